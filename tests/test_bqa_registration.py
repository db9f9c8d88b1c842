"""Drop-in check against the reference package itself (only where /root/reference is mounted, i.e. in the build
container): after ``register_with_bqa()`` the unmodified ``bqa.run_qa`` accepts ``"backend": "b200"``, compiles the
config with bqa's OWN compile step, and executes it on the bqa_b200 engine (kernels = test-only host emulation
here); results must equal bqa's numpy backend on the same config."""
import os
import sys

import numpy as np
import pytest

import instances

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")


@pytest.fixture(scope="module")
def bqa(tmp_path_factory):
    shim = tmp_path_factory.mktemp("shim")
    d = shim / "bqa-0.1.6.dist-info"
    d.mkdir()
    (d / "METADATA").write_text("Metadata-Version: 2.1\nName: bqa\nVersion: 0.1.6\n")
    sys.path[:0] = [REF, str(shim)]
    import bqa
    yield bqa
    sys.path.remove(REF)
    sys.path.remove(str(shim))


@pytest.mark.parametrize("name", ["ring24", "grid4"])
def test_bqa_run_qa_with_b200_backend_equals_numpy_backend(bqa, name):
    from bqa_b200 import _lib, register_with_bqa
    from hostemu.build import build as build_hostemu
    cls = register_with_bqa()
    import bqa.backends as backends
    assert backends.BACKEND_STR_TO_BACKEND["b200"] is cls
    cfg = instances.GOLDEN_CONFIGS[name]()
    want = dict(bqa.run_qa({**cfg, "backend": "numpy"}))
    got = dict(bqa.run_qa({**cfg, "backend": "b200"}, precision="double", _testing_lib=_lib.bind(build_hostemu())))
    assert np.abs(np.array(got["bloch_vectors"]) - np.array(want["bloch_vectors"])).max() < 1e-8
    assert got["measurement_outcomes"] == want["measurement_outcomes"]


def test_b200_backend_has_no_host_arithmetic(bqa):
    """Index tensors of the compile step work; the numerical Tensor methods raise instead of running numpy under the
    name "b200" (no CPU fallback, not even through inheritance)."""
    from bqa_b200 import register_with_bqa
    cls = register_with_bqa()
    idx = cls.make_from_list([3, 1, 2])
    assert idx.numpy.tolist() == [3, 1, 2]
    t = cls.make_from_numpy(np.ones((2, 2, 3, 3), complex))
    for call in (lambda: t.get_density_matrices([t]), lambda: t.pass_msgs([t]), lambda: t.apply_x_gates(0.1),
                 lambda: t.sqrt(), lambda: t.measure(0, 0)):
        with pytest.raises(RuntimeError, match="no per-op host path"):
            call()
