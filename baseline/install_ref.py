"""Installs the UNMODIFIED reference (LuchnikovI/bqa, /root/reference) into baseline/_ref (git-ignored, shipped to the
GPU box by gpurun) so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference's own
numpy backend instead of the oracle port, and so that the GPU tests can drive ``bqa.state`` with the b200 backend.

    python baseline/install_ref.py          (build container only: needs /root/reference)

The reference builds with poetry-core, which is not in this image (``pip install --no-index --no-build-isolation
--find-links /opt/wheelhouse --target baseline/_ref /root/reference`` fails with "No module named 'poetry'").  The
package is pure Python, so the install is done from a copy under /tmp whose build metadata is replaced by an
equivalent setuptools one (same name, version and ``src`` layout); no file of the package itself is touched.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "bqa"
version = "{version}"
description = "reference install for benchmarking (unmodified sources)"

[tool.setuptools.packages.find]
where = ["src"]
"""


def reference_version() -> str:
    with open(os.path.join(REFERENCE, "pyproject.toml")) as f:
        for line in f:
            if line.startswith("version"):
                return line.split("=")[1].strip().strip('"')
    raise RuntimeError("no version in the reference's pyproject.toml")


def installed() -> bool:
    return os.path.exists(os.path.join(TARGET, "bqa", "state.py"))


def install(force: bool = False) -> str:
    if installed() and not force:
        return TARGET
    if not os.path.isdir(REFERENCE):
        raise RuntimeError(f"{REFERENCE} is not present: the reference can only be installed in the build container")
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "bqa_src")
        shutil.copytree(REFERENCE, work, ignore=shutil.ignore_patterns(".git", "poetry.lock"))
        with open(os.path.join(work, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT.format(version=reference_version()))
        shutil.rmtree(TARGET, ignore_errors=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                               "--quiet", "--target", TARGET, work])
    assert installed()
    return TARGET


def add_to_path() -> bool:
    """Makes ``import bqa`` resolve to baseline/_ref; False when the install is absent."""
    if not installed():
        return False
    if TARGET not in sys.path:
        sys.path.insert(0, TARGET)
    return True


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
