"""The drop-in boundary: the CUDA library builds for sm_100a (nvcc cross-compiles without a GPU), loads, and exports
every symbol include/bqa_b200.h declares; the ctypes binding lists exactly those prototypes; the test-only host
emulation exports the same ABI.  No compute calls here (they need a GPU: tests/test_gpu_parity.py)."""
import ctypes
import os
import re
import shutil

import pytest

from bqa_b200 import _lib
from bqa_b200 import build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bqa_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bqa_b200_\w+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    assert set(declared_symbols()) == set(_lib.EXPORTS)


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_cuda_library_builds_loads_and_exports_every_declared_symbol():
    lib = build_mod.build()
    dll = ctypes.CDLL(lib)
    for name in declared_symbols():
        assert hasattr(dll, name), f"{name} declared in include/bqa_b200.h but not exported by {lib}"
    dll.bqa_b200_version.restype = ctypes.c_int
    assert dll.bqa_b200_version() > 0                      # the CUDA build (the host emulation reports < 0)
    dll.bqa_b200_last_error.restype = ctypes.c_char_p
    assert dll.bqa_b200_last_error() is not None
    # argument validation happens before any CUDA call: safe without a device
    dll.bqa_b200_set_kernel_mode.argtypes = [ctypes.c_int]
    assert dll.bqa_b200_set_kernel_mode(7) != 0 and b"kernel mode" in dll.bqa_b200_last_error()
    assert dll.bqa_b200_set_kernel_mode(0) == 0


def test_cuda_sources_target_sm_100a_only():
    flags = " ".join(build_mod.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and flags.count("-gencode") == 1


def test_host_emulation_exports_the_same_abi():
    """the per-class fused entry points (what the engine calls on the host emulation); the raw-op entry points
    bqa_b200_t_* of the backend class and the all-classes-in-one-launch entry points exist in the CUDA library only"""
    from hostemu.build import build as build_hostemu
    dll = ctypes.CDLL(build_hostemu())
    for name in declared_symbols():
        if not name.startswith("bqa_b200_t_") and not name.endswith("_classes") and name not in _lib._CUDA_ONLY:
            assert hasattr(dll, name), name
