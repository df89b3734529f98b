"""pytest configuration: `gpu` marker, the CPU-emulation build of the kernel sources, library handles."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_LIB = os.path.join(EMUL_DIR, "_build", "libvnet_b200_emul.so")
CSRC = os.path.join(ROOT, "vnet_tensorflow_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


@pytest.fixture(scope="session")
def emul_lib():
    """The product's kernel + engine sources compiled for the CPU with tests/emul/cuda_emul.h.
    Test infrastructure only: the product never loads this library."""
    from vnet_tensorflow_b200 import _ffi
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMUL_DIR, "cuda_emul.h"),
                                                                 os.path.join(ROOT, "include", "vnet_b200.h")]
    if _stale(EMUL_LIB, srcs):
        os.makedirs(os.path.dirname(EMUL_LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-DVNB_EMULATE", "-I" + EMUL_DIR, "-x", "c++", "-shared", "-fPIC",
                        os.path.join(CSRC, "libvnet_b200.cu"), "-o", EMUL_LIB], check=True)
    return _ffi.Library(EMUL_LIB)


@pytest.fixture(scope="session")
def gpu_lib():
    from vnet_tensorflow_b200 import _ffi
    return _ffi.Library()  # raises if libvnet_b200.so is missing: no fallback
