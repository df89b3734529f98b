"""The product library's boundary (no GPU needed): libvnet_b200.so loads, exports every symbol that
include/vnet_b200.h declares, and fails loudly -- never falls back -- when no B200 is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from vnet_tensorflow_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    if not os.path.exists(_ffi.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    return _ffi.Library(_ffi.DEFAULT_LIB)


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vnb_[a-z0-9_]+)\s*\(", text)))


def test_header_ffi_and_shared_object_agree(product_lib):
    declared = _header_symbols()
    assert declared == sorted(_ffi.EXPORTED_SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", product_lib.path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vnb_[a-z0-9_]+)", nm))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert "sm_100a" in product_lib.version()


def test_sass_is_blackwell_native(product_lib):
    """The shipped .so carries sm_100a SASS whose convolution kernels are tcgen05 / TMA / TMEM code: every instance of
    the three tensor-core kernels issues UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld) and UTCBAR
    (tcgen05.commit); profiles/r02_sass_summary.txt is the committed listing of the same counts."""
    import importlib.util
    out = subprocess.run(["cuobjdump", "-lelf", product_lib.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    spec = importlib.util.spec_from_file_location("sass_summary", os.path.join(ROOT, "tools", "sass_summary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    per = mod.summarise(product_lib.path)
    names = mod.demangle(list(per))
    seen = {"conv5_tc_kernel": 0, "conv5_col_kernel": 0, "wgrad5_tc_kernel": 0}
    for k, c in per.items():
        for kern in seen:
            if kern + "<" in names[k]:
                seen[kern] += 1
                assert c["UTCHMMA"] >= 7 and c["UTMALDG"] >= 8 and c["LDTM"] >= 1 and c["UTCBAR"] >= 4, (names[k], dict(c))
                assert c["HMMA"] == 0, names[k]          # no legacy mma.sync in the tcgen05 kernels
    assert seen["conv5_tc_kernel"] >= 8 and seen["conv5_col_kernel"] == 2 and seen["wgrad5_tc_kernel"] == 8, seen
    # the 2^3 stride-2 kernels (gather / depth-to-space scatter with TMA store and reduce-add, filter gradient) and the
    # deep-level filter gradient are tcgen05 / TMA code too
    short = {re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("vnb::", ""): c for k, c in per.items()}
    k2 = short["k2_tc_kernel"]
    assert k2["UTCHMMA"] >= 6 and k2["UTMALDG"] >= 3 and k2["UTMASTG"] >= 1 and k2["UTMAREDG"] >= 1 and k2["LDTM"] >= 2 and k2["HMMA"] == 0
    for kern in ("k2_wgrad_tc_kernel", "wgrad5_deep_kernel"):
        c = short[kern]
        assert c["UTCHMMA"] >= 3 and c["UTMALDG"] >= 2 and c["LDTM"] >= 1 and c["UTCBAR"] >= 2 and c["HMMA"] == 0, (kern, dict(c))
    total = sum(c["UTCHMMA"] for c in per.values())
    assert total >= 500, total


def test_create_fails_loudly_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the no-GPU failure path is exercised in the CPU container")
    cfg = _ffi.VnbConfig()
    cfg.in_channels, cfg.num_classes, cfg.num_channels, cfg.num_levels = 1, 2, 16, 1
    cfg.num_convolutions[0] = 1
    cfg.bottom_convolutions = 1
    for i in range(3):
        cfg.patch_shape[i] = 8
    cfg.max_batch = 1
    h = C.c_void_p()
    rc = product_lib.vnb_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -3  # VNB_ERR_CUDA
    assert b"no CPU fallback" in product_lib.vnb_last_error()
    assert not h.value
    from vnet_tensorflow_b200.engine import VNetEngine
    with pytest.raises(_ffi.VnbError):
        VNetEngine(num_classes=2, patch_shape=(16, 16, 16), library=product_lib)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        _ffi.Library(str(tmp_path / "libvnet_b200.so"))
