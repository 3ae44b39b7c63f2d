"""CPU: libcdrhead.so builds for sm_100a, loads, exports every symbol include/cdrhead.h
declares, and validates arguments before touching CUDA.  No compute is launched."""
import ctypes as C
import os
import re
import subprocess

import pytest


def _declared(header_path):
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cdr_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(pkg):
    names = _declared(pkg._lib.HEADER_PATH)
    assert len(names) >= 20
    assert sorted(pkg._lib.EXPORTS) == names, "ctypes table and header disagree"
    handle = C.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in cdrhead.h but not exported"
    assert pkg._lib.lib().cdr_abi_version() == 1


def test_struct_layout_matches_header(pkg):
    L = pkg._lib
    assert C.sizeof(L.CdrConvBn) == 6 * 8
    assert C.sizeof(L.CdrWeightPtrs) == 8 + 9 * 48 + 8          # + fusion_hid_ch1 / fusion_hid_ch2
    assert C.sizeof(L.CdrHeadTaps) == 5 * 8


def test_sm100a_cubin_with_bulk_copies(pkg):
    """The library carries sm_100a SASS; the streaming kernel uses bulk async copies (UBLKCP)."""
    out = subprocess.run(["cuobjdump", "-lelf", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass


def test_argument_validation_without_gpu(pkg):
    L = pkg._lib.lib()
    assert L.cdr_pinv(None, 4, 1e-7, None, None) == 1
    assert b"cdr_pinv" in L.cdr_last_error()
    assert L.cdr_softargmax(None, 1, 64, 64, 4.0, None, None) == 1
    assert L.cdr_dlt(None, None, None, None, 1, 19, None, None) == 1
    assert L.cdr_triangulate_u8(None, None, 5, 0, None, None, 1, 19, None, None) == 1
    h = C.c_void_p()
    assert L.cdr_weights_create(None, 0, None, C.byref(h)) == 1
    assert L.cdr_mpjpe_scratch_bytes(10) == (30 + 1025 * 3) * 8
    with pytest.raises(pkg.CdrError):
        pkg._lib.check(L.cdr_argmax(None, 1, 64, 64, 4.0, None, None, None, None))


def test_missing_library_fails_loudly(pkg, monkeypatch):
    L = pkg._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(os.path.dirname(L.LIB_PATH), "nope.so"))
    with pytest.raises(L.CdrError, match="no CPU or PyTorch fallback"):
        L.lib()
