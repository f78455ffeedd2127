"""CPU tests: the C-ABI library loads without a GPU, exports every symbol include/vppb.h declares,
and its host-only arithmetic (layout, workspace sizes, descriptors) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi(built):
    from vpp_b200 import capi

    return capi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vppb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vppb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(capi):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(capi.lib, n), "libvppb.so does not export %s" % n
        assert n in capi.PROTOTYPES, "capi.py has no prototype for %s" % n
    assert sorted(capi.PROTOTYPES) == names
    assert capi.lib.vppb_version() == 100


def test_layout_matches_reference_formula(capi):
    o = orc.load()
    for nr, nc, e, b, al in [(100, 200, 4, 1, 256), (1080, 1920, 3, 2, 128), (1080, 1920, 3, 2, 32), (2160, 3840, 1, 3, 128),
                             (5, 10, 4, 2, 1), (7, 9, 8, 5, 16), (512, 512, 4, 0, 128)]:
        p1, t1, o1 = C.c_int32(), C.c_int64(), C.c_int64()
        p2, t2, o2 = C.c_int(), C.c_int64(), C.c_int64()
        assert capi.lib.vppb_layout(nr, nc, e, b, al, C.byref(p1), C.byref(t1), C.byref(o1)) == 0
        assert o.vo_layout(nr, nc, e, b, al, C.byref(p2), C.byref(t2), C.byref(o2)) == 0
        assert (p1.value, t1.value, o1.value) == (p2.value, t2.value, o2.value)
        assert p1.value % al == 0 and (o1.value - b * p1.value) % al == 0  # row starts are aligned
    assert capi.lib.vppb_layout(0, 5, 4, 0, 128, None, None, None) == capi.VPPB_E_ARG
    assert b"invalid geometry" in capi.lib.vppb_last_error()


def test_descriptor_helpers_without_gpu(capi):
    # vppb_wrap / vppb_subimage are pure pointer arithmetic
    img = capi.VppbImg()
    assert capi.lib.vppb_wrap(C.byref(img), 0x10000, 100, 200, 4, 1, 256) == 0
    assert img.pitch == 1536 and img.base == 0x10000 + (256 - 4) + 1 * 1536 + 4 and img.alloc is None
    assert capi.lib.vppb_wrap(C.byref(img), 0x10004, 100, 200, 4, 1, 256) == capi.VPPB_E_ARG
    capi.lib.vppb_wrap(C.byref(img), 0x10000, 100, 200, 4, 1, 256)
    sub = capi.VppbImg()
    assert capi.lib.vppb_subimage(C.byref(img), 10, 10, 12, 15, C.byref(sub)) == 0
    assert (sub.nrows, sub.ncols) == (3, 6) and sub.base == img.base + 10 * img.pitch + 40
    assert capi.lib.vppb_subimage(C.byref(img), 10, 10, 200, 15, C.byref(sub)) == capi.VPPB_E_ARG
    assert capi.lib.vppb_fast9_workspace_bytes(2160, 3840, 10) > 2 * 2160 * 120 * 4
    assert capi.lib.vppb_halo_bytes(C.byref(img), 2) == 2 * 202 * 4


def test_no_cpu_fallback(capi):
    """Without a CUDA device the compute entries must fail loudly, never compute on the host."""
    n = C.c_int(0)
    rc = capi.lib.vppb_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    img = capi.VppbImg()
    assert capi.lib.vppb_alloc(C.byref(img), 16, 16, 4, 0, 128) == capi.VPPB_E_CUDA
    assert b"CUDA error" in capi.lib.vppb_last_error()


def test_host_mirror_logic(capi):
    from vpp_b200.image import Box2d, make_box2d, layout

    b = make_box2d(10, 20)
    assert (b.nrows, b.ncols) == (10, 20) and b.has((9, 19)) and not b.has((10, 0))
    assert b == Box2d((0, 0), (9, 19))
    assert layout(1080, 1920, 3, 2, 128)[0] == 6016


def test_box_division_magic():
    """The two exact divisions by 25 used by k_box5_bytes_tma (vpp_b200/csrc/box.cu), checked exhaustively:
    lo lane  (s * 671089) >> 24 == s // 25  for 0 <= s <= 6375 (= 25 * 255), product < 2^32;
    hi lane  byte 1 of mulhi(hi << 16 | lo, 671089) == hi // 25 for every lo in the same range."""
    M = 671089
    s = np.arange(6376, dtype=np.uint64)
    assert ((s * M) >> 24 == s // 25).all() and int((s * M).max()) < 2 ** 32
    hi = s[:, None]
    lo = np.arange(6376, dtype=np.uint64)[None, :]
    q = ((((hi << 16) | lo) * M) >> 32 >> 8) & 0xFF
    assert (q == hi // 25).all()


def test_gray_division_magic():
    """colorspace.cu divides the channel sum by 3 with (s * 43691) >> 17: exact for every s = r + g + b in 0..765
    (the tempting (s * 171) >> 9 is not: it is off by one from s = 512 on)."""
    import numpy as np

    s = np.arange(0, 766, dtype=np.int64)
    assert np.array_equal((s * 43691) >> 17, s // 3)
    assert not np.array_equal((s * 171) >> 9, s // 3)
