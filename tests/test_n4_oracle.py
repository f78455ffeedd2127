"""SURVEY 8(f) N4 on the CPU: the oracle restatements of lbp_transform, local_maxima_filter, fast_detector9_blockwise_rank and the
oriented LK matcher against the reference's own test vector (tests/lbp.cc) and against the reference's headers compiled in
oracle/_ref (lbp_transform.hh, fast.hpp:555-575, lk.hh:180-317; blockwise_rank is not instantiable in the reference - see the oracle)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import oracle as orc
from tests import scenes
from tests.oracle_ops import oracle_grad_pyramid, oracle_pyramid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libvppref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
I = C.POINTER(orc.VoImg)


@pytest.fixture(scope="module")
def o(built):
    return orc.load()


@pytest.fixture(scope="module")
def ref(built):
    r = C.CDLL(REF)
    r.vppref_lbp_u8.argtypes = [I, I]
    r.vppref_local_maxima_filter.argtypes = [I]
    r.vppref_lk_match_oriented.argtypes = [I, I, I, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    return r


def test_lbp_known_answer(o):
    """tests/lbp.cc:9-38: the 3x3 image of the reference's test, lbp(1,1) == 0b10101110"""
    v = np.array([[0, 2, 2], [2, 1, 0], [2, 0, 2]], np.uint8)
    h = orc.HostImage(3, 3, "u8", border=1, data=v)
    out = orc.HostImage(3, 3, "u8")
    o.vo_lbp_u8(h.ptr(), out.ptr())
    assert out.get()[1, 1] == 0b10101110


@needs_ref
@pytest.mark.parametrize("shape", [(3, 3), (37, 53), (64, 128), (5, 301)])
def test_lbp_equals_reference(ref, o, shape):
    img = np.random.default_rng(shape[1]).integers(0, 6, shape, dtype=np.uint8) * 40
    h = orc.HostImage(shape[0], shape[1], "u8", border=1, data=img, fill_border="mirror")
    a, b = orc.HostImage(shape[0], shape[1], "u8"), orc.HostImage(shape[0], shape[1], "u8")
    ref.vppref_lbp_u8(h.ptr(), a.ptr())
    o.vo_lbp_u8(h.ptr(), b.ptr())
    assert np.array_equal(a.get(), b.get()) and len(np.unique(b.get())) > 3


def lmf_scenes(shape, pix, seed):
    """images that exercise the in-place dependence: plateaus, ramps (every pixel's left / upper neighbour is larger and gets zeroed first),
    sparse score-like images and dense noise"""
    r = np.random.default_rng(seed)
    nr, nc = shape
    hi = 250 if pix == "u8" else 100000
    out = [r.integers(0, hi, shape), r.integers(0, 4, shape), np.where(r.random(shape) < 0.05, r.integers(1, hi, shape), 0)]
    rr, cc = np.meshgrid(np.arange(nr), np.arange(nc), indexing="ij")
    out.append((hi - 1 - (rr + cc) % hi))             # decreasing along rows and columns: long chains of dependent decisions
    out.append(((rr * 3 + cc * 2) % 7) * (hi // 8))    # periodic ramps
    out.append(np.full(shape, 9))
    return [a.astype(np.uint8 if pix == "u8" else np.int32) for a in out]


@needs_ref
@pytest.mark.parametrize("pix", ["u8", "i32"])
@pytest.mark.parametrize("shape", [(9, 14), (40, 67), (64, 96)])
def test_local_maxima_filter_serial_equals_reference(ref, o, shape, pix):
    for i, img in enumerate(lmf_scenes(shape, pix, 3)):
        a = orc.HostImage(shape[0], shape[1], pix, border=1, data=img, fill_border="value")
        b = orc.HostImage(shape[0], shape[1], pix, border=1, data=img, fill_border="value")
        ref.vppref_local_maxima_filter(a.ptr())
        o.vo_local_maxima_filter(b.ptr())
        assert np.array_equal(a.get(True), b.get(True)), (i, pix)
        assert (b.get() != img).any() or i == 2


def oriented_case(nr, nc, n, seed, ws):
    f1, f2, pts = scenes.lk_pair(nr, nc, n, seed=seed, shift=(1.3, -0.8), margin=ws + 6)
    r = np.random.default_rng(seed)
    ang1, ang2 = r.uniform(0, 2 * np.pi, len(pts)), r.uniform(0, 2 * np.pi, len(pts))
    ang2[::2] = ang1[::2]  # half of the points search along the template's own direction
    d1 = np.stack([np.cos(ang1), np.sin(ang1)], axis=1).astype(np.float32)
    d2 = np.stack([np.cos(ang2), np.sin(ang2)], axis=1).astype(np.float32)
    d1[::5], d2[::5] = (0.0, 1.0), (0.0, 1.0)  # the axis-aligned window
    pred = r.uniform(-1.5, 1.5, (len(pts), 2)).astype(np.float32)
    pts = pts.copy()
    pts[:4] = [[1.5, 2.5], [nr - 2.0, nc - 3.0], [0.0, nc / 2], [nr / 2, 1.0]]  # windows that leave the domain
    return f1, f2, np.ascontiguousarray(pts, np.float32), pred, d1, d2


@needs_ref
@pytest.mark.parametrize("ws,max_iter,max_step", [(5, 10, 1.0), (7, 21, 0.5), (9, 15, 100.0), (11, 4, 2.0)])
def test_oriented_lk_equals_reference(ref, o, ws, max_iter, max_step):
    nr, nc = 151, 203
    f1, f2, pts, pred, d1, d2 = oriented_case(nr, nc, 300, ws, ws)
    n = len(pts)
    A = orc.HostImage(nr, nc, "u8", border=3, data=f1, fill_border="mirror")
    B = orc.HostImage(nr, nc, "u8", border=3, data=f2, fill_border="mirror")
    G = oracle_grad_pyramid([A], "vfloat2", 3, o)[0]
    fa, ea = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    fb, eb = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    ref.vppref_lk_match_oriented(A.ptr(), B.ptr(), G.ptr(), ws, 1e-3, max_iter, 0.01, max_step, pts.ctypes.data, pred.ctypes.data, d1.ctypes.data,
                                 d2.ctypes.data, n, fa.ctypes.data, ea.ctypes.data)
    o.vo_lk_match_oriented_u8(A.ptr(), B.ptr(), G.ptr(), 1, ws, 1e-3, max_iter, 0.01, max_step, pts.ctypes.data, pred.ctypes.data, d1.ctypes.data,
                              d2.ctypes.data, n, fb.ctypes.data, eb.ctypes.data)
    # points whose rotated windows leave the domain use uninitialised as[] / gs[] in the reference (zero here): the first four
    same = (fa.view(np.int32) == fb.view(np.int32)).all(axis=1) & (ea.view(np.int32) == eb.view(np.int32))
    assert same[4:].all(), (np.flatnonzero(~same), fa[~same][:4], fb[~same][:4])
    ok = eb < 1e30
    assert ok.sum() > n // 2 and (np.abs(fb[ok] - np.array([1.3, -0.8])).max(axis=1) < 1.0).mean() > 0.5


def test_blockwise_rank_properties(o):
    """not instantiable in the reference: the restatement is checked against its own definition - every record is a strict 3x3 maximum of the
    raw score image inside its block, ranks are 0..k-1 in decreasing score order, and with max_points == 1 the rule keeps the LAST of
    the increasing maxima of the raster scan (a candidate replaces a smaller slot)"""
    img = scenes.rectangles_scene(120, 161, seed=3)
    h = orc.HostImage(120, 161, "u8", border=3, data=img, fill_border="mirror")
    cap = img.size
    for bs, mp in ((10, 3), (16, 1), (7, 16)):
        k3, sc = np.zeros((cap, 3), np.int32), np.zeros(cap, np.int32)
        n = o.vo_fast9_blockwise_rank(h.ptr(), 15, None, bs, mp, 0, k3.ctypes.data, sc.ctypes.data, cap)
        assert n > 20
        k3, sc = k3[:n], sc[:n]
        S = np.zeros((122, 163), np.int64)
        ka = np.zeros((cap, 2), np.int32)
        na = o.vo_fast9_u8(h.ptr(), 15, None, 0, bs, 0, ka.ctypes.data, None, cap)
        for (r, c) in ka[:na]:
            S[r + 1, c + 1] = o.vo_fast9_score(h.ptr(), 15, int(r), int(c))
        for (r, c, k), s in zip(k3, sc):
            win = S[r:r + 3, c:c + 3].copy()
            assert s == win[1, 1] and s > 0
            win[1, 1] = -1
            assert s > win.max() and 0 <= k < mp
        blocks = (k3[:, 0] // bs) * 1000 + k3[:, 1] // bs
        assert (np.diff(blocks) >= 0).all()  # blocks in raster order
        for b in np.unique(blocks):
            m = blocks == b
            assert list(k3[m, 2]) == list(range(m.sum())) or mp > 1 and (np.diff(k3[m, 2]) > 0).all()
            assert (np.diff(sc[m]) <= 0).all()


# ---- property-based pins (hypothesis, derandomized like tests/test_oracle_vs_ref_property.py): random geometries and value ranges
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

PSET = dict(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@needs_ref
@settings(**PSET)
@given(nr=st.integers(1, 30), nc=st.integers(1, 70), levels=st.integers(1, 255), aligned=st.sampled_from([1, 4, 16, 128]), seed=st.integers(0, 2 ** 16))
def test_lbp_any_geometry(ref, o, nr, nc, levels, aligned, seed):
    img = np.random.default_rng(seed).integers(0, levels + 1, (nr, nc)).astype(np.uint8)
    h = orc.HostImage(nr, nc, "u8", border=1, aligned=aligned, data=img, fill_border="value", border_value=seed % 256)
    a, b = orc.HostImage(nr, nc, "u8", aligned=aligned), orc.HostImage(nr, nc, "u8", aligned=aligned)
    ref.vppref_lbp_u8(h.ptr(), a.ptr())
    o.vo_lbp_u8(h.ptr(), b.ptr())
    assert np.array_equal(a.get(), b.get())


@needs_ref
@settings(**PSET)
@given(nr=st.integers(1, 24), nc=st.integers(1, 40), levels=st.integers(1, 40), pix=st.sampled_from(["u8", "i32"]), signed=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_local_maxima_filter_any_image(ref, o, nr, nc, levels, pix, signed, seed):
    """few grey levels = many ties and plateaus: the strict comparisons and the in-place order decide everything"""
    lo = -levels if (signed and pix == "i32") else 0
    img = np.random.default_rng(seed).integers(lo, levels + 1, (nr, nc)).astype(np.uint8 if pix == "u8" else np.int32)
    bv = seed % 5
    a = orc.HostImage(nr, nc, pix, border=1, data=img, fill_border="value", border_value=bv)
    b = orc.HostImage(nr, nc, pix, border=1, data=img, fill_border="value", border_value=bv)
    ref.vppref_local_maxima_filter(a.ptr())
    o.vo_local_maxima_filter(b.ptr())
    assert np.array_equal(a.get(True), b.get(True))
