"""SURVEY 8(f) N4 through the C-ABI on the GPU: lbp_transform, local_maxima_filter (serial in-place semantics), fast_detector9_blockwise_rank
and the oriented LK matcher against the oracle (which tests/test_n4_oracle.py pins to the reference's headers / test vector).
tests/test_emulated_parity.py re-runs every test of this module on the CPU emulator."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle as orc
from tests import scenes
from tests.oracle_ops import oracle_grad_pyramid
from tests.test_gpu_parity import vpp  # noqa: F401  (the module-scoped CUDA fixture)
from tests.test_n4_oracle import lmf_scenes, oriented_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 3), (37, 53), (64, 128), (5, 301), (131, 1000)])
def test_lbp_transform(vpp, shape):
    o = orc.load()
    if shape == (3, 3):
        img = np.array([[0, 2, 2], [2, 1, 0], [2, 0, 2]], np.uint8)  # tests/lbp.cc:9-38
    else:
        img = (np.random.default_rng(shape[1]).integers(0, 6, shape, dtype=np.uint8) * 40).astype(np.uint8)
    A = vpp.Image2d.from_host(img, "u8", border=1)
    vpp.fill_border_mirror(A)
    B = vpp.lbp_transform(A)
    h = orc.HostImage(shape[0], shape[1], "u8", border=1, data=img, fill_border="mirror")
    ref = orc.HostImage(shape[0], shape[1], "u8")
    o.vo_lbp_u8(h.ptr(), ref.ptr())
    assert np.array_equal(B.download(), ref.get())
    if shape == (3, 3):
        assert B.download()[1, 1] == 0b10101110
    # a sub-image view (rows not 16-byte aligned): the byte path
    if shape[1] > 40:
        sub = vpp.Image2d.from_host(img[1:-1, 3:-2], "u8", border=1, aligned=4)
        vpp.fill_border_mirror(sub)
        hs = orc.HostImage(shape[0] - 2, shape[1] - 5, "u8", border=1, data=img[1:-1, 3:-2], fill_border="mirror")
        rs = orc.HostImage(shape[0] - 2, shape[1] - 5, "u8")
        o.vo_lbp_u8(hs.ptr(), rs.ptr())
        assert np.array_equal(vpp.lbp_transform(sub, vpp.Image2d(shape[0] - 2, shape[1] - 5, "u8", aligned=4)).download(), rs.get())


@pytest.mark.parametrize("pix", ["u8", "i32"])
@pytest.mark.parametrize("shape", [(9, 14), (40, 67), (64, 96)])
def test_local_maxima_filter_serial_semantics(vpp, shape, pix):
    o = orc.load()
    for i, img in enumerate(lmf_scenes(shape, pix, 5)):
        A = vpp.Image2d.from_host(img, pix, border=1)
        vpp.fill_border_with_value(A, 0)
        h = orc.HostImage(shape[0], shape[1], pix, border=1, data=img, fill_border="value")
        vpp.local_maxima_filter(A)
        o.vo_local_maxima_filter(h.ptr())
        assert np.array_equal(A.download(with_border=True), h.get(True)), (i, pix)
    if pix == "i32":  # negative values: signed comparisons, a zeroed neighbour is larger than a negative pixel
        img = np.random.default_rng(2).integers(-50, 50, shape).astype(np.int32)
        A = vpp.Image2d.from_host(img, pix, border=1)
        vpp.fill_border_with_value(A, 7)
        h = orc.HostImage(shape[0], shape[1], pix, border=1, data=img, fill_border="value", border_value=7)
        vpp.local_maxima_filter(A)
        o.vo_local_maxima_filter(h.ptr())
        assert np.array_equal(A.download(with_border=True), h.get(True))


@pytest.mark.parametrize("shape,th,bs,mp,ring,maskval", [((120, 161), 15, 10, 3, "reference", None), ((97, 203), 8, 16, 1, "true", None),
                                                        ((120, 161), 15, 7, 16, "reference", 0xFF), ((64, 70), 5, 9, 4, "reference", 0x01)])
def test_fast9_blockwise_rank(vpp, shape, th, bs, mp, ring, maskval):
    o = orc.load()
    img = scenes.rectangles_scene(shape[0], shape[1], seed=shape[1])
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    h = orc.HostImage(shape[0], shape[1], "u8", border=3, data=img, fill_border="mirror")
    M = hm = None
    if maskval is not None:
        m = np.full(shape, maskval, np.uint8)
        m[::5, :] = 0
        M, hm = vpp.Image2d.from_host(m, "u8"), orc.HostImage(shape[0], shape[1], "u8", data=m)
    sc = []
    got = vpp.fast9_blockwise_rank(G, th, block_size=bs, max_points_per_block=mp, mask=M, scores=sc, ring=ring)
    cap = img.size
    k3, s = np.zeros((cap, 3), np.int32), np.zeros(cap, np.int32)
    n = o.vo_fast9_blockwise_rank(h.ptr(), th, hm.ptr() if hm else None, bs, mp, 0 if ring == "reference" else 1, k3.ctypes.data, s.ctypes.data, cap)
    assert n > 10 and len(got) == n
    assert np.array_equal(got, k3[:n]) and np.array_equal(np.array(sc, np.int32), s[:n])
    if mp > 1:
        assert (got[:, 2] > 0).any()


@pytest.mark.parametrize("ws,max_iter,max_step,grad", [(5, 10, 1.0, "vfloat2"), (7, 21, 0.5, "vfloat2"), (9, 15, 100.0, "vint2"), (11, 4, 2.0, "vfloat2"),
                                                      (15, 6, 1.5, "vfloat2")])
def test_oriented_lk_matcher(vpp, ws, max_iter, max_step, grad):
    """lk.hh:180-317: failure codes identical, displacements and errors within 1e-4 relative of the oracle (observed: bit-identical)"""
    o = orc.load()
    nr, nc = 151, 203
    f1, f2, pts, pred, d1, d2 = oriented_case(nr, nc, 300 if ws <= 7 else 90, ws, ws)  # (the CPU emulator runs this test too: one fiber per thread)
    n = len(pts)
    A = orc.HostImage(nr, nc, "u8", border=3, data=f1, fill_border="mirror")
    B = orc.HostImage(nr, nc, "u8", border=3, data=f2, fill_border="mirror")
    Gh = oracle_grad_pyramid([A], grad, 3, o)[0]
    rf, re = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    o.vo_lk_match_oriented_u8(A.ptr(), B.ptr(), Gh.ptr(), 1 if grad == "vfloat2" else 0, ws, 1e-3, max_iter, 0.01, max_step, pts.ctypes.data, pred.ctypes.data,
                              d1.ctypes.data, d2.ctypes.data, n, rf.ctypes.data, re.ctypes.data)
    dA, dB = vpp.Image2d.from_host(f1, "u8", border=3), vpp.Image2d.from_host(f2, "u8", border=3)
    vpp.fill_border_mirror(dA)
    vpp.fill_border_mirror(dB)
    dG = vpp.Image2d(nr, nc, grad, border=3)
    vpp.scharr(dA, dG)
    vpp.fill_border_mirror(dG)
    assert np.array_equal(dG.download(with_border=True).view(np.int32), Gh.get(True).view(np.int32))
    gf, ge = vpp.oriented_lk_match(dA, dB, dG, pts, pred, d1, d2, ws, 1e-3, max_iter, 0.01, max_step)
    fail_ref, fail_got = re > 1e30, ge > 1e30
    assert np.array_equal(fail_ref, fail_got)
    assert np.array_equal(gf[fail_got].view(np.int32), rf[fail_ref].view(np.int32))  # (-1,-1) / (0,0) codes
    ok = ~fail_ref
    assert ok.sum() > n // 3
    assert np.abs(gf[ok] - rf[ok]).max() <= 1e-4 * max(1.0, np.abs(rf[ok]).max())
    assert np.abs(ge[ok] - re[ok]).max() <= 1e-4 * max(1.0, np.abs(re[ok]).max())
    assert (gf.view(np.int32) == rf.view(np.int32)).all() and (ge.view(np.int32) == re.view(np.int32)).all()
