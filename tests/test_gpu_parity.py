"""GPU parity: every CUDA entry point (through the C-ABI via vpp_b200.capi) against the CPU oracle
on the same seeded inputs.  Integer / byte / index work must be bit-exact; LK displacements within
1e-4 relative (north_star), in practice bit-exact because the kernel replays the reference's
float evaluation order."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle as orc
from tests import scenes
from tests.oracle_ops import oracle_grad_pyramid, oracle_lk, oracle_lucas_kanade, oracle_pyramid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vpp(gpu):
    import vpp_b200

    return vpp_b200


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------ containers
def test_layout_alignment_and_roundtrip(vpp):
    # tests/imageNd.cc:46-50: first pixel of each row and the pitch are aligned
    for (nr, nc, pix, b, al) in [(100, 200, "i32", 1, 256), (5, 10, "i32", 2, 128), (37, 61, "vuchar3", 2, 128), (1080, 1920, "u8", 3, 128)]:
        img = vpp.Image2d(nr, nc, pix, border=b, aligned=al)
        assert img.desc.base % al == 0 and img.pitch % al == 0
        a = rng(1).integers(0, 255, img._host_shape(True)).astype(img.dtype)
        img.upload(a, with_border=True)
        assert np.array_equal(img.download(with_border=True), a)
        assert np.array_equal(img.download(), a[b:b + nr, b:b + nc] if b else a)


def test_subimage_aliases_pixels(vpp):
    # tests/imageNd.cc:74-83
    a = rng(2).integers(0, 1000, (100, 200), dtype=np.int32)
    img = vpp.Image2d.from_host(a, "i32")
    sub = img | vpp.Box2d((10, 10), (12, 15))
    assert (sub.nrows, sub.ncols) == (3, 6)
    assert np.array_equal(sub.download(), a[10:13, 10:16])
    vpp.fill(sub, 7)
    a[10:13, 10:16] = 7
    assert np.array_equal(img.download(), a)


# ------------------------------------------------------------------ pixel_wise named kernels
@pytest.mark.parametrize("shape", [(512, 512), (1, 1), (3, 5), (101, 203), (64, 1024), (1080, 1920)])
def test_pixel_wise_add_bit_exact(vpp, shape):
    r = rng(42)
    b = r.integers(0, 2 ** 30, shape, dtype=np.int32)
    c = r.integers(0, 2 ** 30, shape, dtype=np.int32)
    A = vpp.Image2d(*shape, "i32")
    vpp.fill(A, 0)
    vpp.pixel_wise_add(A, vpp.Image2d.from_host(b, "i32"), vpp.Image2d.from_host(c, "i32"))
    ha, hb, hc = orc.HostImage(*shape, "i32"), orc.HostImage(*shape, "i32", data=b), orc.HostImage(*shape, "i32", data=c)
    orc.load().vo_pw_add_i32(ha.ptr(), hb.ptr(), hc.ptr())
    assert np.array_equal(A.download(), ha.get())
    assert np.array_equal(A.download(), b + c)  # benchmarks/image_add.cc:21-28 check()


def test_pixel_wise_add_wraps_and_views(vpp):
    r = rng(3)
    b = r.integers(-2 ** 31, 2 ** 31 - 1, (40, 70), dtype=np.int32)
    c = r.integers(-2 ** 31, 2 ** 31 - 1, (40, 70), dtype=np.int32)
    B, Cc = vpp.Image2d.from_host(b, "i32", border=1), vpp.Image2d.from_host(c, "i32", border=3)
    A = vpp.Image2d(40, 70, "i32", border=2)
    vpp.fill_with_border(A, -5)
    vpp.pixel_wise_add(A, B, Cc)
    exp = (b.astype(np.int64) + c).astype(np.int32)
    got = A.download(with_border=True)
    assert np.array_equal(got[2:-2, 2:-2], exp)
    got[2:-2, 2:-2] = -5
    assert (got == -5).all()  # the border frame is never touched
    # unaligned views (odd column offset) take the scalar path
    box = vpp.Box2d((3, 5), (30, 61))
    vpp.fill(A, 0)
    vpp.pixel_wise_add(A | box, B | box, Cc | box)
    assert np.array_equal(A.download()[3:31, 5:62], exp[3:31, 5:62])


@pytest.mark.parametrize("pix,val", [("u8", 42), ("i32", -123456), ("vuchar3", (1, 2, 3)), ("vint2", (7, -9)), ("f32", 1.5)])
def test_fill_variants(vpp, pix, val):
    img = vpp.Image2d(37, 53, pix, border=2)
    vpp.fill_with_border(img, np.zeros(1)[0])
    vpp.fill(img, val)
    a = img.download(with_border=True)
    inner = a[2:-2, 2:-2]
    assert (inner == np.asarray(val, dtype=img.dtype)).all()
    a[2:-2, 2:-2] = 0
    assert (a == 0).all()
    vpp.fill_with_border(img, val)
    assert (img.download(with_border=True) == np.asarray(val, dtype=img.dtype)).all()


@pytest.mark.parametrize("pix", ["u8", "i32", "vuchar3", "vint2"])
@pytest.mark.parametrize("mode", ["value", "mirror", "closest"])
def test_border_fills(vpp, pix, mode):
    nr, nc, b = 5, 10, 2
    dt, ch = orc.PIXEL_TYPES[pix]
    a = rng(5).integers(0, 100, (nr, nc) + ((ch,) if ch > 1 else ())).astype(dt)
    img = vpp.Image2d.from_host(a, pix, border=b, aligned=128)
    vpp.fill_border_with_value(img, 0)
    h = orc.HostImage(nr, nc, pix, border=b, data=a)
    o = orc.load()
    if mode == "value":
        v = np.full(ch, 6, dtype=dt)
        vpp.fill_border_with_value(img, 6)
        o.vo_fill_border_value(h.ptr(), v.ctypes.data)
    elif mode == "mirror":
        vpp.fill_border_mirror(img)
        o.vo_fill_border_mirror(h.ptr())
    else:
        vpp.fill_border_closest(img)
        o.vo_fill_border_closest(h.ptr())
    assert np.array_equal(img.download(with_border=True), h.get(with_border=True))


def test_border_closest_closed_form(vpp):
    # tests/border.cc:36-60
    img = vpp.Image2d(5, 10, "i32", border=2, aligned=128)
    rr, cc = np.meshgrid(np.arange(5), np.arange(10), indexing="ij")
    img.upload(((rr + cc) % 10).astype(np.int32))
    vpp.fill_border_closest(img)
    got = img.download(with_border=True)
    r2, c2 = np.meshgrid(np.arange(-2, 7), np.arange(-2, 12), indexing="ij")
    assert np.array_equal(got, (np.clip(r2, 0, 4) + np.clip(c2, 0, 9)) % 10)


def test_copy_clone_sum(vpp):
    a = rng(6).integers(-100, 100, (100, 200), dtype=np.int32)
    img = vpp.Image2d.from_host(a, "i32", border=1)
    vpp.fill_border_mirror(img)
    cl = vpp.clone(img, border=3)  # tests/imageNd.cc:58-71
    got = cl.download(with_border=True)
    assert np.array_equal(got[2:-2, 2:-2], img.download(with_border=True))
    dst = vpp.Image2d(100, 200, "i32")
    vpp.copy(img, dst)
    assert np.array_equal(dst.download(), a)
    assert vpp.sum(img) == int(a.sum())
    # tests/sum.cc:11-15: char image with wrapping counter
    ch = (np.arange(100 * 200) % 256).astype(np.uint8).view(np.int8).reshape(100, 200)
    assert vpp.sum(vpp.Image2d.from_host(ch, "i8")) == int(ch.astype(np.int64).sum())
    u = rng(7).integers(0, 256, (33, 77), dtype=np.uint8)
    assert vpp.sum(vpp.Image2d.from_host(u, "u8")) == int(u.sum())


# ------------------------------------------------------------------ 5x5 box
@pytest.mark.parametrize("shape", [(270, 480), (1080, 1920), (7, 9), (64, 341), (57, 342), (200, 1000)])
def test_box5x5_vuchar3_bit_exact(vpp, shape):
    src = rng(11).integers(0, 256, shape + (3,), dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "vuchar3", border=2)
    vpp.fill_border_mirror(S)
    D = vpp.Image2d(*shape, "vuchar3")
    vpp.fill(D, 0)
    vpp.box5x5(S, D)
    hs = orc.HostImage(*shape, "vuchar3", border=2, data=src, fill_border="mirror")
    hd = orc.HostImage(*shape, "vuchar3")
    orc.load().vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
    assert np.array_equal(D.download(), hd.get())


def test_box5x5_extremes_and_u8(vpp):
    o = orc.load()
    for fillv in (0, 255):
        S = vpp.Image2d(40, 50, "vuchar3", border=2)
        vpp.fill_with_border(S, fillv)
        D = vpp.Image2d(40, 50, "vuchar3")
        vpp.box5x5(S, D)
        assert (D.download() == fillv).all()
    src = rng(12).integers(0, 256, (123, 457), dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "u8", border=3)
    vpp.fill_border_mirror(S)
    D = vpp.Image2d(123, 457, "u8", border=1)
    vpp.box5x5(S, D)
    hs = orc.HostImage(123, 457, "u8", border=3, data=src, fill_border="mirror")
    hd = orc.HostImage(123, 457, "u8")
    o.vo_box5x5_u8(hs.ptr(), hd.ptr(), 1)
    assert np.array_equal(D.download(), hd.get())


def test_box5x5_direct_path_on_views_matches(vpp):
    # a subimage is not TMA-describable -> direct kernel; must agree with the oracle on the same window
    src = rng(13).integers(0, 256, (90, 130, 3), dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "vuchar3", border=2)
    vpp.fill_border_mirror(S)
    D = vpp.Image2d(90, 130, "vuchar3")
    vpp.fill(D, 0)
    box = vpp.Box2d((10, 7), (70, 100))
    vpp.box5x5(S | box, D | box)
    hs = orc.HostImage(90, 130, "vuchar3", border=2, data=src, fill_border="mirror")
    hd = orc.HostImage(90, 130, "vuchar3")
    orc.load().vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
    exp = np.zeros_like(src)
    exp[10:71, 7:101] = hd.get()[10:71, 7:101]
    assert np.array_equal(D.download(), exp)


def test_box5x5_i32(vpp):
    # benchmarks/box_5x5_filter.cc:191 style values in [0,1000) + negative values (trunc toward 0)
    for lo, hi in ((0, 1000), (-1000, 1000)):
        src = rng(14).integers(lo, hi, (101, 203), dtype=np.int32)
        S = vpp.Image2d.from_host(src, "i32", border=2)
        vpp.fill_border_mirror(S)
        D = vpp.Image2d(101, 203, "i32")
        vpp.box5x5(S, D)
        hs = orc.HostImage(101, 203, "i32", border=2, data=src, fill_border="mirror")
        hd = orc.HostImage(101, 203, "i32")
        orc.load().vo_box5x5_i32(hs.ptr(), hd.ptr())
        assert np.array_equal(D.download(), hd.get())


def test_box_border_too_small_is_an_error(vpp):
    from vpp_b200 import capi

    S, D = vpp.Image2d(20, 20, "vuchar3", border=1), vpp.Image2d(20, 20, "vuchar3")
    with pytest.raises(capi.VppbError) as e:
        vpp.box5x5(S, D)
    assert e.value.code == capi.VPPB_E_BORDER


# ------------------------------------------------------------------ scharr + pyramid
@pytest.mark.parametrize("gpix", ["vint2", "vfloat2"])
@pytest.mark.parametrize("shape", [(100, 100), (33, 77), (541, 961)])
def test_scharr(vpp, gpix, shape):
    src = rng(21).integers(0, 256, shape, dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "u8", border=1)
    vpp.fill_border_mirror(S)
    G = vpp.Image2d(*shape, gpix)
    vpp.scharr(S, G)
    hs = orc.HostImage(*shape, "u8", border=1, data=src, fill_border="mirror")
    hg = orc.HostImage(*shape, gpix)
    orc.load().vo_scharr_u8(hs.ptr(), hg.ptr(), 1 if gpix == "vfloat2" else 0)
    got, exp = G.download(), hg.get()
    assert np.array_equal(got.view(np.int32), exp.view(np.int32))


@pytest.mark.parametrize("shape", [(100, 100), (101, 77), (270, 481), (1080, 1920)])
def test_pyramid_u8(vpp, shape):
    src = rng(22).integers(0, 256, shape, dtype=np.uint8)
    pyr = vpp.Pyramid2d(vpp.Image2d.from_host(src, "u8"), 3, 2, border=3)
    ref = oracle_pyramid(src, 3, "u8", 3)
    for l in range(3):
        assert (pyr[l].nrows, pyr[l].ncols) == (ref[l].nrows, ref[l].ncols)
        assert np.array_equal(pyr[l].download(with_border=True), ref[l].get(with_border=True)), "level %d" % l
    # pyramid.hh:140: 1080 -> 541 -> 271, 1920 -> 961 -> 481
    if shape == (1080, 1920):
        assert [(p.nrows, p.ncols) for p in pyr.levels] == [(1080, 1920), (541, 961), (271, 481)]


@pytest.mark.parametrize("gpix", ["vint2", "vfloat2"])
def test_gradient_pyramid(vpp, gpix):
    src = scenes.rectangles_scene(203, 301, seed=4)
    o = orc.load()
    prev = vpp.Pyramid2d(vpp.Image2d.from_host(src, "u8"), 3, 2, border=3)
    grad = vpp.Pyramid2d((203, 301), 3, 2, pixel=gpix, border=3)
    vpp.scharr(prev[0], grad[0])
    grad.propagate_level0()
    rprev = oracle_pyramid(src, 3, "u8", 3, o)
    rgrad = oracle_grad_pyramid(rprev, gpix, 3, o)
    for l in range(3):
        got, exp = grad[l].download(with_border=True), rgrad[l].get(with_border=True)
        assert np.array_equal(got.view(np.int32), exp.view(np.int32)), "level %d" % l


# ------------------------------------------------------------------ FAST9
def _oracle_fast(img, th, mask=None, mode=0, bs=10, ring=0, want_scores=False):
    o = orc.load()
    h = orc.HostImage(img.shape[0], img.shape[1], "u8", border=3, data=img, fill_border="mirror")
    hm = orc.HostImage(img.shape[0], img.shape[1], "u8", data=mask) if mask is not None else None
    cap = img.size
    kps = np.zeros((cap, 2), dtype=np.int32)
    sc = np.zeros(cap, dtype=np.int32)
    n = o.vo_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, bs, ring, kps.ctypes.data, sc.ctypes.data if want_scores else None, cap)
    assert n >= 0
    return kps[:n], sc[:n]


@pytest.mark.parametrize("th", [10, 20, 40])
@pytest.mark.parametrize("ring", ["reference", "true"])
def test_fast9_keypoints_bit_exact(vpp, th, ring):
    img = scenes.rectangles_scene(317, 403, seed=8)
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    sc = []
    kps = vpp.fast9(G, th, ring=ring, scores=sc)
    rk, rs = _oracle_fast(img, th, ring=0 if ring == "reference" else 1, want_scores=True)
    assert len(rk) > 50
    assert np.array_equal(kps, rk)
    assert np.array_equal(np.asarray(sc, dtype=np.int32), rs)


@pytest.mark.parametrize("maskval", [0xFF, 0x01, 0x10])
def test_fast9_mask_semantics(vpp, maskval):
    # fast.hpp:310-317: the mask byte seeds `possible`: 0x01 keeps only darker arcs (video_extruder.hpp:101)
    img = scenes.rectangles_scene(200, 260, seed=9)
    mask = np.zeros(img.shape, dtype=np.uint8)
    mask[20:150, 30:200] = maskval
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    M = vpp.Image2d.from_host(mask, "u8")
    kps = vpp.fast9(G, 15, mask=M)
    rk, _ = _oracle_fast(img, 15, mask=mask)
    assert len(rk) > 5 and np.array_equal(kps, rk)


@pytest.mark.parametrize("mode", ["local_maxima", "blockwise"])
def test_fast9_maxima_modes(vpp, mode):
    img = scenes.rectangles_scene(241, 322, seed=10)
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    sc = []
    kps = vpp.fast9(G, 12, local_maxima=mode == "local_maxima", blockwise=mode == "blockwise", block_size=10, scores=sc)
    rk, rs = _oracle_fast(img, 12, mode=1 if mode == "local_maxima" else 2, bs=10, want_scores=True)
    assert len(rk) > 5
    assert np.array_equal(kps, rk)
    assert np.array_equal(np.asarray(sc, dtype=np.int32), rs)


def test_fast9_edges_empty_and_errors(vpp):
    from vpp_b200 import capi

    flat = np.full((64, 100), 77, dtype=np.uint8)
    G = vpp.Image2d.from_host(flat, "u8", border=3)
    vpp.fill_border_mirror(G)
    assert len(vpp.fast9(G, 10)) == 0
    # corners at the image edge read the 3-px border
    img = np.zeros((40, 40), dtype=np.uint8)
    img[0:4, 0:4] = 255
    img[36:, 36:] = 255
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    rk, _ = _oracle_fast(img, 20)
    assert np.array_equal(vpp.fast9(G, 20), rk)
    with pytest.raises(capi.VppbError) as e:  # fast.hpp:937-938
        vpp.fast9(vpp.Image2d(10, 10, "u8", border=2), 10)
    assert e.value.code == capi.VPPB_E_BORDER
    pts = np.array([[5, 5], [20, 21]], dtype=np.int32)
    got = vpp.fast9_scores(G, 20, pts)
    h = orc.HostImage(40, 40, "u8", border=3, data=img, fill_border="mirror")
    assert list(got) == [orc.load().vo_fast9_score(h.ptr(), 20, 5, 5), orc.load().vo_fast9_score(h.ptr(), 20, 20, 21)]


# ------------------------------------------------------------------ Lucas-Kanade
def _relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


@pytest.mark.parametrize("nscales", [2, 3])
@pytest.mark.parametrize("winsize", [5, 7, 11])
def test_lucas_kanade_driver(vpp, winsize, nscales):
    f1, f2, pts = scenes.lk_pair(300, 400, 400, seed=31)
    flow, dist = vpp.lucas_kanade(vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), pts, winsize=winsize, nscales=nscales)
    rflow, rdist = oracle_lucas_kanade(f1, f2, pts, winsize=winsize, nscales=nscales)
    ok = rdist < 3e38
    assert ok.sum() > 300
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)  # identical failure flags
    assert (_relerr(flow, rflow) <= 1e-4).all()        # north_star tolerance
    assert np.allclose(dist[ok], rdist[ok], rtol=1e-4)
    if nscales == 2:
        # the synthetic flow is (2.3,-1.7) + 0.5 px sinusoid.  (With 3 levels the reference itself is
        # unstable: the gradient pyramid is the *blurred level-0 gradient* (lucas_kanade.hpp:156-157), i.e.
        # 2^S too small at level S, so the Gauss-Newton step overshoots 4x at level 2 - parity still holds.)
        good = np.abs(flow[ok] - np.array([2.3, -1.7])).max(axis=1) < 1.0
        assert good.mean() > 0.9


def test_lucas_kanade_prediction_and_failures(vpp):
    f1, f2, pts = scenes.lk_pair(200, 260, 100, seed=32)
    pred = np.tile(np.array([[2.0, -2.0]], dtype=np.float32), (len(pts), 1))
    # points hugging the border exercise the A.has() gating and the out-of-domain abort
    pts2 = np.concatenate([pts, np.array([[1, 1], [198, 258], [0, 130], [100, 0]], dtype=np.float32)])
    pred2 = np.concatenate([pred, np.zeros((4, 2), dtype=np.float32)])
    flow, dist = vpp.lucas_kanade(vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), pts2, winsize=7, nscales=2, prediction=pred2)
    rflow, rdist = oracle_lucas_kanade(f1, f2, pts2, winsize=7, nscales=2, prediction=pred2)
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)
    assert (_relerr(flow, rflow) <= 1e-4).all()


@pytest.mark.parametrize("gpix", ["vfloat2", "vint2"])
def test_pyrlk_match(vpp, gpix):
    f1, f2, pts = scenes.lk_pair(270, 480, 500, seed=33)
    o = orc.load()
    prev = vpp.Pyramid2d(vpp.Image2d.from_host(f1, "u8"), 3, 2, border=4)
    nxt = vpp.Pyramid2d(vpp.Image2d.from_host(f2, "u8"), 3, 2, border=4)
    grad = vpp.Pyramid2d((270, 480), 3, 2, pixel=gpix, border=4)
    vpp.scharr(prev[0], grad[0])
    grad.propagate_level0()
    flow, dist, keep = vpp.pyrlk_match(prev, grad, nxt, pts, 7, 0.01, 0.6, 21, 0.01)
    rprev, rnxt = oracle_pyramid(f1, 3, "u8", 4, o), oracle_pyramid(f2, 3, "u8", 4, o)
    rgrad = oracle_grad_pyramid(rprev, gpix, 4, o)
    P = orc.VoLkParams(nlevels=3, min_scale=0, winsize=7, max_iter=21, grad_is_float=1 if gpix == "vfloat2" else 0, err_mode=1,
                       gate_on_max_err=1, min_ev=0.01, delta=0.01, max_err=0.6, factor=2.0, pred_div=1.0)
    rflow, rdist = oracle_lk(rprev, rnxt, rgrad, P, pts)
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)
    assert (_relerr(flow, rflow) <= 1e-4).all()
    ok = rdist < 3e38
    assert np.allclose(dist[ok], rdist[ok], rtol=1e-4)
    assert keep.sum() > 100


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
def test_full_size_properties(vpp):
    # 4K add: linearity / checksum-of-checksums
    shape = (2160, 3840)
    r = rng(50)
    b = r.integers(0, 2 ** 30, shape, dtype=np.int32)
    c = r.integers(0, 2 ** 30, shape, dtype=np.int32)
    A, B, Cc = vpp.Image2d(*shape, "i32"), vpp.Image2d.from_host(b, "i32"), vpp.Image2d.from_host(c, "i32")
    vpp.pixel_wise_add(A, B, Cc)
    tot = (int(b.astype(np.int64).sum()) + int(c.astype(np.int64).sum())) & 0xFFFFFFFF
    assert vpp.sum(A) == (tot - (1 << 32) if tot >= (1 << 31) else tot)  # checksum of checksums (int wrap)
    assert np.array_equal(A.download(), b + c)
    # 4K box on vuchar3: constant image is a fixed point; box(x + k) == box(x) + k away from saturation
    src = r.integers(0, 200, shape + (3,), dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "vuchar3", border=2)
    vpp.fill_border_mirror(S)
    D1, D2 = vpp.Image2d(*shape, "vuchar3"), vpp.Image2d(*shape, "vuchar3")
    vpp.box5x5(S, D1)
    S.upload(src + 50)
    vpp.fill_border_mirror(S)
    vpp.box5x5(S, D2)
    assert np.array_equal(D2.download(), D1.download() + 50)
    # spot-check rows against the oracle on a horizontal band (oracle on the full frame takes seconds)
    hs = orc.HostImage(64, 3840, "vuchar3", border=2)
    band = (src + 50)[1000 - 2:1064 + 2]
    full = np.zeros((68, 3844, 3), dtype=np.uint8)
    full[:, 2:-2] = band
    full[:, :2] = band[:, 1::-1]
    full[:, -2:] = band[:, :-3:-1]
    hs.set(full, with_border=True)
    hd = orc.HostImage(64, 3840, "vuchar3")
    orc.load().vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
    assert np.array_equal(D2.download()[1000:1064], hd.get())


# ------------------------------------------------------------------ row-tile halos
def test_halo_pack_unpack_single_and_batch(vpp):
    from vpp_b200 import capi
    from vpp_b200.ops import _DeviceBuffer

    nr, nc, b, n = 12, 37, 2, 5
    r = rng(70)
    hosts = [r.integers(0, 256, (nr + 2 * b, nc + 2 * b, 3), dtype=np.uint8) for _ in range(n)]
    imgs = [vpp.Image2d(nr, nc, "vuchar3", border=b) for _ in range(n)]
    for im, h in zip(imgs, hosts):
        im.upload(h, with_border=True)
    hb = capi.lib.vppb_halo_bytes(imgs[0].ptr(), 2)
    assert hb == 2 * (nc + 2 * b) * 3
    descs = (capi.VppbImg * n)(*[im.desc for im in imgs])
    for which, rows in ((0, slice(b, b + 2)), (1, slice(b + nr - 2, b + nr))):
        st = _DeviceBuffer(n * hb)
        capi.check(capi.lib.vppb_halo_pack_batch(descs, n, 2, which, st.ptr, None))
        got = st.to_host(np.uint8, n * hb).reshape(n, 2, nc + 2 * b, 3)
        assert np.array_equal(got, np.stack([h[rows] for h in hosts]))
        st1 = _DeviceBuffer(hb)
        capi.check(capi.lib.vppb_halo_pack(imgs[3].ptr(), 2, which, st1.ptr, None))
        assert np.array_equal(st1.to_host(np.uint8, hb).reshape(2, nc + 2 * b, 3), hosts[3][rows])
    # unpack: staging -> border rows above (which 0) / below (which 1); the domain must stay untouched
    payload = r.integers(0, 256, (n, 2, nc + 2 * b, 3), dtype=np.uint8)
    st = _DeviceBuffer(payload.nbytes).from_host(payload)
    capi.check(capi.lib.vppb_halo_unpack_batch(descs, n, 2, 0, st.ptr, None))
    capi.check(capi.lib.vppb_halo_unpack_batch(descs, n, 2, 1, st.ptr, None))
    for i, im in enumerate(imgs):
        a = im.download(with_border=True)
        assert np.array_equal(a[0:2], payload[i]) and np.array_equal(a[-2:], payload[i])
        assert np.array_equal(a[2:-2], hosts[i][2:-2])


# ------------------------------------------------------------------ semi-dense optical flow (video_extruder)
@pytest.mark.parametrize("shape,ws,nscales,min_scale,prop,patch", [((121, 161), 9, 3, 0, 2, 5), ((145, 209), 7, 4, 0, 2, 5), ((129, 97), 9, 3, 1, 3, 3),
                                                                  ((240, 322), 9, 3, 0, 2, 5), ((121, 161), 9, 2, 0, 0, 5)])
def test_semi_dense_optical_flow_bit_exact(vpp, shape, ws, nscales, min_scale, prop, patch):
    f1, f2, _ = scenes.lk_pair(shape[0], shape[1], 4, seed=21, shift=(3.0, -2.0), margin=10)
    o = orc.load()
    G = vpp.Image2d.from_host(f1, "u8", border=3)
    vpp.fill_border_mirror(G)
    kps = vpp.fast9(G, 8, blockwise=True, block_size=6)  # what video_extruder feeds the flow with
    n = len(kps)
    assert n > 100
    pos, dist, valid = vpp.semi_dense_optical_flow(kps, vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), winsize=ws, nscales=nscales,
                                                   min_scale=min_scale, propagation=prop, patchsize=patch)
    h1, h2 = orc.HostImage(shape[0], shape[1], "u8", data=f1), orc.HostImage(shape[0], shape[1], "u8", data=f2)
    rpos, rdist, rvalid = np.zeros((n, 2), np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    k = np.ascontiguousarray(kps)
    o.vo_semi_dense_flow(h1.ptr(), h2.ptr(), k.ctypes.data, n, ws, nscales, min_scale, prop, patch, rpos.ctypes.data, rdist.ctypes.data, rvalid.ctypes.data)
    assert np.array_equal(valid, rvalid.astype(bool)) and valid.sum() > 50
    assert np.array_equal(pos, rpos)
    assert np.array_equal(dist, rdist)


# ------------------------------------------------------------------ video_extruder (semi-dense flow + FAST orchestration)
def test_video_extruder_gpu_equals_oracle(vpp):
    from vpp_b200 import video_extruder as ve
    from tests.oracle_video import OracleOps
    from tests.test_oracle_vs_ref import _moving_frames

    nr, nc, nf = 161, 241, 6
    frames = _moving_frames(nr, nc, nf)
    kw = dict(detector_th=6, keypoint_spacing=10, detector_period=3, max_trajectory_length=5, nscales=3, winsize=9, propagation=2)
    g, c = ve.video_extruder_init(nr, nc), ve.video_extruder_init(nr, nc)
    gops, cops = ve.GpuOps(), OracleOps()
    for f in range(1, nf):
        ve.video_extruder_update(g, frames[f - 1], frames[f], gops, **kw)
        ve.video_extruder_update(c, frames[f - 1], frames[f], cops, **kw)
        assert np.array_equal(ve.state_table(g), ve.state_table(c)), "frame %d" % f
    t = ve.state_table(g)
    assert len(t) > 20 and (t[:, 2] > 1).sum() > 5
