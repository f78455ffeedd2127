"""GPU parity AT THE SIZES BASELINE.json STATES (configs 2-5), not at reduced test sizes: every CUDA result is compared with
the CPU oracle (OpenMP build of the same restatement, pinned to the reference headers by tests/test_oracle_vs_ref.py) on
the full frame.  Integer / byte / index outputs bit-exact; LK displacements <= 1e-4 relative with identical failure flags
(north_star's tolerance; for 3 levels the oracle is the clamped-read definition, see tests/test_oracle_vs_ref.py:181-198)."""
import os

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


@pytest.fixture(scope="module")
def omp():
    o = orc.load(omp=True)
    return o


# ------------------------------------------------------------------ config 2 at 4K: full-frame box on vuchar3
@pytest.mark.parametrize("shape", [(2160, 3840), (1080, 1920), (4320, 7680)])
def test_box5x5_vuchar3_full_frame(vpp, omp, shape):
    r = np.random.default_rng(202)
    src = r.integers(0, 256, shape + (3,), dtype=np.uint8)
    S = vpp.Image2d.from_host(src, "vuchar3", border=2)
    vpp.fill_border_mirror(S)
    D = vpp.Image2d(*shape, "vuchar3")
    vpp.box5x5(S, D)
    hs = orc.HostImage(shape[0], shape[1], "vuchar3", border=2, data=src, fill_border="mirror")
    hd = orc.HostImage(shape[0], shape[1], "vuchar3")
    omp.vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
    assert np.array_equal(D.download(), hd.get())


# ------------------------------------------------------------------ config 3: FAST9 on 3840x2160, keypoint arrays bit-exact
def _oracle_fast(o, img, th, mask=None, mode=0, bs=10, ring=0):
    h = orc.HostImage(img.shape[0], img.shape[1], "u8", border=3, data=img, fill_border="mirror")
    hm = orc.HostImage(img.shape[0], img.shape[1], "u8", data=mask) if mask is not None else None
    cap = img.size // 4
    kps = np.zeros((cap, 2), dtype=np.int32)
    sc = np.zeros(cap, dtype=np.int32)
    n = o.vo_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, bs, ring, kps.ctypes.data, sc.ctypes.data, cap)
    assert 0 <= n <= cap
    return kps[:n], sc[:n]


@pytest.fixture(scope="module")
def scene4k():
    return scenes.rectangles_scene(2160, 3840, seed=42)


@pytest.mark.parametrize("maskval", [None, 0xFF, 0x01])
def test_fast9_4k_keypoints_bit_exact(vpp, scene4k, maskval):
    img = scene4k
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    mask, M = None, None
    if maskval is not None:
        mask = np.full(img.shape, maskval, dtype=np.uint8)
        mask[::7, ::5] = 0  # holes, so that the mask is really consulted per pixel
        M = vpp.Image2d.from_host(mask, "u8")
    sc = []
    kps = vpp.fast9(G, 20, mask=M, scores=sc)
    rk, rs = _oracle_fast(orc.load(), img, 20, mask=mask)  # serial oracle: raster order
    assert len(rk) > 20000
    assert np.array_equal(kps, rk)
    assert np.array_equal(np.asarray(sc, dtype=np.int32), rs)


@pytest.mark.parametrize("mode", ["local_maxima", "blockwise"])
def test_fast9_4k_maxima_modes(vpp, scene4k, mode):
    img = scene4k
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    sc = []
    kps = vpp.fast9(G, 20, local_maxima=mode == "local_maxima", blockwise=mode == "blockwise", block_size=10, scores=sc)
    rk, rs = _oracle_fast(orc.load(), img, 20, mode=1 if mode == "local_maxima" else 2, bs=10)
    assert len(rk) > 5000
    assert np.array_equal(kps, rk)
    assert np.array_equal(np.asarray(sc, dtype=np.int32), rs)


# ------------------------------------------------------------------ config 4: pyrLK 1080p, 3 levels, 10 000 keypoints, 7x7
def _relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


@pytest.fixture(scope="module")
def pair1080():
    return scenes.lk_pair(1080, 1920, 10000, seed=44)


def test_lucas_kanade_1080p_10k_vint2(vpp, omp, pair1080):
    f1, f2, pts = pair1080
    assert len(pts) == 10000
    flow, dist = vpp.lucas_kanade(vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), pts, winsize=7, nscales=3)
    rflow, rdist = oracle_lucas_kanade(f1, f2, pts, winsize=7, nscales=3, lib=omp)
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)  # identical failure flags
    ok = rdist < 3e38
    assert ok.sum() > 5000
    assert (_relerr(flow, rflow) <= 1e-4).all()
    assert np.allclose(dist[ok], rdist[ok], rtol=1e-4)


@pytest.mark.parametrize("gpix", ["vfloat2", "vint2"])
def test_pyrlk_match_1080p_10k(vpp, omp, pair1080, gpix):
    f1, f2, pts = pair1080
    prev = vpp.Pyramid2d(vpp.Image2d.from_host(f1, "u8"), 3, 2, border=4)
    nxt = vpp.Pyramid2d(vpp.Image2d.from_host(f2, "u8"), 3, 2, border=4)
    grad = vpp.Pyramid2d((1080, 1920), 3, 2, pixel=gpix, border=4)
    vpp.scharr(prev[0], grad[0])
    grad.propagate_level0()
    flow, dist, keep = vpp.pyrlk_match(prev, grad, nxt, pts, 7, 0.01, 0.6, 21, 0.01)
    rprev, rnxt = oracle_pyramid(f1, 3, "u8", 4, omp), oracle_pyramid(f2, 3, "u8", 4, omp)
    rgrad = oracle_grad_pyramid(rprev, gpix, 4, omp)
    # pyramids and gradients are integer / float-exact images: compare them whole (incl. borders) before the matcher
    for lvl in range(3):
        assert np.array_equal(prev[lvl].download(with_border=True), rprev[lvl].get(with_border=True)), "prev level %d" % lvl
        g, rg = grad[lvl].download(with_border=True), rgrad[lvl].get(with_border=True)
        assert np.array_equal(g.view(np.uint32), rg.view(np.uint32)), "grad level %d" % lvl
    P = orc.VoLkParams(nlevels=3, min_scale=0, winsize=7, max_iter=21, grad_is_float=1 if gpix == "vfloat2" else 0, err_mode=1,
                       gate_on_max_err=1, min_ev=0.01, delta=0.01, max_err=0.6, factor=2.0, pred_div=1.0)
    rflow, rdist = oracle_lk(rprev, rnxt, rgrad, P, pts, lib=omp)
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)
    assert (_relerr(flow, rflow) <= 1e-4).all()
    ok = rdist < 3e38
    assert np.allclose(dist[ok], rdist[ok], rtol=1e-4)
    assert keep.sum() > 3000


# ------------------------------------------------------------------ config 5's kernel: semi-dense flow with video_extruder's settings
def _sdof_case(vpp, shape, seed):
    f1, f2, _ = scenes.lk_pair(shape[0], shape[1], 4, seed=seed, shift=(3.0, -2.0), margin=10)
    G = vpp.Image2d.from_host(f1, "u8", border=3)
    vpp.fill_border_mirror(G)
    kps = vpp.fast9(G, 10, blockwise=True, block_size=10)  # video_extruder.hpp:111: blockwise FAST, one keypoint per 10x10 block
    return f1, f2, kps


@pytest.mark.parametrize("shape", [(1080, 1920), (2160, 3840)])
def test_semi_dense_flow_full_frame(vpp, shape):
    f1, f2, kps = _sdof_case(vpp, shape, 55)
    n = len(kps)
    assert n > 2000
    pos, dist, valid = vpp.semi_dense_optical_flow(kps, vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), winsize=9, nscales=3,
                                                   min_scale=0, propagation=2, patchsize=5)
    o = orc.load()
    h1, h2 = orc.HostImage(shape[0], shape[1], "u8", data=f1), orc.HostImage(shape[0], shape[1], "u8", data=f2)
    rpos, rdist, rvalid = np.zeros((n, 2), np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    k = np.ascontiguousarray(kps)
    o.vo_semi_dense_flow(h1.ptr(), h2.ptr(), k.ctypes.data, n, 9, 3, 0, 2, 5, rpos.ctypes.data, rdist.ctypes.data, rvalid.ctypes.data)
    assert np.array_equal(valid, rvalid.astype(bool)) and valid.sum() > 1000
    assert np.array_equal(pos, rpos)
    assert np.array_equal(dist, rdist)


# ------------------------------------------------------------------ SURVEY 8(f) N4 at frame sizes (the inputs bench.py's extras time and check)
def test_lbp_transform_4k(vpp, omp):
    f = np.random.default_rng(4).integers(0, 256, (2160, 3840), dtype=np.uint8)
    A = vpp.Image2d.from_host(f, "u8", border=1)
    vpp.fill_border_mirror(A)
    B = vpp.lbp_transform(A)
    hs = orc.HostImage(2160, 3840, "u8", border=1, data=f, fill_border="mirror")
    hd = orc.HostImage(2160, 3840, "u8")
    omp.vo_lbp_u8(hs.ptr(), hd.ptr())
    assert np.array_equal(B.download(), hd.get())


def test_local_maxima_filter_1080p(vpp):
    """a FAST-like sparse score image with 200-pixel ramps (chains of dependent decisions): the serial raster-order result, many CTAs"""
    r_ = np.random.default_rng(6)
    sc_ = np.where(r_.random((1080, 1920)) < 0.03, r_.integers(1, 250, (1080, 1920)), 0).astype(np.uint8)
    sc_[100:140, 200:900] = (250 - (np.arange(700) % 200))[None, :].astype(np.uint8)
    A = vpp.Image2d.from_host(sc_, "u8", border=1)
    vpp.fill_border_with_value(A, 0)
    vpp.local_maxima_filter(A)
    hs = orc.HostImage(1080, 1920, "u8", border=1, data=sc_)
    orc.load().vo_local_maxima_filter(hs.ptr())
    assert np.array_equal(A.download(), hs.get())
