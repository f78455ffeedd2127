"""Pin the oracle restatement to the reference's OWN code: oracle/_ref/libvppref*.so is the reference's
headers (/root/reference/vpp, verbatim) compiled against the Eigen / iod stand-ins in oracle/ref_shim.
Runs wherever oracle/_ref has been built (this container; the .so also travels to the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import oracle as orc
from tests import scenes
from tests.oracle_ops import oracle_grad_pyramid, oracle_lk, oracle_lucas_kanade, oracle_pyramid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libvppref.so")
REF_OMP = os.path.join(ROOT, "oracle", "_ref", "libvppref_omp.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")

I = C.POINTER(orc.VoImg)


def _load(path):
    r = C.CDLL(path)
    r.vppref_pw_add_i32.argtypes = [I, I, I]
    r.vppref_fill_border_mirror.argtypes = [I]
    r.vppref_fill_border_closest.argtypes = [I]
    r.vppref_box5x5_i32.argtypes = [I, I]
    r.vppref_box5x5_u8c3.argtypes = [I, I]
    r.vppref_scharr_u8.argtypes = [I, I, C.c_int]
    r.vppref_rgb_to_graylevel.argtypes = [I, I]
    r.vppref_rgb_to_graylevel_v1.argtypes = [I, I]
    r.vppref_lowpass_u8.argtypes = [I, I]
    r.vppref_pyramid.argtypes = [I, C.c_int, I, C.c_int]
    r.vppref_fast9_u8.argtypes = [I, C.c_int, I, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    r.vppref_fast9_blockwise_native_order.argtypes = [I, C.c_int, I, C.c_int, C.c_void_p, C.c_int]
    r.vppref_fast9_score.argtypes = [I, C.c_int, C.c_int, C.c_int]
    r.vppref_is_fast9_keypoint.argtypes = [I, C.c_int, C.c_int, C.c_int]
    r.vppref_interp_u8.argtypes = [I, C.c_float, C.c_float]
    r.vppref_lucas_kanade.argtypes = [I, I, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    r.vppref_video_extruder.argtypes = [I, C.c_int] + [C.c_int] * 7 + [C.c_void_p, C.c_int]
    r.vppref_semi_dense_flow.argtypes = [I, I, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    r.vppref_pyrlk_levels.argtypes = [I, I, I, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p]
    return r


@pytest.fixture(scope="module")
def ref(built):
    return _load(REF)


@pytest.fixture(scope="module")
def o(built):
    return orc.load()


def rng(s):
    return np.random.default_rng(s)


def test_add_borders_box(ref, o):
    b, c = rng(1).integers(-2 ** 30, 2 ** 30, (2, 37, 53), dtype=np.int32)
    a1, a2 = orc.HostImage(37, 53, "i32", aligned=16), orc.HostImage(37, 53, "i32", aligned=16)
    hb, hc = orc.HostImage(37, 53, "i32", aligned=16, data=b), orc.HostImage(37, 53, "i32", aligned=16, data=c)
    ref.vppref_pw_add_i32(a1.ptr(), hb.ptr(), hc.ptr())
    o.vo_pw_add_i32(a2.ptr(), hb.ptr(), hc.ptr())
    assert np.array_equal(a1.get(), a2.get())
    for pix in ("u8", "vuchar3", "i32", "vint2"):
        dt, ch = orc.PIXEL_TYPES[pix]
        d = rng(2).integers(0, 200, (9, 14) + ((ch,) if ch > 1 else ())).astype(dt)
        h1, h2 = orc.HostImage(9, 14, pix, border=3, data=d), orc.HostImage(9, 14, pix, border=3, data=d)
        ref.vppref_fill_border_mirror(h1.ptr()); o.vo_fill_border_mirror(h2.ptr())
        assert np.array_equal(h1.get(True), h2.get(True)), pix
        if pix != "vint2":
            ref.vppref_fill_border_closest(h1.ptr()); o.vo_fill_border_closest(h2.ptr())
            assert np.array_equal(h1.get(True), h2.get(True)), pix
    s = rng(3).integers(-1000, 1000, (41, 67), dtype=np.int32)
    hs = orc.HostImage(41, 67, "i32", border=2, data=s, fill_border="mirror")
    d1, d2 = orc.HostImage(41, 67, "i32"), orc.HostImage(41, 67, "i32")
    ref.vppref_box5x5_i32(hs.ptr(), d1.ptr()); o.vo_box5x5_i32(hs.ptr(), d2.ptr())
    assert np.array_equal(d1.get(), d2.get())
    u = rng(4).integers(0, 256, (45, 71, 3), dtype=np.uint8)
    hs = orc.HostImage(45, 71, "vuchar3", border=2, data=u, fill_border="mirror")
    d1, d2 = orc.HostImage(45, 71, "vuchar3"), orc.HostImage(45, 71, "vuchar3")
    ref.vppref_box5x5_u8c3(hs.ptr(), d1.ptr()); o.vo_box5x5_u8(hs.ptr(), d2.ptr(), 3)
    assert np.array_equal(d1.get(), d2.get())


def test_interp_scharr_lowpass(ref, o):
    a = rng(5).integers(0, 256, (33, 47), dtype=np.uint8)
    h = orc.HostImage(33, 47, "u8", border=2, data=a, fill_border="mirror")
    for (pr, pc) in rng(6).uniform(0, 30, (200, 2)).astype(np.float32):
        assert ref.vppref_interp_u8(h.ptr(), pr, pc) == o.vo_interp_u8(h.ptr(), pr, pc)
    for gpix in ("vint2", "vfloat2"):
        g1, g2 = orc.HostImage(33, 47, gpix), orc.HostImage(33, 47, gpix)
        ref.vppref_scharr_u8(h.ptr(), g1.ptr(), 1 if gpix == "vfloat2" else 0)
        o.vo_scharr_u8(h.ptr(), g2.ptr(), 1 if gpix == "vfloat2" else 0)
        assert np.array_equal(g1.get().view(np.int32), g2.get().view(np.int32)), gpix
    l1, l2 = orc.HostImage(33, 47, "u8"), orc.HostImage(33, 47, "u8")
    ref.vppref_lowpass_u8(h.ptr(), l1.ptr()); o.vo_lowpass(h.ptr(), l2.ptr(), 0)
    assert np.array_equal(l1.get(), l2.get())


@pytest.mark.parametrize("kind,pix", [(0, "u8"), (1, "vint2"), (2, "vfloat2")])
@pytest.mark.parametrize("shape", [(101, 77), (100, 80)])
def test_pyramids(ref, o, kind, pix, shape):
    a = scenes.rectangles_scene(shape[0], shape[1], seed=7)
    src = orc.HostImage(shape[0], shape[1], "u8", data=a)
    mine = oracle_pyramid(a, 3, "u8", 3, o)
    if kind:
        mine = oracle_grad_pyramid(mine, pix, 3, o)
    theirs = [orc.HostImage(l.nrows, l.ncols, pix, border=3) for l in mine]
    ref.vppref_pyramid(src.ptr(), 3, orc.desc_array(theirs), kind)
    for lvl in range(3):
        x, y = theirs[lvl].get(True), mine[lvl].get(True)
        if pix != "u8":
            x, y = x.view(np.int32), y.view(np.int32)
        if shape[0] % 2 == 1 or lvl == 0:
            # odd parent sizes never read the reference's uninitialised low-pass border (pyramid.hh:179-181)
            assert np.array_equal(x, y), "level %d" % lvl
        else:
            # even parent size: the last row/col of the level (and what mirrors / blurs it) is garbage in the
            # reference; everything that does not depend on it must still match
            m = 3 + 4 * lvl
            assert np.array_equal(x[3:-m - 3, 3:-m - 3], y[3:-m - 3, 3:-m - 3]), "level %d interior" % lvl


@pytest.mark.parametrize("lib_path", [REF, REF_OMP], ids=["scalar-tree", "avx2-tree"])
def test_fast9_reference_ring(o, built, lib_path):
    """The pruning tree of fast.hpp:253-508 (scalar fallback and the AVX2 build) == the oracle's 9-arc
    test on the ring as implemented; mask, threshold, maxima modes and scores included."""
    if not os.path.exists(lib_path):
        pytest.skip("not built")
    r = _load(lib_path)
    cases = []
    for seed, th in [(8, 10), (9, 20), (10, 40)]:
        cases.append((scenes.rectangles_scene(131, 160, seed=seed), th))
    # dense random patterns: pixels drawn from 3 levels exercise a large share of the 2^16 ring patterns
    cases.append((rng(11).choice(np.array([20, 100, 180], np.uint8), (96, 128)), 30))
    cases.append((rng(12).integers(0, 256, (64, 96), dtype=np.uint8), 15))
    for img, th in cases:
        nr, nc = img.shape
        h = orc.HostImage(nr, nc, "u8", border=3, aligned=32, data=img, fill_border="mirror")
        for mode in (0, 1, 2):
            for maskval in (None, 0xFF, 0x01, 0x10):
                hm = None
                if maskval is not None:
                    m = np.zeros(img.shape, np.uint8)
                    m[5:nr - 9, 7:nc - 11] = maskval
                    hm = orc.HostImage(nr, nc, "u8", aligned=32, data=m)
                k1, k2 = np.zeros((img.size, 2), np.int32), np.zeros((img.size, 2), np.int32)
                s1, s2 = np.zeros(img.size, np.int32), np.zeros(img.size, np.int32)
                n1 = r.vppref_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, 10, k1.ctypes.data, s1.ctypes.data, img.size)
                n2 = o.vo_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, mode, 10, 0, k2.ctypes.data, s2.ctypes.data, img.size)
                assert n1 == n2, (th, mode, maskval, n1, n2)
                if mode == 2:  # the oracle emits blockwise keypoints in cell order (the reference's serial order); the wrapper sorts by pixel
                    order = np.lexsort((k2[:n2, 1], k2[:n2, 0]))
                    k2[:n2], s2[:n2] = k2[:n2][order], s2[:n2][order]
                assert np.array_equal(k1[:n1], k2[:n2]) and np.array_equal(s1[:n1], s2[:n2]), (th, mode, maskval)
                if mode == 2 and lib_path == REF:  # serial build: the reference's own output order == the oracle's
                    k3, k4 = np.zeros((img.size, 2), np.int32), np.zeros((img.size, 2), np.int32)
                    n3 = r.vppref_fast9_blockwise_native_order(h.ptr(), th, hm.ptr() if hm else None, 10, k3.ctypes.data, img.size)
                    n4 = o.vo_fast9_u8(h.ptr(), th, hm.ptr() if hm else None, 2, 10, 0, k4.ctypes.data, None, img.size)
                    assert n3 == n4 and np.array_equal(k3[:n3], k4[:n4])
        assert n1 >= 0


def test_fast9_true_ring_and_score(ref, o):
    img = scenes.rectangles_scene(90, 120, seed=13)
    h = orc.HostImage(90, 120, "u8", border=3, data=img, fill_border="mirror")
    k = np.zeros((img.size, 2), np.int32)
    n = o.vo_fast9_u8(h.ptr(), 20, None, 0, 10, 1, k.ctypes.data, None, img.size)
    mine = set(map(tuple, k[:n]))
    theirs = {(r_, c_) for r_ in range(90) for c_ in range(120) if ref.vppref_is_fast9_keypoint(h.ptr(), 20, r_, c_)}
    assert len(mine) > 20 and mine == theirs
    for (r_, c_) in list(mine)[:50]:
        assert ref.vppref_fast9_score(h.ptr(), 20, r_, c_) == o.vo_fast9_score(h.ptr(), 20, r_, c_)


@pytest.mark.parametrize("winsize,nscales", [(5, 2), (7, 2), (11, 2), (7, 1), (7, 3), (11, 3)])
def test_lucas_kanade_bit_exact(ref, o, winsize, nscales):
    """lucas_kanade() of the reference vs the oracle: bit-identical flows and distances.  With 3 levels the
    reference's 4x overshoot at level 2 (blurred level-0 gradient, lucas_kanade.hpp:156-157) throws a few
    tracks against the image edge where its un-checked bilinear taps read outside the allocated border
    (undefined values; the oracle clamps): those (< 2 % of the points) are allowed to differ."""
    f1, f2, pts = scenes.lk_pair(301, 401, 400, seed=14, margin=40)
    h1, h2 = orc.HostImage(301, 401, "u8", data=f1), orc.HostImage(301, 401, "u8", data=f2)
    n = len(pts)
    for pred in (None, np.tile(np.array([[2.0, -2.0]], np.float32), (n, 1))):
        flow, dist = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
        ref.vppref_lucas_kanade(h1.ptr(), h2.ptr(), pts.ctypes.data, pred.ctypes.data if pred is not None else None, n, 21, winsize, nscales,
                                0.0001, 0.1, flow.ctypes.data, dist.ctypes.data)
        rflow, rdist = oracle_lucas_kanade(f1, f2, pts, niterations=21, winsize=winsize, nscales=nscales, prediction=pred, lib=o)
        bad = (flow.view(np.int32) != rflow.view(np.int32)).any(axis=1) | (dist.view(np.int32) != rdist.view(np.int32))
        if nscales <= 2:
            assert not bad.any(), (bad.sum(), np.abs(flow - rflow).max())
        else:
            assert bad.mean() < 0.02, bad.sum()


def test_reference_pyrlk_kat_through_real_headers(ref):
    # tests/pyrlk.cc:14-50 executed by the reference's own lucas_kanade()
    d = np.load(os.path.join(ROOT, "tests", "golden", "pyrlk_scene.npz"))
    h1, h2 = orc.HostImage(100, 100, "u8", data=d["i1"]), orc.HostImage(100, 100, "u8", data=d["i2"])
    kp = np.array([[50, 50]], np.float32)
    flow, dist = np.zeros((1, 2), np.float32), np.zeros(1, np.float32)
    ref.vppref_lucas_kanade(h1.ptr(), h2.ptr(), kp.ctypes.data, None, 1, 50, 5, 2, 0.001, 0.01, flow.ctypes.data, dist.ctypes.data)
    assert np.linalg.norm(flow[0] - np.array([2.0, 2.0])) < 0.05, flow


@pytest.mark.parametrize("winsize", [5, 7])
def test_lk_square_win_matcher(ref, o, winsize):
    """lk_match_point_square_win<WS> (lk.hh:42-175) inside the pyrlk_match loop, float gradient pyramid."""
    f1, f2, pts = scenes.lk_pair(141, 181, 120, seed=15, margin=30)
    prev, nxt = oracle_pyramid(f1, 2, "u8", 4, o), oracle_pyramid(f2, 2, "u8", 4, o)
    grad = oracle_grad_pyramid(prev, "vfloat2", 4, o)
    n = len(pts)
    flow, dist = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    ref.vppref_pyrlk_levels(orc.desc_array(prev), orc.desc_array(nxt), orc.desc_array(grad), 2, 0, winsize, pts.ctypes.data, n, 0.01, 0.6, 21.0, 0.01,
                            flow.ctypes.data, dist.ctypes.data)
    P = orc.VoLkParams(nlevels=2, min_scale=0, winsize=winsize, max_iter=21, grad_is_float=1, err_mode=1, gate_on_max_err=1, min_ev=0.01, delta=0.01,
                       max_err=0.6, factor=2.0, pred_div=1.0)
    rflow, rdist = oracle_lk(prev, nxt, grad, P, pts, lib=o)
    assert np.array_equal(dist >= 3e38, rdist >= 3e38)
    assert np.allclose(flow, rflow, rtol=1e-5, atol=1e-5), np.abs(flow - rflow).max()
    ok = rdist < 3e38
    assert np.allclose(dist[ok], rdist[ok], rtol=1e-4)


@pytest.mark.parametrize("shape,ws,nscales,min_scale,prop,patch", [((121, 161), 9, 3, 0, 2, 5), ((145, 209), 7, 4, 0, 2, 5),
                                                                  ((129, 97), 9, 3, 1, 3, 3), ((121, 161), 9, 2, 0, 0, 5)])
def test_semi_dense_flow_serial_semantics(ref, o, shape, ws, nscales, min_scale, prop, patch):
    """semi_dense_optical_flow (semi_dense_optical_flow.hpp:46-214 + gradient_descent.hh) executed by the reference's
    own headers, serial build, vs the oracle restatement: positions, distances and the set of reported keypoints.
    Sizes of the form 2^k m + 1 keep every pyramid level odd, so the reference never reads its uninitialised
    low-pass border (pyramid.hh:179-181)."""
    f1, f2, _ = scenes.lk_pair(shape[0], shape[1], 4, seed=21, shift=(3.0, -2.0), margin=10)
    k = np.zeros((f1.size, 2), np.int32)
    h = orc.HostImage(shape[0], shape[1], "u8", border=3, data=f1, fill_border="mirror")
    n = o.vo_fast9_u8(h.ptr(), 8, None, 2, 6, 0, k.ctypes.data, None, len(k))  # blockwise FAST keypoints, as video_extruder feeds it
    kps = np.ascontiguousarray(k[:n])
    assert n > 100
    h1, h2 = orc.HostImage(shape[0], shape[1], "u8", data=f1), orc.HostImage(shape[0], shape[1], "u8", data=f2)
    res = []
    for fn in (ref.vppref_semi_dense_flow, o.vo_semi_dense_flow):
        pos, dist, valid = np.zeros((n, 2), np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        fn(h1.ptr(), h2.ptr(), kps.ctypes.data, n, ws, nscales, min_scale, prop, patch, pos.ctypes.data, dist.ctypes.data, valid.ctypes.data)
        res.append((pos, dist, valid))
    assert np.array_equal(res[0][2], res[1][2]) and res[0][2].sum() > 50
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    flow = (res[1][0] - kps)[res[1][2] > 0]
    assert np.median(np.abs(flow - np.array([3, -2])).max(axis=1)) <= 1  # the synthetic motion is (3,-2) +- 0.5 px


def _moving_frames(nr, nc, nframes, seed=31):
    """a textured scene translating by (2,-1) px per frame (+ a little per-frame noise)"""
    base = scenes.rectangles_scene(nr + 64, nc + 64, seed=seed, noise=2)
    rngf = np.random.default_rng(seed)
    out = []
    for f in range(nframes):
        a = base[32 - 2 * f:32 - 2 * f + nr, 32 + f:32 + f + nc].astype(np.int32) + rngf.integers(-1, 2, (nr, nc))
        out.append(np.clip(a, 0, 255).astype(np.uint8))
    return out


def test_video_extruder_orchestration(ref, o):
    """video_extruder_update (video_extruder.hpp:24-135) run by the reference's own headers over 7 frames vs the
    Python orchestration of vpp_b200.video_extruder on the oracle backend: identical keypoints, ages and trajectories."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("video_extruder", os.path.join(ROOT, "vpp_b200", "video_extruder.py"))
    ve = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ve)
    from tests.oracle_video import OracleOps

    # the -DNDEBUG build (the flags of the reference's examples / benchmarks), single thread = serial semantics: with
    # asserts on, keypoint_container.hpp:82 aborts as soon as a dead keypoint is revived by move() - which the
    # reference's own update loop does (video_extruder.hpp:48-51 on entries that died but were not compacted yet)
    ref = _load(REF_OMP)
    ref.vppref_set_num_threads(1)
    nr, nc, nf = 161, 241, 7
    frames = _moving_frames(nr, nc, nf)
    hosts = [orc.HostImage(nr, nc, "u8", border=10, aligned=32, data=f, fill_border="mirror") for f in frames]
    out = np.zeros((nr * nc, 6), np.int32)
    n = ref.vppref_video_extruder(orc.desc_array(hosts), nf, 6, 10, 3, 5, 3, 9, 2, out.ctypes.data, len(out))
    assert n > 20
    ctx = ve.video_extruder_init(nr, nc)
    ops = OracleOps(o)
    for f in range(1, nf):
        ve.video_extruder_update(ctx, frames[f - 1], frames[f], ops, detector_th=6, keypoint_spacing=10, detector_period=3,
                                 max_trajectory_length=5, nscales=3, winsize=9, propagation=2)
    mine = ve.state_table(ctx)
    assert len(mine) == n
    assert np.array_equal(mine, out[:n])
    assert (mine[:, 2] > 1).sum() > 5  # some keypoints really were tracked across frames


@pytest.mark.parametrize("pix", ["vuchar3", "vuchar4"])
def test_rgb_to_graylevel(ref, o, pix):
    """rgb_to_graylevel<unsigned char> of the reference (colorspace_conversions.hh:22-47) vs the oracle: domain and border,
    every channel sum 0..765 occurs; + the reference's own KAT (tests/colorspace_conversions.cc:8-23): gray(i,i,i) == i."""
    ch = 3 if pix == "vuchar3" else 4
    data = rng(77).integers(0, 256, (45, 67, ch), dtype=np.uint8)
    data[0, :, :3] = 255
    data[1, :, :3] = 0
    for b in (0, 3):
        src = orc.HostImage(45, 67, pix, border=b, aligned=32, data=data, fill_border="mirror" if b else None)
        g1, g2 = orc.HostImage(45, 67, "u8", border=b, aligned=32), orc.HostImage(45, 67, "u8", border=b, aligned=32)
        ref.vppref_rgb_to_graylevel(src.ptr(), g1.ptr())
        o.vo_rgb_to_graylevel(src.ptr(), g2.ptr())
        assert np.array_equal(g1.get(True), g2.get(True))
        assert np.array_equal(g2.get(), (data[..., :3].astype(np.int32).sum(axis=2) // 3).astype(np.uint8))
    if pix == "vuchar3":
        kat = (np.arange(100 * 100) % 256).astype(np.uint8).reshape(100, 100)
        src = orc.HostImage(100, 100, "vuchar3", data=np.repeat(kat[..., None], 3, axis=2))
        for fn in (ref.vppref_rgb_to_graylevel_v1, o.vo_rgb_to_graylevel):
            g = orc.HostImage(100, 100, "u8")
            fn(src.ptr(), g.ptr())
            assert np.array_equal(g.get(), kat)
