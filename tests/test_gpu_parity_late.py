"""GPU parity tests of what was built AFTER the round's last GPU run (verified on the CPU emulator only, see
profiles/r1_emulator_verification.md): they live in their own module, which sorts after tests/test_gpu_parity.py, so that
`pytest -m gpu -x` reaches them only once everything that already ran on hardware has been re-checked.
tests/test_emulated_parity.py collects them for the emulator like the others."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import oracle as orc
from tests.test_gpu_parity import rng, vpp  # noqa: F401  (the module-scoped CUDA fixture)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_linear_copy_path_of_upload_download(vpp):
    # gap-free images (row bytes a multiple of the alignment, no border) take the single linear copy of vppb_upload / download
    for (nr, nc, pix) in [(33, 128, "vuchar3"), (17, 64, "i32"), (9, 256, "u8")]:
        img = vpp.Image2d(nr, nc, pix)
        assert img.pitch == nc * img.elem_bytes
        a = rng(3).integers(0, 255, img._host_shape(False)).astype(img.dtype)
        img.upload(a)
        assert np.array_equal(img.download(), a)
        sub = img | vpp.Box2d((2, 8), (5, 40))  # a view of it is pitched again
        assert np.array_equal(sub.download(), a[2:6, 8:41])


@pytest.mark.parametrize("pix,shape,n", [("vuchar3", (64, 341), 3), ("vuchar3", (270, 480), 5), ("vuchar3", (40, 330), 33), ("u8", (57, 1000), 4),
                                         ("vuchar3", (1080, 1920), 2), ("vuchar3", (7, 9), 2)])
def test_box5x5_batch_equals_oracle(vpp, pix, shape, n):
    """vppb_box5x5_*_batch: one persistent launch over the tiles of all images (two launches for 33 images), every image
    with its own tensor map / output base; each result must equal the oracle's"""
    o = orc.load()
    ch = 3 if pix == "vuchar3" else 1
    srcs, dsts, exp = [], [], []
    for i in range(n):
        data = rng(1000 * i + shape[0]).integers(0, 256, shape + ((ch,) if ch > 1 else ()), dtype=np.uint8)
        S = vpp.Image2d.from_host(data, pix, border=2)
        vpp.fill_border_mirror(S)
        srcs.append(S)
        dsts.append(vpp.Image2d(shape[0], shape[1], pix))
        hs = orc.HostImage(shape[0], shape[1], pix, border=2, data=data, fill_border="mirror")
        hd = orc.HostImage(shape[0], shape[1], pix)
        o.vo_box5x5_u8(hs.ptr(), hd.ptr(), ch)
        exp.append(hd.get())
    vpp.box5x5_batch(srcs, dsts)
    for i in range(n):
        assert np.array_equal(dsts[i].download(), exp[i]), i


def test_box5x5_batch_fallbacks_and_errors(vpp):
    """mixed shapes, views and small alignments take the image-by-image route with the same results; errors as the single entry"""
    from vpp_b200 import capi

    o = orc.load()
    shapes = [(40, 50), (64, 341), (40, 50)]
    srcs, dsts, exp = [], [], []
    for i, sh in enumerate(shapes):
        data = rng(70 + i).integers(0, 256, sh + (3,), dtype=np.uint8)
        S = vpp.Image2d.from_host(data, "vuchar3", border=2, aligned=128 if i else 4)
        vpp.fill_border_mirror(S)
        srcs.append(S)
        dsts.append(vpp.Image2d(sh[0], sh[1], "vuchar3"))
        hs, hd = orc.HostImage(sh[0], sh[1], "vuchar3", border=2, data=data, fill_border="mirror"), orc.HostImage(sh[0], sh[1], "vuchar3")
        o.vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
        exp.append(hd.get())
    vpp.box5x5_batch(srcs, dsts)
    for i in range(len(shapes)):
        assert np.array_equal(dsts[i].download(), exp[i]), i
    bad = vpp.Image2d(40, 50, "vuchar3", border=1)
    with pytest.raises(capi.VppbError) as e:
        vpp.box5x5_batch([srcs[0], bad], [dsts[0], vpp.Image2d(40, 50, "vuchar3")])
    assert e.value.code == capi.VPPB_E_BORDER
    assert capi.lib.vppb_box5x5_u8c3_batch(None, None, 0, None) == capi.VPPB_E_ARG


@pytest.mark.parametrize("pix,kind", [("u8", 0), ("vint2", 1), ("vfloat2", 2)])
def test_fused_level_equals_lowpass_then_mirror(vpp, pix, kind):
    """vppb_lowpass_sub2_mirror == vppb_lowpass_sub2 + vppb_fill_border_mirror on every geometry: even / odd parents
    (tail work items of the u8 launch), tiny levels where one pixel mirrors into both borders, prefixes of the level,
    unaligned parents (generic kernel), borders 0..5."""
    from vpp_b200 import capi

    dt, ch = orc.PIXEL_TYPES[pix]
    cases = [(4, 4, 2, 128), (5, 9, 3, 128), (8, 8, 4, 128), (9, 6, 2, 128), (16, 40, 5, 128), (33, 70, 3, 128), (64, 65, 0, 128), (97, 130, 4, 128),
             (270, 481, 3, 128), (21, 30, 3, 1), (40, 41, 2, 4), (12, 200, 5, 128), (200, 12, 5, 128)]
    for nr, nc, b, al in cases:
        shape = (nr, nc) + ((ch,) if ch > 1 else ())
        data = rng(nr * 1000 + nc).integers(0, 256, shape).astype(dt)
        parent = vpp.Image2d.from_host(data, pix, border=2, aligned=al)
        vpp.fill_border_mirror(parent)
        for onr, onc in ((1 + nr // 2, 1 + nc // 2), (max(nr // 2, b, 1), max(nc // 2 - 1, b, 1))):
            if b > onr or b > onc:
                continue
            a = vpp.Image2d(onr, onc, pix, border=b, aligned=al)
            f = vpp.Image2d(onr, onc, pix, border=b, aligned=al)
            for im in (a, f):
                vpp.fill(im, 77 if ch == 1 else [77] * ch)
                vpp.fill_border_with_value(im, 33 if ch == 1 else [33] * ch)
            capi.check(capi.lib.vppb_lowpass_sub2(parent.ptr(), a.ptr(), kind, None))
            vpp.fill_border_mirror(a)
            capi.check(capi.lib.vppb_lowpass_sub2_mirror(parent.ptr(), f.ptr(), kind, None))
            x, y = a.download(with_border=True), f.download(with_border=True)
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (nr, nc, b, al, onr, onc)
    # a border wider than the level cannot be mirrored
    small = vpp.Image2d(3, 3, pix, border=4)
    big = vpp.Image2d.from_host(rng(1).integers(0, 256, (5, 5) + ((ch,) if ch > 1 else ())).astype(dt), pix, border=2)
    assert capi.lib.vppb_lowpass_sub2_mirror(big.ptr(), small.ptr(), kind, None) == capi.VPPB_E_BORDER


@pytest.mark.parametrize("pix", ["vuchar3", "vuchar4"])
@pytest.mark.parametrize("shape", [(45, 67), (270, 480), (1, 1), (33, 16), (64, 1000)])
def test_rgb_to_graylevel_and_frame_ingest(vpp, pix, shape):
    """vppb_rgb_to_graylevel_u8 (domain_with_border form of colorspace_conversions.hh:22-47) and the fused ingest
    clone(_border) + fill_border_mirror + rgb_to_graylevel (examples/video_extruder.cc:46-48) vs the oracle"""
    from vpp_b200 import capi

    o = orc.load()
    ch = 3 if pix == "vuchar3" else 4
    data = rng(shape[0] * 7 + shape[1]).integers(0, 256, shape + (ch,), dtype=np.uint8)
    for b in (0, 2):
        if b > min(shape):
            continue
        for al in (128, 16, 4):
            hs = orc.HostImage(shape[0], shape[1], pix, border=b, aligned=al, data=data, fill_border="mirror" if b else None)
            exp = orc.HostImage(shape[0], shape[1], "u8", border=b, aligned=al)
            o.vo_rgb_to_graylevel(hs.ptr(), exp.ptr())
            src = vpp.Image2d.from_host(data, pix, border=b, aligned=al)
            if b:
                vpp.fill_border_mirror(src)
            got = vpp.rgb_to_graylevel(src)
            assert (got.border, got.nrows, got.ncols) == (b, shape[0], shape[1])
            assert np.array_equal(got.download(with_border=True), exp.get(True)), (b, al)
    # ingest: tight source without border -> gray with a mirror border of 3 (what fast9 / the pyramids want)
    for bb in (0, 1, 3):
        if bb > min(shape):
            continue
        hs = orc.HostImage(shape[0], shape[1], pix, data=data)
        exp = orc.HostImage(shape[0], shape[1], "u8", border=bb)
        o.vo_rgb_to_graylevel(hs.ptr(), exp.ptr())
        o.vo_fill_border_mirror(exp.ptr())
        for al in (128, 1):
            got = vpp.ingest_rgb_frame(vpp.Image2d.from_host(data, pix, aligned=al), bb)
            assert np.array_equal(got.download(with_border=True), exp.get(True)), (bb, al)
    # errors: output border wider than the input's (plain form), border wider than the image (ingest form)
    a, g = vpp.Image2d(8, 8, pix), vpp.Image2d(8, 8, "u8", border=2)
    assert capi.lib.vppb_rgb_to_graylevel_u8(a.ptr(), g.ptr(), None) == capi.VPPB_E_BORDER
    g9 = vpp.Image2d(8, 8, "u8", border=9)
    assert capi.lib.vppb_rgb_to_graylevel_u8_mirror(a.ptr(), g9.ptr(), None) == capi.VPPB_E_BORDER


@pytest.mark.parametrize("tag,nframes,th", [("7f_th4", 7, 4), ("9f_th5", 9, 5)])
def test_video_extruder_eventful_sequence_equals_reference_tables(vpp, tag, nframes, th):
    """the committed eventful sequence (an occluder appears, a patch is mirrored: keypoints die, merge and are re-detected)
    through the CUDA path (Python orchestration) against the tables the REFERENCE's own video_extruder_update produced
    (tests/golden/make_video_extruder_fixture.py); frame by frame against the oracle-backed orchestration too"""
    import os

    from vpp_b200 import video_extruder as ve
    from tests.oracle_video import OracleOps

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    frames = np.fromfile(os.path.join(gold, "video_extruder_frames_9x121x161.u8"), np.uint8).reshape(9, 121, 161)
    expected = np.fromfile(os.path.join(gold, "video_extruder_expected_%s.i32" % tag), np.int32).reshape(-1, 6)
    kw = dict(detector_th=th, keypoint_spacing=10, detector_period=3, max_trajectory_length=5, nscales=3, winsize=9, propagation=2)
    g, c = ve.video_extruder_init(121, 161), ve.video_extruder_init(121, 161)
    gops, cops = ve.GpuOps(), OracleOps()
    for f in range(1, nframes):
        ve.video_extruder_update(g, frames[f - 1], frames[f], gops, **kw)
        ve.video_extruder_update(c, frames[f - 1], frames[f], cops, **kw)
        assert np.array_equal(ve.state_table(g), ve.state_table(c)), "frame %d" % f
    assert np.array_equal(ve.state_table(g), expected)
    assert (expected[:, 2] == 0).any()  # the sequence does leave dead, not yet compacted keypoints behind


@pytest.mark.parametrize("name", ["core_tests", "algo_tests", "nbh_tests", "extruder_tests"])
def test_cpp_programs(gpu, name):
    """The C++ host API on the GPU: the reference's own tests rewritten with device kernels (core_tests, algo_tests - they ran on
    hardware before, but now go through the reworked pyramid launches and the lvalue kernel invocation of pixel_wise.hh) and
    the headers added after the last GPU run: box_nbh2d / window.hh / colorspace_conversions.hh (nbh_tests), video_extruder.hh /
    keypoint_container.hh (extruder_tests); built by build.sh"""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", name)
    assert os.path.exists(exe), "build.sh did not produce %s" % exe
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("schedule", ["levels", "dataflow", "antidiagonals"])
@pytest.mark.parametrize("density", ["sparse", "dense"])
def test_semi_dense_flow_level_schedule(vpp, monkeypatch, density, schedule):
    """The opt-in schedules of the propagation sweeps (the default, one cooperative launch that solves every sweep by
    relaxation, is what every other flow test runs): VPPB_SDOF_SCHEDULE=levels runs dependency levels of the marked cells,
    =dataflow one persistent launch per sweep with flags between cells, =antidiagonals one launch per anti-diagonal - same
    serial semantics, so the results must equal the oracle bit for bit.  sparse: blockwise-FAST keypoints (video_extruder's
    case: few levels); dense: a keypoint in every cell (the level schedule falls back to anti-diagonals wherever levels would
    not halve the launches)."""
    from tests import scenes

    monkeypatch.setenv("VPPB_SDOF_SCHEDULE", schedule)
    nr, nc = 121, 161
    f1, f2, _ = scenes.lk_pair(nr, nc, 4, seed=23, shift=(3.0, -2.0), margin=10)
    o = orc.load()
    if density == "sparse":
        h = orc.HostImage(nr, nc, "u8", border=3, data=f1, fill_border="mirror")
        k = np.zeros((f1.size, 2), np.int32)
        n = o.vo_fast9_u8(h.ptr(), 8, None, 2, 10, 0, k.ctypes.data, None, len(k))
        kps = np.ascontiguousarray(k[:n])
    else:
        rr, cc = np.meshgrid(np.arange(2, nr, 5), np.arange(2, nc, 5), indexing="ij")
        kps = np.stack([rr.ravel(), cc.ravel()], axis=1).astype(np.int32)
    for (ws, nscales, prop, patch) in [(9, 3, 2, 5), (7, 2, 3, 5), (9, 1, 1, 3)]:
        pos, dist, valid = vpp.semi_dense_optical_flow(kps, vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), winsize=ws, nscales=nscales,
                                                       propagation=prop, patchsize=patch)
        h1, h2 = orc.HostImage(nr, nc, "u8", data=f1), orc.HostImage(nr, nc, "u8", data=f2)
        m = len(kps)
        rpos, rdist, rvalid = np.zeros((m, 2), np.int32), np.zeros(m, np.int32), np.zeros(m, np.uint8)
        o.vo_semi_dense_flow(h1.ptr(), h2.ptr(), kps.ctypes.data, m, ws, nscales, 0, prop, patch, rpos.ctypes.data, rdist.ctypes.data, rvalid.ctypes.data)
        assert np.array_equal(valid, rvalid.astype(bool)) and valid.sum() > 10
        ok = rvalid > 0
        assert np.array_equal(pos[ok], rpos[ok]) and np.array_equal(dist[ok], rdist[ok]), (density, ws, nscales, prop, patch)


@pytest.mark.parametrize("shape,nlevels,grad", [((121, 163), 3, "vfloat2"), ((64, 96), 4, "vint2"), ((270, 481), 3, "vint2"), ((37, 50), 1, "vfloat2")])
def test_pyrlk_prepare_one_launch_equals_streams(vpp, monkeypatch, shape, nlevels, grad):
    """vppb_pyrlk_prepare as ONE cooperative launch (phases of concatenated work items, grid barriers between them) against its
    multi-stream form (one launch per step): both u8 pyramids and the gradient pyramid, every level, whole buffers incl. borders;
    and against the oracle's pyramids."""
    from tests import scenes
    from tests.oracle_ops import oracle_grad_pyramid, oracle_pyramid

    nr, nc = shape
    f1, f2, _ = scenes.lk_pair(nr, nc, 4, seed=nc, margin=10)
    I1, I2 = vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8")
    got = {}
    for form in ("fused", "streams"):
        monkeypatch.setenv("VPPB_PREPARE", form)
        prev, nxt = vpp.Pyramid2d((nr, nc), nlevels, 2, pixel="u8", border=4), vpp.Pyramid2d((nr, nc), nlevels, 2, pixel="u8", border=4)
        g = vpp.Pyramid2d((nr, nc), nlevels, 2, pixel=grad, border=4)
        vpp.pyrlk_prepare(I1, I2, prev, nxt, g)
        got[form] = [[l.download(with_border=True) for l in p.levels] for p in (prev, nxt, g)]
    for a, b in zip(got["fused"], got["streams"]):
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    o = orc.load()
    rprev = oracle_pyramid(f1, nlevels, "u8", 4, o)
    rgrad = oracle_grad_pyramid(rprev, grad, 4, o)
    for l in range(nlevels):
        assert np.array_equal(got["fused"][0][l], rprev[l].get(True))
        assert np.array_equal(got["fused"][2][l].view(np.int32), rgrad[l].get(True).view(np.int32))


@pytest.mark.parametrize("shape,shift,nscales,prop", [((96, 131), (6.0, -5.0), 1, 4), ((121, 161), (9.0, 8.0), 2, 3), ((64, 203), (-5.0, 6.0), 1, 2)])
def test_semi_dense_flow_long_propagation_chains(vpp, shape, shift, nscales, prop):
    """A shift the greedy descent cannot reach from a zero prediction: most cells start on a wrong local minimum and the few
    lucky ones spread their flow along the sweeps, cell after cell - the worst case of the relaxation schedule (one round
    per link of the chain, many rounds, work lists refilled round after round).  A keypoint in every cell; exact against
    the oracle."""
    from tests import scenes

    nr, nc = shape
    f1, f2, _ = scenes.lk_pair(nr, nc, 4, seed=31, shift=shift, margin=10)
    o = orc.load()
    rr, cc = np.meshgrid(np.arange(2, nr, 5), np.arange(2, nc, 5), indexing="ij")
    kps = np.stack([rr.ravel(), cc.ravel()], axis=1).astype(np.int32)
    m = len(kps)
    res = {}
    for p in (0, prop):
        pos, dist, valid = vpp.semi_dense_optical_flow(kps, vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8"), winsize=9, nscales=nscales,
                                                       propagation=p, patchsize=5)
        h1, h2 = orc.HostImage(nr, nc, "u8", data=f1), orc.HostImage(nr, nc, "u8", data=f2)
        rpos, rdist, rvalid = np.zeros((m, 2), np.int32), np.zeros(m, np.int32), np.zeros(m, np.uint8)
        o.vo_semi_dense_flow(h1.ptr(), h2.ptr(), kps.ctypes.data, m, 9, nscales, 0, p, 5, rpos.ctypes.data, rdist.ctypes.data, rvalid.ctypes.data)
        assert np.array_equal(valid, rvalid.astype(bool)) and valid.sum() > m // 2
        assert np.array_equal(pos, rpos) and np.array_equal(dist, rdist), (shape, shift, p)
        res[p] = pos
    moved = (res[0] != res[prop]).any(axis=1).sum()
    assert moved > m // 4, "the sweeps were meant to move many cells (moved %d of %d)" % (moved, m)


# ------------------------------------------------------------------ FAST9 band kernel: several TMA boxes per band, ragged tails
@pytest.mark.parametrize("shape", [(21, 4100), (9, 2017), (8, 2016), (17, 6050)])
def test_fast9_wide_images_multibox(vpp, shape):
    """The band kernel walks a row band in boxes of 2016 pixels: widths around the box size, a partial last band, a mask and
    both rings, exact keypoint arrays and scores against the oracle."""
    from tests import scenes
    from tests.test_gpu_parity import _oracle_fast

    img = scenes.rectangles_scene(shape[0], shape[1], seed=shape[1], nrect=shape[1] // 12)
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    mask = np.full(img.shape, 0xFF, dtype=np.uint8)
    mask[:, ::3] = 0x01
    mask[::4, :] = 0
    M = vpp.Image2d.from_host(mask, "u8")
    for ring in ("reference", "true"):
        for m, hm in ((None, None), (M, mask)):
            sc = []
            kps = vpp.fast9(G, 12, mask=m, ring=ring, scores=sc)
            rk, rs = _oracle_fast(img, 12, mask=hm, ring=0 if ring == "reference" else 1, want_scores=True)
            assert len(rk) > 10
            assert np.array_equal(kps, rk)
            assert np.array_equal(np.asarray(sc, dtype=np.int32), rs)


@pytest.mark.parametrize("th", [0, 127, 128, 200, 255, 300])
def test_fast9_threshold_extremes(vpp, th):
    """The packed prefilter switches formula at th >= 127 and nothing can pass at (th & 255) == 255; th is used modulo 256 as the
    reference's S::repeat(th) does."""
    from tests.test_gpu_parity import _oracle_fast

    r = np.random.default_rng(th)
    img = r.integers(0, 256, (40, 300), dtype=np.uint8)
    img[10:30, 50:200] = 0
    img[12:28, 60:190] = 255
    G = vpp.Image2d.from_host(img, "u8", border=3)
    vpp.fill_border_mirror(G)
    kps = vpp.fast9(G, th)
    rk, _ = _oracle_fast(img, th)
    assert np.array_equal(kps, rk)


# ------------------------------------------------------------------ row tiles: halo rows pulled from the neighbouring tiles by the box kernel itself
@pytest.mark.parametrize("pix,ntiles,rows,cols", [("vuchar3", 3, 37, 300), ("vuchar3", 2, 2, 700), ("u8", 4, 23, 1100), ("vuchar3", 8, 16, 190)])
def test_box5x5_row_tiles_read_neighbours(vpp, pix, ntiles, rows, cols):
    """vppb_box5x5_*_tiles: a frame cut into equally tall row tiles held in SEPARATE allocations (on the GPUs of a node: one
    per device, mapped with vppb_ipc_open); the tiles' top / bottom border rows are filled with garbage, the kernel must take
    the halo rows from the neighbours' domain rows.  Every tile of the result equals the oracle's full-frame box.  """
    from vpp_b200 import capi

    ch = 3 if pix == "vuchar3" else 1
    r = rng(900 + rows)
    H = ntiles * rows
    frame = r.integers(0, 256, (H, cols, ch) if ch > 1 else (H, cols), dtype=np.uint8)
    hs = orc.HostImage(H, cols, pix, border=2, data=frame, fill_border="mirror")
    hd = orc.HostImage(H, cols, pix)
    orc.load().vo_box5x5_u8(hs.ptr(), hd.ptr(), ch)
    want = hd.get()
    full = hs.get(with_border=True)  # (H + 4, cols + 4[, ch]) with the mirror border
    tiles, outs = [], []
    for t in range(ntiles):
        T = vpp.Image2d(rows, cols, pix, border=2)
        blk = np.array(full[t * rows:t * rows + rows + 4])
        if t > 0:
            blk[:2] = 0xA5   # the halo rows above are NOT in the tile
        if t < ntiles - 1:
            blk[-2:] = 0x5A  # nor the ones below
        T.upload(blk, with_border=True)
        tiles.append(T)
        outs.append(vpp.Image2d(rows, cols, pix))
    mapped = []
    emulated = hasattr(capi.lib, "vppb_emu_set_reverse")
    for T in tiles:
        h, off, m = (C.c_char * 64)(), C.c_int64(), capi.VppbImg()
        capi.check(capi.lib.vppb_ipc_export(T.ptr(), h, C.byref(off)))
        assert off.value == T.desc.base - T.desc.alloc
        if emulated:  # a CUDA IPC handle cannot be opened by the process that exported it: the open leg runs in the 2-rank tools/tiles_check.py
            capi.check(capi.lib.vppb_ipc_open(h, off.value, T.ptr(), C.byref(m)))
        else:
            C.memmove(C.byref(m), C.byref(T.desc), C.sizeof(capi.VppbImg))
        mapped.append(m)
    null = capi.VppbImg()
    ins = (capi.VppbImg * ntiles)(*[T.desc for T in tiles])
    ups = (capi.VppbImg * ntiles)(*[mapped[t - 1] if t > 0 else null for t in range(ntiles)])
    dns = (capi.VppbImg * ntiles)(*[mapped[t + 1] if t < ntiles - 1 else null for t in range(ntiles)])
    dst = (capi.VppbImg * ntiles)(*[o.desc for o in outs])
    fn = capi.lib.vppb_box5x5_u8c3_tiles if ch == 3 else capi.lib.vppb_box5x5_u8_tiles
    capi.check(fn(ins, ups, dns, dst, ntiles, None))
    capi.check(capi.lib.vppb_sync(None))
    for t in range(ntiles):
        assert np.array_equal(outs[t].download(), want[t * rows:(t + 1) * rows]), "tile %d" % t
    # one tile at a time, as each rank of a multi-GPU job calls it
    for t in range(ntiles):
        o = vpp.Image2d(rows, cols, pix)
        capi.check(fn(C.byref(ins[t]), C.byref(ups[t]), C.byref(dns[t]), o.ptr(), 1, None))
        assert np.array_equal(o.download(), want[t * rows:(t + 1) * rows]), "single tile %d" % t


# ------------------------------------------------------------------ video_extruder with the keypoint container in HBM (N3)
@pytest.mark.parametrize("tag,nframes,th", [("7f_th4", 7, 4), ("9f_th5", 9, 5)])
def test_video_extruder_device_container_equals_reference_tables(vpp, tag, nframes, th):
    """video_extruder_update_device: container, merge grid, score filter, detector mask, add / compact and trajectories are kernels
    (vppb_kpc_*).  Frame by frame its table equals the host orchestration's (which equals the oracle's), and the final table equals
    the one the REFERENCE's own video_extruder_update produced on the committed eventful sequence."""
    from vpp_b200 import video_extruder as ve

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    frames = np.fromfile(os.path.join(gold, "video_extruder_frames_9x121x161.u8"), np.uint8).reshape(9, 121, 161)
    expected = np.fromfile(os.path.join(gold, "video_extruder_expected_%s.i32" % tag), np.int32).reshape(-1, 6)
    kw = dict(detector_th=th, keypoint_spacing=10, detector_period=3, nscales=3, winsize=9, propagation=2)
    d, g = ve.DeviceVideoExtruderCtx(121, 161, max_trajectory_length=5), ve.video_extruder_init(121, 161)
    gops = ve.GpuOps()
    for f in range(1, nframes):
        ve.video_extruder_update_device(d, frames[f - 1], frames[f], **kw)
        ve.video_extruder_update(g, frames[f - 1], frames[f], gops, max_trajectory_length=5, **kw)
        assert np.array_equal(ve.device_state_table(d), ve.state_table(g)), "frame %d" % f
    assert np.array_equal(ve.device_state_table(d), expected)


def test_video_extruder_device_container_merge_cases(vpp):
    """The merge rule on hand-made containers (ages and cell sharing chosen to hit every branch: older arrives later, younger
    arrives later, ties, dead entries in the cell) against a direct transcription of video_extruder.hpp:59-84."""
    import ctypes as C

    from vpp_b200 import capi, video_extruder as ve
    from vpp_b200.ops import _DeviceBuffer

    r = rng(4242)
    for trial in range(20):
        n = int(r.integers(1, 60))
        pos = np.stack([r.integers(0, 40, n), r.integers(0, 50, n)], 1).astype(np.int32)
        age = r.integers(0, 4, n).astype(np.int32)
        ctx = ve.DeviceVideoExtruderCtx(40, 50, capacity=128, max_trajectory_length=3)
        det, cnt = _DeviceBuffer(pos.nbytes).from_host(pos), _DeviceBuffer(4).from_host(np.array([n], np.int32))
        capi.check(capi.lib.vppb_kpc_add_and_compact(ctx.handle, det.ptr, cnt.ptr, n, 0, None))  # all enter with age 1 ...
        assert capi.lib.vppb_kpc_size(ctx.handle) == n
        # ... then the ages are set through the flow step: entry i "moves in place" age[i] - 1 times, or is removed (target age 0)
        far = np.full((n, 2), -5, np.int32)
        for k in range(1, 4):
            newpos = np.where((age == 0)[:, None], far, pos).astype(np.int32)
            v = ((age == 0) | (age > k)).astype(np.uint8) if k == 1 else (age > k).astype(np.uint8)
            a, b = _DeviceBuffer(newpos.nbytes).from_host(newpos), _DeviceBuffer(n).from_host(v)
            capi.check(capi.lib.vppb_kpc_flow_update(ctx.handle, a.ptr, b.ptr, 40, 50, None))
        cur = age.copy()
        tab = _DeviceBuffer(n * 24)
        capi.check(capi.lib.vppb_kpc_state_table(ctx.handle, tab.ptr, None))
        got_age = tab.to_host(np.int32, n * 6).reshape(-1, 6)[:, 2]
        assert np.array_equal(got_age, cur), (got_age, cur)
        # reference merge, transcribed
        want, idx = cur.copy(), {}
        for i in range(n):
            cell = (pos[i, 0] // 10, pos[i, 1] // 10)
            j = idx.get(cell, -1)
            if j >= 0:
                other = want[j]
                if other < want[i]:
                    want[j] = 0; idx[cell] = i
                if other > want[i]:
                    want[i] = 0
            else:
                idx[cell] = i
        capi.check(capi.lib.vppb_kpc_merge(ctx.handle, 40, 50, 10, None))
        capi.check(capi.lib.vppb_kpc_state_table(ctx.handle, tab.ptr, None))
        assert np.array_equal(tab.to_host(np.int32, n * 6).reshape(-1, 6)[:, 2], want), trial


def test_out_of_frame_keypoints_are_skipped(vpp):
    """A keypoint outside the frame is an input error the reference answers with undefined behaviour; the C-ABI skips it: the
    semi-dense flow reports it invalid and leaves every other result untouched, fast9_scores gives it a score of 0."""
    from tests import scenes

    f1, f2, _ = scenes.lk_pair(121, 161, 4, seed=21, shift=(3.0, -2.0), margin=10)
    G = vpp.Image2d.from_host(f1, "u8", border=3)
    vpp.fill_border_mirror(G)
    kps = vpp.fast9(G, 8, blockwise=True, block_size=6)
    bad = np.array([[-1, 5], [5, -3], [121, 10], [10, 161], [4000, 4000], [-7, -7]], dtype=np.int32)
    both = np.concatenate([kps, bad])
    I1, I2 = vpp.Image2d.from_host(f1, "u8"), vpp.Image2d.from_host(f2, "u8")
    p0, d0, v0 = vpp.semi_dense_optical_flow(kps, I1, I2, winsize=9, nscales=3)
    p1, d1, v1 = vpp.semi_dense_optical_flow(both, I1, I2, winsize=9, nscales=3)
    n = len(kps)
    assert np.array_equal(p1[:n], p0) and np.array_equal(d1[:n], d0) and np.array_equal(v1[:n], v0) and v0.sum() > 50
    assert not v1[n:].any()
    sc = vpp.fast9_scores(G, 8, both)
    assert np.array_equal(sc[:n], vpp.fast9_scores(G, 8, kps)) and not sc[n:].any()
