"""The library's stateless CUDA kernels (pyramid.cu, pixelwise.cu: one thread = independent loads/stores, no shared
memory / barriers / shuffles) compiled by g++ and run thread by thread on the CPU (tests/emu/), against the oracle.

What this catches without a GPU: wrong index arithmetic on ragged / tiny / unaligned geometries, stray writes (whole
buffers incl. row padding are compared, guard zones around every allocation), misaligned 64/128-bit accesses (UBSan
reports them; on the GPU they fault), and results that depend on the order in which threads run (every case is run with
the launch order forward and reversed).  It does not replace the `-m gpu` parity tests; kernels that use shared
memory, TMA, shuffles or atomics (box, FAST, LK, semi-dense flow, sum) are not emulated."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

from tests import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GUARD = 4096
UBSAN_LOG = os.path.join(ROOT, "tests", "emu", "_build", "ubsan.log")


@pytest.fixture(autouse=True)
def no_sanitizer_reports():
    yield
    logs = glob.glob(UBSAN_LOG + "*")
    text = "".join(open(f).read() for f in logs)
    for f in logs:
        os.remove(f)
    assert not text, "UBSan (misaligned / out-of-bounds access in an emulated kernel):\n" + text[:4000]


class EImg(C.Structure):  # vppb_img
    _fields_ = [("base", C.c_void_p), ("alloc", C.c_void_p), ("nrows", C.c_int32), ("ncols", C.c_int32), ("pitch", C.c_int32),
                ("border", C.c_int32), ("elem_bytes", C.c_int32), ("align", C.c_int32)]


@pytest.fixture(scope="module")
def emu(built):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    path = build_emu.build()
    for f in glob.glob(UBSAN_LOG + "*"):
        os.remove(f)
    os.environ["UBSAN_OPTIONS"] = "log_path=%s" % UBSAN_LOG  # read when the sanitizer runtime inside the .so initialises
    os.environ.setdefault("VPPB_EMU_LOG", os.path.join(ROOT, "tests", "emu", "_build", "emu_fail.log"))  # why the emulator aborted, if it does
    lib = C.CDLL(path)
    I, VP = C.POINTER(EImg), C.c_void_p
    for name, args in {"vppb_pw_add_i32": [I, I, I, VP], "vppb_fill": [I, VP, C.c_int, VP], "vppb_copy2d": [I, I, C.c_int, VP],
                       "vppb_copy2d_mirror": [I, I, VP], "vppb_fill_border_value": [I, VP, VP], "vppb_fill_border_mirror": [I, VP],
                       "vppb_fill_border_closest": [I, VP], "vppb_rgb_to_graylevel_u8": [I, I, VP], "vppb_rgb_to_graylevel_u8_mirror": [I, I, VP], "vppb_scharr_u8": [I, I, C.c_int, VP], "vppb_scharr_u8_mirror": [I, I, C.c_int, VP],
                       "vppb_lowpass_sub2": [I, I, C.c_int, VP], "vppb_lowpass_sub2_mirror": [I, I, C.c_int, VP],
                       "vppb_halo_pack": [I, C.c_int32, C.c_int, VP, VP], "vppb_halo_unpack": [I, C.c_int32, C.c_int, VP, VP],
                       "vppb_halo_pack_batch": [I, C.c_int32, C.c_int32, C.c_int, VP, VP], "vppb_halo_unpack_batch": [I, C.c_int32, C.c_int32, C.c_int, VP, VP]}.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, C.c_int
    lib.vppb_halo_bytes.argtypes, lib.vppb_halo_bytes.restype = [I, C.c_int32], C.c_int64
    lib.vppb_last_error.restype = C.c_char_p
    return lib


def guarded(nrows, ncols, pixel, border=0, aligned=128, data=None, seed=None):
    """HostImage re-homed between two guard zones; row padding and border start as 0xCD, or random bytes with `seed`"""
    h = orc.HostImage(nrows, ncols, pixel, border=border, aligned=aligned)
    raw = np.full(h.total + 2 * GUARD + aligned, 0xA5, np.uint8)
    off = GUARD + ((-(raw.ctypes.data + GUARD)) % aligned)
    h._raw, h.buf = raw, raw[off:off + h.total]
    h.buf[:] = 0xCD if seed is None else np.random.default_rng(seed).integers(0, 256, h.total, dtype=np.uint8)
    h.desc = orc.VoImg(h.buf.ctypes.data + h.origin, nrows, ncols, h.pitch, border, h.elem)
    h.guard_off = off
    if data is not None:
        h.set(data)
    return h


def twin(h):
    """a second image with the same geometry and byte-identical content (padding included)"""
    t = guarded(h.nrows, h.ncols, h.pixel, h.border, h.aligned)
    t.buf[:] = h.buf
    return t


def guards_ok(*imgs):
    return all((h._raw[:h.guard_off] == 0xA5).all() and (h._raw[h.guard_off + h.total:] == 0xA5).all() for h in imgs)


def E(h):
    return C.byref(EImg(h.desc.base, None, h.nrows, h.ncols, h.pitch, h.border, h.elem, h.aligned))


def data_for(pixel, nr, nc, seed):
    dt, ch = orc.PIXEL_TYPES[pixel]
    r = np.random.default_rng(seed)
    shape = (nr, nc) + ((ch,) if ch > 1 else ())
    if dt == np.float32:
        return ((r.random(shape) - 0.5) * 300).astype(dt)
    info = np.iinfo(dt)
    return r.integers(max(info.min, -1000), min(info.max, 1000) + 1, shape).astype(dt)


def both_orders(emu):
    for rev in (0, 1):
        emu.vppb_emu_set_reverse(rev)
        yield rev
    emu.vppb_emu_set_reverse(0)


GEOMS = [(1, 1), (2, 3), (4, 4), (5, 9), (8, 8), (9, 6), (16, 40), (33, 70), (64, 65), (97, 130), (31, 257), (130, 17)]


@pytest.mark.parametrize("aligned", [128, 16, 4, 1])
def test_scharr_and_fused_mirror(emu, aligned):
    o = orc.load()
    for nr, nc in GEOMS:
        for gpix, asf in (("vint2", 0), ("vfloat2", 1)):
            for b in (0, 1, 3, 5):
                if b > nr or b > nc:
                    continue
                src = guarded(nr, nc, "u8", border=2, aligned=aligned, data=data_for("u8", nr, nc, nr * 7 + nc))
                o.vo_fill_border_mirror(src.ptr())
                exp = guarded(nr, nc, gpix, border=b, aligned=aligned)
                o.vo_scharr_u8(src.ptr(), exp.ptr(), asf)
                for rev in both_orders(emu):
                    got = guarded(nr, nc, gpix, border=b, aligned=aligned)
                    assert emu.vppb_scharr_u8(E(src), E(got), asf, None) == 0, emu.vppb_last_error()
                    assert np.array_equal(got.buf, exp.buf), ("scharr", nr, nc, gpix, b, aligned, rev)
                    assert guards_ok(got, src)
                if b:
                    o.vo_fill_border_mirror(exp.ptr())
                    for rev in both_orders(emu):
                        got = guarded(nr, nc, gpix, border=b, aligned=aligned)
                        assert emu.vppb_scharr_u8_mirror(E(src), E(got), asf, None) == 0, emu.vppb_last_error()
                        assert np.array_equal(got.buf, exp.buf), ("scharr_mirror", nr, nc, gpix, b, aligned, rev)
                        assert guards_ok(got, src)


@pytest.mark.parametrize("aligned", [128, 32, 4, 1])
@pytest.mark.parametrize("pix,kind", [("u8", 0), ("vint2", 1), ("vfloat2", 2)])
def test_lowpass_sub2_and_fused_mirror(emu, pix, kind, aligned):
    o = orc.load()
    for nr, nc in GEOMS:
        if nr < 2 or nc < 2:
            continue  # the parent needs a mirror-filled border of 2
        parent = guarded(nr, nc, pix, border=2, aligned=aligned, data=data_for(pix, nr, nc, nr * 13 + nc))
        o.vo_fill_border_mirror(parent.ptr())
        for onr, onc in {(1 + nr // 2, 1 + nc // 2), ((nr + 1) // 2, (nc + 1) // 2), (max(nr // 2 - 1, 1), max(nc // 2, 1))}:
            for b in (0, 2, 3, 6):
                if b > onr or b > onc:
                    continue
                exp = guarded(onr, onc, pix, border=b, aligned=aligned)
                o.vo_lowpass_sub2(parent.ptr(), exp.ptr(), kind)
                for rev in both_orders(emu):
                    got = guarded(onr, onc, pix, border=b, aligned=aligned)
                    assert emu.vppb_lowpass_sub2(E(parent), E(got), kind, None) == 0, emu.vppb_last_error()
                    assert np.array_equal(got.buf, exp.buf), ("lowpass", pix, nr, nc, onr, onc, b, aligned, rev)
                    assert guards_ok(got, parent)
                if b:
                    o.vo_fill_border_mirror(exp.ptr())
                    for rev in both_orders(emu):
                        got = guarded(onr, onc, pix, border=b, aligned=aligned)
                        assert emu.vppb_lowpass_sub2_mirror(E(parent), E(got), kind, None) == 0, emu.vppb_last_error()
                        assert np.array_equal(got.buf, exp.buf), ("lowpass_mirror", pix, nr, nc, onr, onc, b, aligned, rev)
                        assert guards_ok(got, parent)


def test_fused_entries_reject_oversized_borders(emu):
    big = guarded(5, 5, "u8", border=2, data=data_for("u8", 5, 5, 1))
    small = guarded(3, 3, "u8", border=4)
    g = guarded(3, 3, "vint2", border=4)
    assert emu.vppb_lowpass_sub2_mirror(E(big), E(small), 0, None) == -3  # VPPB_E_BORDER
    assert emu.vppb_scharr_u8_mirror(E(big), E(g), 0, None) == -3
    src = guarded(3, 3, "u8")
    assert emu.vppb_copy2d_mirror(E(src), E(small), None) == -3
    assert emu.vppb_copy2d_mirror(E(big), E(small), None) == -2  # VPPB_E_ARG: domains differ


@pytest.mark.parametrize("pix", ["u8", "vuchar3", "i32", "vint2", "vfloat2"])
def test_copy_mirror_and_pyramid_chain(emu, pix):
    o = orc.load()
    for nr, nc in GEOMS:
        for b, al_src, al_dst in ((0, 128, 128), (2, 128, 128), (3, 16, 128), (3, 128, 4), (5, 1, 1)):
            if b > nr or b > nc:
                continue
            src = guarded(nr, nc, pix, border=1, aligned=al_src, data=data_for(pix, nr, nc, nr + 31 * nc), seed=5)
            exp = guarded(nr, nc, pix, border=b, aligned=al_dst)
            o.vo_copy(src.ptr(), exp.ptr(), 0)
            o.vo_fill_border_mirror(exp.ptr())
            for rev in both_orders(emu):
                got = guarded(nr, nc, pix, border=b, aligned=al_dst)
                assert emu.vppb_copy2d_mirror(E(src), E(got), None) == 0, emu.vppb_last_error()
                assert np.array_equal(got.buf, exp.buf), ("copy_mirror", pix, nr, nc, b, al_src, al_dst, rev)
                assert guards_ok(got, src)
    if pix in ("u8", "vint2", "vfloat2"):  # Pyramid2d.update as vpp_b200/ops.py issues it: copy_mirror, then one fused launch per level
        from tests.oracle_ops import oracle_pyramid

        kind = {"u8": 0, "vint2": 1, "vfloat2": 2}[pix]
        for nr, nc, b in ((101, 77, 3), (64, 96, 4), (270, 481, 2)):
            base = data_for(pix, nr, nc, 99)
            src = guarded(nr, nc, pix, data=base)
            ref = oracle_pyramid(base, 3, pix, b, o)
            for rev in both_orders(emu):
                levels = [guarded(l.nrows, l.ncols, pix, border=b) for l in ref]
                assert emu.vppb_copy2d_mirror(E(src), E(levels[0]), None) == 0
                for i in (1, 2):
                    assert emu.vppb_lowpass_sub2_mirror(E(levels[i - 1]), E(levels[i]), kind, None) == 0
                for i in range(3):
                    assert np.array_equal(levels[i].get(True).view(np.uint8), ref[i].get(True).view(np.uint8)), (pix, nr, nc, i, rev)
                assert guards_ok(*levels)


@pytest.mark.parametrize("aligned", [128, 16, 1])
def test_add_fill_copy_borders(emu, aligned):
    o = orc.load()
    for nr, nc in GEOMS:
        b_, c_ = data_for("i32", nr, nc, 1) * 1000003, data_for("i32", nr, nc, 2) * 2000003
        hb, hc = guarded(nr, nc, "i32", aligned=aligned, data=b_), guarded(nr, nc, "i32", aligned=aligned, data=c_)
        exp = guarded(nr, nc, "i32", aligned=aligned)
        o.vo_pw_add_i32(exp.ptr(), hb.ptr(), hc.ptr())
        for rev in both_orders(emu):
            got = guarded(nr, nc, "i32", aligned=aligned)
            assert emu.vppb_pw_add_i32(E(got), E(hb), E(hc), None) == 0
            assert np.array_equal(got.buf, exp.buf) and guards_ok(got), ("add", nr, nc, aligned, rev)
        for pix in ("u8", "vuchar3", "i32", "vint2"):
            dt, ch = orc.PIXEL_TYPES[pix]
            val = np.arange(7, 7 + ch).astype(dt)
            for b in (0, 1, 4):
                for wb in (0, 1):
                    exp, got = guarded(nr, nc, pix, border=b, aligned=aligned), guarded(nr, nc, pix, border=b, aligned=aligned)
                    o.vo_fill(exp.ptr(), val.ctypes.data, wb)
                    assert emu.vppb_fill(E(got), val.ctypes.data, wb, None) == 0
                    assert np.array_equal(got.buf, exp.buf) and guards_ok(got), ("fill", pix, nr, nc, b, wb, aligned)
                src = guarded(nr, nc, pix, border=b, aligned=aligned, data=data_for(pix, nr, nc, 3), seed=8)
                for wb in (0, 1):
                    exp, got = guarded(nr, nc, pix, border=b, aligned=aligned), guarded(nr, nc, pix, border=b, aligned=aligned)
                    o.vo_copy(src.ptr(), exp.ptr(), wb)
                    assert emu.vppb_copy2d(E(src), E(got), wb, None) == 0
                    assert np.array_equal(got.buf, exp.buf) and guards_ok(got), ("copy", pix, nr, nc, b, wb, aligned)
                if b and b <= nr and b <= nc:
                    for mode in ("value", "mirror", "closest"):
                        if mode == "closest" and pix == "vint2":
                            continue
                        exp = twin(src)
                        got = twin(src)
                        if mode == "value":
                            o.vo_fill_border_value(exp.ptr(), val.ctypes.data)
                            rc = emu.vppb_fill_border_value(E(got), val.ctypes.data, None)
                        else:
                            getattr(o, "vo_fill_border_" + mode)(exp.ptr())
                            rc = getattr(emu, "vppb_fill_border_" + mode)(E(got), None)
                        assert rc == 0 and np.array_equal(got.buf, exp.buf) and guards_ok(got), (mode, pix, nr, nc, b, aligned)


def test_halo_pack_unpack_roundtrip(emu):
    """pack the 2 rows next to each tile edge, unpack them into the neighbour's border rows: the border rows of tile k+1
    must equal the last rows of tile k (tiles.py protocol), single and batched entry points"""
    nr, nc, halo, nimg = 12, 37, 2, 3
    tiles = [guarded(nr, nc, "vuchar3", border=halo, data=data_for("vuchar3", nr, nc, 40 + i)) for i in range(nimg)]
    per = emu.vppb_halo_bytes(E(tiles[0]), halo)
    assert per == halo * (nc + 2 * halo) * 3  # full buffer width, column border included
    for rev in both_orders(emu):
        top, bot = np.zeros(per, np.uint8), np.zeros(per, np.uint8)
        assert emu.vppb_halo_pack(E(tiles[0]), halo, 0, top.ctypes.data, None) == 0  # side 0: first rows
        assert emu.vppb_halo_pack(E(tiles[0]), halo, 1, bot.ctypes.data, None) == 0  # side 1: last rows
        dst = twin(tiles[1])
        assert emu.vppb_halo_unpack(E(dst), halo, 0, bot.ctypes.data, None) == 0      # into the rows above row 0
        assert emu.vppb_halo_unpack(E(dst), halo, 1, top.ctypes.data, None) == 0      # into the rows below the last
        v, s = dst.get(True), tiles[0].get(True)  # frames: row index = r + halo
        assert np.array_equal(v[:halo], s[nr:nr + halo])          # rows [-halo, 0) <- the neighbour's rows [nr-halo, nr)
        assert np.array_equal(v[-halo:], s[halo:2 * halo])        # rows [nr, nr+halo) <- the neighbour's rows [0, halo)
        assert np.array_equal(v[halo:-halo], tiles[1].get(True)[halo:-halo])  # the domain rows are untouched
        assert guards_ok(dst)
        # batched == the single calls image by image
        descs = (EImg * nimg)(*[EImg(t.desc.base, None, t.nrows, t.ncols, t.pitch, t.border, t.elem, t.aligned) for t in tiles])
        staging = np.zeros(per * nimg, np.uint8)
        assert emu.vppb_halo_pack_batch(descs, nimg, halo, 1, staging.ctypes.data, None) == 0
        for i in range(nimg):
            one = np.zeros(per, np.uint8)
            assert emu.vppb_halo_pack(E(tiles[i]), halo, 1, one.ctypes.data, None) == 0
            assert np.array_equal(staging[i * per:(i + 1) * per], one), i
        outs = [twin(t) for t in tiles]
        odesc = (EImg * nimg)(*[EImg(t.desc.base, None, t.nrows, t.ncols, t.pitch, t.border, t.elem, t.aligned) for t in outs])
        assert emu.vppb_halo_unpack_batch(odesc, nimg, halo, 0, staging.ctypes.data, None) == 0
        for i in range(nimg):
            single = twin(tiles[i])
            assert emu.vppb_halo_unpack(E(single), halo, 0, staging[i * per:].ctypes.data, None) == 0
            assert np.array_equal(outs[i].buf, single.buf), i


@pytest.mark.parametrize("aligned", [128, 16, 4, 1])
@pytest.mark.parametrize("pix", ["vuchar3", "vuchar4"])
def test_rgb_to_graylevel_and_ingest(emu, pix, aligned):
    o = orc.load()
    for nr, nc in GEOMS + [(3, 16), (5, 32), (2, 47), (7, 129)]:
        data = data_for(pix, nr, nc, nr * 5 + nc)
        for b in (0, 1, 3):
            if b > nr or b > nc:
                continue
            src = guarded(nr, nc, pix, border=b, aligned=aligned, data=data, seed=3)
            if b:
                o.vo_fill_border_mirror(src.ptr())
            exp = guarded(nr, nc, "u8", border=b, aligned=aligned)
            o.vo_rgb_to_graylevel(src.ptr(), exp.ptr())
            for rev in both_orders(emu):
                got = guarded(nr, nc, "u8", border=b, aligned=aligned)
                assert emu.vppb_rgb_to_graylevel_u8(E(src), E(got), None) == 0, emu.vppb_last_error()
                assert np.array_equal(got.buf, exp.buf) and guards_ok(got, src), ("gray", pix, nr, nc, b, aligned, rev)
            # ingest: the source border is never read (random bytes there), the result carries the mirror border
            tight = guarded(nr, nc, pix, border=0, aligned=aligned, data=data)
            exp = guarded(nr, nc, "u8", border=b, aligned=aligned)
            o.vo_rgb_to_graylevel(tight.ptr(), exp.ptr())
            o.vo_fill_border_mirror(exp.ptr())
            for rev in both_orders(emu):
                got = guarded(nr, nc, "u8", border=b, aligned=aligned)
                assert emu.vppb_rgb_to_graylevel_u8_mirror(E(tight), E(got), None) == 0, emu.vppb_last_error()
                assert np.array_equal(got.buf, exp.buf) and guards_ok(got, tight), ("ingest", pix, nr, nc, b, aligned, rev)
