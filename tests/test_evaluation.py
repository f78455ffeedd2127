"""The KITTI-style evaluation harness: vpp_b200.evaluation.flow_error_stats against the C restatement of
kitti::flow_error_stats (evaluation/utils/kitti.hh:75-134), the 16-bit flow decoding of load_flow (:9-21), and - on the
GPU - the whole KITTI.cc loop body: identical statistics whether the flow comes from the CUDA path or from the oracle."""
import numpy as np
import pytest

from tests import oracle as orc


def _rand_flow(rng, shape, density):
    f = np.zeros(shape + (3,), dtype=np.float32)
    m = rng.random(shape) < density
    f[..., 0][m] = rng.uniform(-12, 12, m.sum()).astype(np.float32)
    f[..., 1][m] = rng.uniform(-12, 12, m.sum()).astype(np.float32)
    f[..., 2][m] = 1.0
    return f


@pytest.mark.parametrize("shape,d1,d2", [((37, 53), 0.3, 0.7), ((64, 64), 0.0, 0.5), ((20, 31), 1.0, 1.0), ((8, 9), 0.2, 0.0)])
def test_flow_error_stats_equals_reference_restatement(shape, d1, d2):
    from vpp_b200 import evaluation  # imports the CUDA library's ctypes binding; no GPU call is made here

    rng = np.random.default_rng(shape[0])
    flow, ref = _rand_flow(rng, shape, d1), _rand_flow(rng, shape, d2)
    ref[..., :2] += np.where(rng.random(shape)[..., None] < 0.5, 0, 4).astype(np.float32) * (ref[..., 2:3] > 0)
    got = evaluation.flow_error_stats(flow, ref)
    hf, hr = orc.HostImage(shape[0], shape[1], "vfloat3", data=flow), orc.HostImage(shape[0], shape[1], "vfloat3", data=ref)
    hm = orc.HostImage(shape[0], shape[1], "u8")
    out = np.zeros(6, dtype=np.float32)
    n = orc.load().vo_flow_error_stats(hf.ptr(), hr.ptr(), out.ctypes.data, hm.ptr())
    assert n == got["compared"]
    for k, name in enumerate(("n1", "n3", "n5", "n10", "avg", "density")):
        assert np.float32(got[name]) == out[k], name
    assert np.array_equal(got["errors_map"], hm.get())
    assert np.all(np.diff(got["errors"]) >= 0)


def test_kitti_flow_decoding():
    from vpp_b200 import evaluation

    png = np.zeros((2, 2, 3), dtype=np.uint16)
    png[0, 0] = (1, (1 << 15) + 64, (1 << 15) - 128)   # BGR as OpenCV reads it: valid, v, u
    f = evaluation.decode_kitti_flow(png)
    assert tuple(f[0, 0]) == (1.0, -2.0, 1.0) and tuple(f[1, 1]) == (-512.0, -512.0, 0.0)


@pytest.mark.gpu
def test_kitti_loop_gpu_equals_oracle(gpu):
    """KITTI.cc:126-193 on synthetic KITTI-sized pairs: the flow image of the CUDA path equals the one built from the oracle's
    rgb_to_graylevel / blockwise FAST9 / semi-dense flow, so every statistic is identical; and most vectors are right."""
    import ctypes as C

    from vpp_b200 import evaluation

    o = orc.load()
    for f1, f2, ref in evaluation.synthetic_pairs(2, nrows=187, ncols=400, seed=3):
        flow, nk = evaluation.semi_dense_flow_image(f1, f2, nscales=3, winsize=9, propagation=2)
        H, W = f1.shape[:2]
        g = []
        for f in (f1, f2):
            hi = orc.HostImage(H, W, "vuchar3", border=9, data=f, fill_border="mirror")
            hg = orc.HostImage(H, W, "u8", border=9)
            o.vo_rgb_to_graylevel(hi.ptr(), hg.ptr())
            g.append(hg)
        cap = H * W
        kps = np.zeros((cap, 2), dtype=np.int32)
        n = o.vo_fast9_u8(g[0].ptr(), 10, None, 2, 10, 0, kps.ctypes.data, None, cap)
        assert n == nk
        kps = np.ascontiguousarray(kps[:n])
        rp, rd, rv = np.zeros((n, 2), np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        o.vo_semi_dense_flow(g[0].ptr(), g[1].ptr(), kps.ctypes.data, n, 9, 3, 0, 2, 5, rp.ctypes.data, rd.ctypes.data, rv.ctypes.data)
        want = np.zeros((H, W, 3), dtype=np.float32)
        v = rv.astype(bool)
        want[kps[v, 0], kps[v, 1], 0:2] = (rp[v] - kps[v]).astype(np.float32)
        want[kps[v, 0], kps[v, 1], 2] = 1.0
        assert np.array_equal(flow, want)
        st = evaluation.flow_error_stats(flow, ref)
        assert st["compared"] > 100 and st["n3"] < 30.0
