"""KITTI-style evaluation of the semi-dense flow: the reference's only accuracy yardstick.

Mirrors evaluation/semi_dense_optical_flow/KITTI.cc:126-193 (per training pair: rgb_to_graylevel, clone with border,
blockwise FAST9, semi_dense_optical_flow, flow_error_stats, averages over the pairs) and evaluation/utils/kitti.hh:75-134
(flow_error_stats) and :9-21 (load_flow's 16-bit PNG decoding).  Pixel work runs on the GPU path (frame ingest, FAST9,
semi-dense flow); the statistics are a handful of reductions over the sparse flow image and stay on the host as in the
reference.  Without a KITTI tree (none in this repository's environment) `synthetic_pairs` supplies frame pairs whose true
flow is known, so the harness itself is exercised end to end."""
import os

import numpy as np

from . import ops
from .image import Image2d


def decode_kitti_flow(png_u16_bgr):
    """load_flow (kitti.hh:9-21): OpenCV reads the 16-bit PNG as BGR = (valid, v, u); the reference stores
    (float(v[1]) - 2^15) / 64, (float(v[2]) - 2^15) / 64, float(v[0])."""
    a = np.asarray(png_u16_bgr)
    out = np.zeros(a.shape[:2] + (3,), dtype=np.float32)
    out[..., 0] = (a[..., 1].astype(np.float32) - float(1 << 15)) / np.float32(64.0)
    out[..., 1] = (a[..., 2].astype(np.float32) - float(1 << 15)) / np.float32(64.0)
    out[..., 2] = a[..., 0].astype(np.float32)
    return out


def flow_error_stats(flow, ref):
    """kitti::flow_error_stats (kitti.hh:75-134).  flow, ref: (H, W, 3) float32, channel 2 > 0 marks a defined vector.
    Returns n1 / n3 / n5 / n10 (percent of the compared vectors whose end-point error exceeds 1 / 3 / 5 / 10 px), avg (mean
    end-point error), density (percent of pixels with a computed flow), errors (sorted), errors_map (u8, min(err * 20, 255))."""
    flow = np.asarray(flow, dtype=np.float32)
    ref = np.asarray(ref, dtype=np.float32)
    has = flow[..., 2] > 0
    both = has & (ref[..., 2] > 0)
    d = flow[..., :2][both] - ref[..., :2][both]
    err = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
    emap = np.zeros(flow.shape[:2], dtype=np.uint8)
    emap[both] = np.minimum(err * np.float32(20.0), np.float32(255.0)).astype(np.uint8)  # float -> uchar truncation
    n = len(err)
    pct = lambda k: float(np.float32(100.0) * np.float32((err > k).sum()) / np.float32(n)) if (err > k).sum() else 0.0
    esum = np.float32(0.0)
    for e in err:  # the reference accumulates in raster order in float
        esum = np.float32(esum + e)
    return {"n1": pct(1.0), "n3": pct(3.0), "n5": pct(5.0), "n10": pct(10.0), "avg": float(esum / np.float32(n if n else 1)),
            "density": float(np.float32(100.0) * np.float32(has.sum()) / np.float32(flow.shape[0] * flow.shape[1])),
            "errors": np.sort(err), "errors_map": emap, "compared": int(n)}


def semi_dense_flow_image(frame1_rgb, frame2_rgb, winsize=9, nscales=1, min_scale=0, propagation=2, patchsize=5, detector_th=10, block_size=10,
                          stream=None):
    """One pair of KITTI.cc's loop body on the GPU path: returns (flow image (H, W, 3) float32, number of keypoints)."""
    f1 = Image2d.from_host(np.ascontiguousarray(frame1_rgb), "vuchar3")
    f2 = Image2d.from_host(np.ascontiguousarray(frame2_rgb), "vuchar3")
    g1 = ops.ingest_rgb_frame(f1, winsize, stream=stream)   # rgb_to_graylevel + clone(_border = winsize) + mirror, one launch
    g2 = ops.ingest_rgb_frame(f2, winsize, stream=stream)
    kps = ops.fast9(g1, detector_th, blockwise=True, block_size=block_size, stream=stream)
    pos, dist, valid = ops.semi_dense_optical_flow(kps, g1, g2, winsize=winsize, nscales=nscales, min_scale=min_scale, propagation=propagation,
                                                   patchsize=patchsize, stream=stream)
    flow = np.zeros(frame1_rgb.shape[:2] + (3,), dtype=np.float32)
    k = kps[valid]
    flow[k[:, 0], k[:, 1], 0:2] = (pos[valid] - k).astype(np.float32)
    flow[k[:, 0], k[:, 1], 2] = 1.0
    return flow, len(kps)


def evaluate(pairs, **params):
    """KITTI.cc's main loop: pairs yields (frame1_rgb, frame2_rgb, ref_flow); returns the averages the reference writes
    (_errors = mean n3, _nkeypoints) plus the per-pair statistics."""
    per_pair, n3, nk = [], [], []
    for f1, f2, ref in pairs:
        flow, n = semi_dense_flow_image(f1, f2, **params)
        st = flow_error_stats(flow, ref)
        per_pair.append({k: v for k, v in st.items() if k not in ("errors", "errors_map")})
        n3.append(st["n3"]); nk.append(n)
    return {"errors": float(np.mean(n3)) if n3 else 0.0, "nkeypoints": float(np.mean(nk)) if nk else 0.0, "pairs": per_pair}


def synthetic_pairs(n, nrows=375, ncols=1242, seed=0, max_shift=6.0):
    """KITTI-sized textured frame pairs with a known piecewise-constant integer flow (the second frame is the first one
    shifted region by region), ground truth valid where both regions agree.  (row, col) order like the rest of the path."""
    rng = np.random.default_rng(seed)
    for _ in range(n):
        base = rng.integers(0, 256, (nrows // 4 + 8, ncols // 4 + 8), dtype=np.uint8)
        base = np.kron(base, np.ones((4, 4), dtype=np.uint8))[:nrows + 32, :ncols + 32]
        base = (base.astype(np.int32) + rng.integers(-6, 7, base.shape)).clip(0, 255).astype(np.uint8)
        f1 = base[16:16 + nrows, 16:16 + ncols]
        ref = np.zeros((nrows, ncols, 3), dtype=np.float32)
        f2 = np.zeros_like(f1)
        half = ncols // 2
        for c0, c1 in ((0, half), (half, ncols)):
            dr, dc = [int(v) for v in rng.integers(-int(max_shift), int(max_shift) + 1, 2)]
            f2[:, c0:c1] = base[16 - dr:16 - dr + nrows, 16 - dc + c0:16 - dc + c1]
            ref[:, c0:c1, 0], ref[:, c0:c1, 1] = dr, dc
            ref[8:-8, c0 + 8:c1 - 8, 2] = 1.0
        yield (np.repeat(f1[:, :, None], 3, axis=2), np.repeat(f2[:, :, None], 3, axis=2), ref)


def kitti_pairs(kitti_root, n):
    """foreach_training_pair (kitti.hh:24-51) for a real KITTI flow tree: training/colored_0/%06d_10.png, _11.png, flow_noc."""
    import cv2

    for i in range(n):
        a = cv2.imread(os.path.join(kitti_root, "training", "colored_0", "%06d_10.png" % i))
        b = cv2.imread(os.path.join(kitti_root, "training", "colored_0", "%06d_11.png" % i))
        r = cv2.imread(os.path.join(kitti_root, "training", "flow_noc", "%06d_10.png" % i), cv2.IMREAD_UNCHANGED)
        if a is None or b is None or r is None:
            break
        ref = decode_kitti_flow(r)
        ref = ref[..., [1, 0, 2]]  # KITTI stores (u = x, v = y); the path is (row, col)
        yield a, b, ref
