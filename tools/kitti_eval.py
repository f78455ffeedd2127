#!/usr/bin/env python
"""evaluation/semi_dense_optical_flow/KITTI.cc on the GPU path: python tools/kitti_eval.py [kitti_root] [n_pairs]
(without a KITTI tree: synthetic pairs with known flow).  Prints the averages the reference writes to its result file."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vpp_b200 import capi, evaluation  # noqa: E402

capi.check(capi.lib.vppb_init(0))
root = sys.argv[1] if len(sys.argv) > 1 and os.path.isdir(sys.argv[1]) else None
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pairs = evaluation.kitti_pairs(root, n) if root else evaluation.synthetic_pairs(n)
res = evaluation.evaluate(pairs, nscales=3, winsize=9, propagation=2, min_scale=0, patchsize=5, detector_th=10, block_size=10)
res["data"] = root or "synthetic"
print(json.dumps(res))
