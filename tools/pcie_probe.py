#!/usr/bin/env python
"""What the host link of the box allows: pinned host <-> device copy rates, one direction and both at once, in chunks of one
1080p vuchar3 frame (6.2 MB, what the e2e leg of bench.py moves per call) and of 64 MB.  Usage (GPU box): python tools/pcie_probe.py"""
import json
import time

import torch

dev = torch.device("cuda:0")
out = {}
for label, chunk in (("frame_6MB", 1920 * 1080 * 3), ("chunk_64MB", 64 << 20)):
    n = max(4, (1 << 30) // chunk)  # ~1 GB per direction per pass
    hin = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(4)]
    hout = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(4)]
    din = [torch.empty(chunk, dtype=torch.uint8, device=dev) for _ in range(4)]
    dout = [torch.empty(chunk, dtype=torch.uint8, device=dev) for _ in range(4)]
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def up():
        with torch.cuda.stream(s_up):
            for i in range(n):
                din[i % 4].copy_(hin[i % 4], non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            for i in range(n):
                hout[i % 4].copy_(dout[i % 4], non_blocking=True)

    def both():
        up()
        down()

    res = {}
    for name, fn, mult in (("h2d", up, 1), ("d2h", down, 1), ("both", both, 1)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        res[name + "_GBps_per_direction"] = n * chunk / dt / 1e9
    out[label] = res
print(json.dumps(out))
