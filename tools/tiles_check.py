#!/usr/bin/env python
"""N-rank check + timing of the row-tiled 5x5 box (run under torchrun, one rank per GPU):
  fused : vppb_box5x5_u8c3_tiles - halo rows pulled from the neighbours' memory (CUDA IPC) inside the box kernel
  nccl  : vppb_halo_exchange (one grouped NCCL send/recv through the C-ABI) + vppb_box5x5_u8c3_batch on the tiles
Both are compared per rank with the oracle's full-frame result; step times are CUDA events, max over ranks.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/tiles_check.py [H W frames]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vpp_b200 as vpp  # noqa: E402
from vpp_b200 import capi, tiles  # noqa: E402
from tests import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    H, W, nframes = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4320, 7680, 8)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.check(capi.lib.vppb_init(local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    r0, r1 = tiles.tile_rows(H, rank, world)
    th = r1 - r0
    rng = np.random.default_rng(42)
    uniq = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(2)]
    pad = [np.pad(f, ((2, 2), (2, 2), (0, 0)), mode="symmetric") for f in uniq]
    src, dst = [], []
    for i in range(nframes):
        s = vpp.Image2d(th, W, "vuchar3", border=2)
        blk = np.array(pad[i % 2][r0:r1 + 4])
        if rank > 0:
            blk[:2] = 0xA5          # interior tiles do not hold their halo rows
        if rank < world - 1:
            blk[-2:] = 0x5A
        s.upload(blk, with_border=True)
        src.append(s)
        dst.append(vpp.Image2d(th, W, "vuchar3"))
    dist.barrier()
    ups, downs, opened = tiles.open_neighbour_tiles(dist, rank, world, src)
    ins = (capi.VppbImg * nframes)(*[s.desc for s in src])
    outs = (capi.VppbImg * nframes)(*[d.desc for d in dst])

    want = []
    for k in range(2):
        hs = orc.HostImage(th, W, "vuchar3", border=2, aligned=32)
        hs.set(pad[k][r0:r1 + 4], with_border=True)
        hd = orc.HostImage(th, W, "vuchar3", aligned=32)
        orc.load(omp=True).vo_box5x5_u8(hs.ptr(), hd.ptr(), 3)
        want.append(hd.get())

    def check(tag):
        torch.cuda.synchronize()
        ok = all(np.array_equal(dst[i].download(), want[i % 2]) for i in (0, 1, nframes - 1))
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"world": world, "frame": [H, W], "frames_per_step": nframes, "tile_rows": th}
    # ---- fused
    fused = lambda: capi.check(capi.lib.vppb_box5x5_u8c3_tiles(ins, ups, downs, outs, nframes, sp))
    fused()
    res["fused_parity"] = check("fused")
    res["fused_ms_per_step"] = timed(fused)
    # no neighbours at all (what the kernel costs without the peer reads; result wrong at the tile edges by construction)
    res["no_halo_ms_per_step"] = timed(lambda: capi.check(capi.lib.vppb_box5x5_u8c3_batch(ins, outs, nframes, sp)))
    # ---- materialised halos through the C-ABI NCCL exchange
    for d in dst:
        vpp.fill(d, [0, 0, 0])
    try:
        comm = tiles.nccl_comm(dist, rank, world)
        ex = lambda: capi.check(capi.lib.vppb_halo_exchange(comm, rank, world, ins, nframes, 2, sp))
        box = lambda: capi.check(capi.lib.vppb_box5x5_u8c3_batch(ins, outs, nframes, sp))
        ex(); box()
        res["nccl_parity"] = check("nccl")
        res["nccl_exchange_ms"] = timed(ex)
        res["nccl_step_ms"] = timed(lambda: (ex(), box()))
        capi.lib.vppb_comm_destroy(comm)
    except Exception as e:  # noqa: BLE001
        res["nccl_error"] = repr(e)[:300]
    mpix = nframes * H * W / 1e6
    res["fused_mpix_per_s"] = mpix / (res["fused_ms_per_step"] / 1e3)
    if rank == 0:
        print(json.dumps(res))
    tiles.close_neighbour_tiles(opened)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
