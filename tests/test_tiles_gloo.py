"""world_size-2 (and 3) gloo test of the N>1 host logic: row tiles + one grouped halo exchange.
Each rank owns a row tile of a frame (border rows = halo), swaps edge rows with its neighbours
through vpp_b200.tiles.exchange_halos, runs the ORACLE 5x5 box on its tile and the concatenated
result must equal the oracle on the whole frame (bit-exact)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from tests import oracle as orc
    import importlib.util

    spec = importlib.util.spec_from_file_location("tiles", os.path.join(ROOT, "vpp_b200", "tiles.py"))
    tiles = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tiles)

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = orc.load()
    frame = np.random.default_rng(99).integers(0, 256, (H, W, 3), dtype=np.uint8)  # same on every rank
    r0, r1 = tiles.tile_rows(H, rank, world)
    th = r1 - r0
    tile = orc.HostImage(th, W, "vuchar3", border=2, data=frame[r0:r1], fill_border="mirror")  # mirror everywhere first
    v = tile.view(True)
    # edge rows incl. the (already mirrored) column border = exactly what vppb_halo_pack ships
    send_up, send_dn = torch.from_numpy(v[2:4].copy()).flatten(), torch.from_numpy(v[-4:-2].copy()).flatten()
    recv_up, recv_dn = torch.empty_like(send_up), torch.empty_like(send_dn)
    up, down = tiles.exchange_halos(dist, rank, world, send_up, send_dn, recv_up, recv_dn)
    if up is not None:
        v[0:2] = recv_up.numpy().reshape(2, W + 4, 3)
    if down is not None:
        v[-2:] = recv_dn.numpy().reshape(2, W + 4, 3)
    out = orc.HostImage(th, W, "vuchar3")
    o.vo_box5x5_u8(tile.ptr(), out.ptr(), 3)
    np.save(os.path.join(out_dir, "tile%d.npy" % rank), out.get())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_tiles_with_halo_exchange_match_whole_frame(built, tmp_path, world):
    import torch.multiprocessing as mp

    from tests import oracle as orc

    H, W = 61, 47
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, W, str(tmp_path)), nprocs=world, join=True)
    frame = np.random.default_rng(99).integers(0, 256, (H, W, 3), dtype=np.uint8)
    whole_in = orc.HostImage(H, W, "vuchar3", border=2, data=frame, fill_border="mirror")
    whole_out = orc.HostImage(H, W, "vuchar3")
    orc.load().vo_box5x5_u8(whole_in.ptr(), whole_out.ptr(), 3)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "tile%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, whole_out.get())


def test_tile_bounds():
    import importlib.util

    spec = importlib.util.spec_from_file_location("tiles", os.path.join(ROOT, "vpp_b200", "tiles.py"))
    tiles = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tiles)
    for H in (1080, 4320, 61):
        for world in (1, 2, 3, 4, 8):
            b = [tiles.tile_rows(H, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == H and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    assert tiles.neighbours(0, 1) == (None, None) and tiles.neighbours(0, 2) == (None, 1) and tiles.neighbours(3, 8) == (2, 4)
