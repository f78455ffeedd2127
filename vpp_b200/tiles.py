"""Row-tile sharding of a frame over the ranks of one node and the single grouped halo exchange.

Host logic only (no kernels): tile bounds, neighbour ranks, and one `batch_isend_irecv` that swaps the
edge rows of every frame of a step with both neighbours.  Works with any torch.distributed backend:
NCCL over NVLink on the GPUs (bench.py), gloo on CPU (tests/test_tiles_gloo.py)."""


def tile_rows(nrows, rank, world):
    """Rows [r0, r1) of the frame owned by `rank` (contiguous, floor split as SURVEY §8e)."""
    return (nrows * rank) // world, (nrows * (rank + 1)) // world


def neighbours(rank, world):
    """(up, down) ranks or None at the frame edge (outer tiles keep their mirror border)."""
    return (rank - 1 if rank > 0 else None), (rank + 1 if rank + 1 < world else None)


def halo_ops(dist, rank, world, send_up, send_dn, recv_up, recv_dn):
    """The P2P operations of one grouped exchange (build once per set of staging buffers, run every step)."""
    up, down = neighbours(rank, world)
    ops = []
    if up is not None:
        ops += [dist.P2POp(dist.isend, send_up, up), dist.P2POp(dist.irecv, recv_up, up)]
    if down is not None:
        ops += [dist.P2POp(dist.isend, send_dn, down), dist.P2POp(dist.irecv, recv_dn, down)]
    return ops


def run_halo_ops(dist, ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_halos(dist, rank, world, send_up, send_dn, recv_up, recv_dn):
    """One grouped exchange: my top edge rows go up, my bottom edge rows go down; the neighbours'
    edge rows land in recv_up / recv_dn.  Tensors are flat staging buffers (all frames packed)."""
    run_halo_ops(dist, halo_ops(dist, rank, world, send_up, send_dn, recv_up, recv_dn))
    return neighbours(rank, world)


def exchange_halos_inplace(dist, rank, world, send_up, send_dn, recv_up, recv_dn):
    """Same single grouped exchange, but straight from / into the images: because the edge rows of a pitched
    tile (with their column border) are one contiguous block, the per-frame lists of tensor views
    (top edge rows, bottom edge rows, top border rows, bottom border rows) need no pack / unpack kernels."""
    up, down = neighbours(rank, world)
    ops = []
    if up is not None:
        ops += [dist.P2POp(dist.isend, t, up) for t in send_up] + [dist.P2POp(dist.irecv, t, up) for t in recv_up]
    if down is not None:
        ops += [dist.P2POp(dist.isend, t, down) for t in send_dn] + [dist.P2POp(dist.irecv, t, down) for t in recv_dn]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return up, down


# ---- peer-memory form: every rank maps its neighbours' tiles and the kernels read the halo rows from there -----------
def open_neighbour_tiles(dist, rank, world, images):
    """Exchange CUDA IPC handles of `images` (Image2d that own their allocation, same count and geometry on every rank) and
    map the tiles of rank - 1 and rank + 1.  Returns (ups, downs): ctypes arrays of vppb_img (base == NULL where there is
    no neighbour) to pass to vppb_box5x5_*_tiles, plus the list of opened descriptors to close with close_neighbour_tiles."""
    import ctypes as C

    from . import capi

    mine = []
    for im in images:
        h, off = (C.c_char * 64)(), C.c_int64()
        capi.check(capi.lib.vppb_ipc_export(im.ptr(), h, C.byref(off)))
        mine.append((bytes(h.raw), int(off.value)))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    up, down = neighbours(rank, world)
    n = len(images)
    ups, downs, opened = (capi.VppbImg * n)(), (capi.VppbImg * n)(), []
    for who, arr in ((up, ups), (down, downs)):
        if who is None:
            continue
        for i, (hb, off) in enumerate(everyone[who]):
            buf = (C.c_char * 64).from_buffer_copy(hb)
            capi.check(capi.lib.vppb_ipc_open(buf, off, images[i].ptr(), C.byref(arr[i])))
            opened.append(arr[i])
    return ups, downs, opened


def close_neighbour_tiles(opened):
    import ctypes as C

    from . import capi

    for d in opened:
        capi.lib.vppb_ipc_close(C.byref(d))


def nccl_comm(dist, rank, world):
    """A vppb (NCCL) communicator for vppb_halo_exchange: rank 0 draws the unique id, torch.distributed carries it."""
    import ctypes as C

    from . import capi

    idb = (C.c_char * 128)()
    if rank == 0:
        capi.check(capi.lib.vppb_comm_unique_id(idb))
    box = [bytes(idb.raw)]
    dist.broadcast_object_list(box, src=0)
    idb = (C.c_char * 128).from_buffer_copy(box[0])
    comm = C.c_void_p()
    capi.check(capi.lib.vppb_comm_init(idb, rank, world, C.byref(comm)))
    return comm
