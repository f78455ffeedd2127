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
