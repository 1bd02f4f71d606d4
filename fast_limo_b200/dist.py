"""Multi-GPU driver of the iterated update: scan points shard across ranks, the map is replicated.

SURVEY 8(e): query points are independent (Mapper.cpp:68-76), so rank r matches the contiguous slice
[n*r/W, n*(r+1)/W) of the scan against its own map replica and produces the 96-double packed normal
equations of its slice; ONE all-reduce (sum) of those 96 doubles per pass gives every rank the
whole-scan H^T H / H^T h, and every rank then runs the identical host filter algebra, so no pose
broadcast is needed.  The collective is supplied by the caller (NCCL on device tensors in bench.py,
gloo on CPU tensors in the tests).
"""
import numpy as np

from . import api


def shard_bounds(n, rank, world):
    """Contiguous slice of an n-point scan owned by `rank` out of `world`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def pack96(HTH, HTh, n_rows, sum_sq_res, n_valid):
    """Inverse of flimo_unpack96 (used by tests that synthesise per-shard results)."""
    p = np.zeros(96)
    e = 0
    for i in range(12):
        for j in range(i, 12):
            p[e] = HTH[i, j]
            e += 1
    p[78:90] = HTh
    p[90], p[91], p[92] = n_rows, sum_sq_res, n_valid
    return p


def sharded_update(mapper, x0, P0, max_iter, limits, local_pass, all_reduce, R=0.001, D=5.0):
    """esekf::update_iterated_dyn_share_modified with the measurement pass split over ranks.

    local_pass(state26) -> this rank's packed 96 doubles (any array-like the collective accepts)
    all_reduce(packed)  -> numpy array of the 96 sums over all ranks
    Returns (state26, P, passes).
    """
    mapper.ekf_begin(x0, P0, max_iter, limits, R, D)
    passes, done = 0, max_iter < 0
    while not done:
        summed = np.asarray(all_reduce(local_pass(mapper.ekf_state())), dtype=np.float64)
        r = api.unpack96(summed)
        done = mapper.ekf_step(r.HTH, r.HTh, r.n_rows)
        passes += 1
    x, P = mapper.ekf_end()
    return x, P, passes


EXCHANGE_BYTES_PER_RANK = 4096


def attach_exchange(mapper, rank, world, broadcast_object):
    """Create (rank 0) / open a POSIX shared-memory segment for the fused exchange and attach it.

    broadcast_object(obj_or_None) must return rank 0's object on every rank (e.g. a wrapper around
    torch.distributed.broadcast_object_list).  Returns the SharedMemory object (keep it alive; rank 0
    unlinks it at the end)."""
    from multiprocessing import shared_memory
    size = max(4096, world * EXCHANGE_BYTES_PER_RANK)
    size = (size + 4095) // 4096 * 4096
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=size)
        shm.buf[:size] = bytes(size)
        name = broadcast_object(shm.name)
    else:
        name = broadcast_object(None)
        shm = shared_memory.SharedMemory(name=name)
    mapper.exchange_attach(shm.buf, rank, world)
    return shm


def sharded_update_exchange(mapper, x0, P0, max_iter, limits, R=0.001, D=5.0):
    """sharded_update with the measurement pass + exchange done by flimo_match_reduce_exchange."""
    mapper.ekf_begin(x0, P0, max_iter, limits, R, D)
    passes, done = 0, max_iter < 0
    while not done:
        r = mapper.match_exchange(mapper.ekf_state())
        done = mapper.ekf_step(r.HTH, r.HTh, r.n_rows)
        passes += 1
    x, P = mapper.ekf_end()
    return x, P, passes
