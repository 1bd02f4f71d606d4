"""CPU, world_size 2 over gloo: the sharded-scan update (fast_limo_b200.dist) — slice bounds, the
packed 96-double exchange format, and the product EKF state machine under a real collective.
The per-shard measurement here comes from the CPU oracle (there is no GPU in this container); on the
GPU box bench.py runs the same driver with the CUDA pass and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_limo_b200 import api, synth
from fast_limo_b200.dist import pack96, shard_bounds, sharded_update


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 131072, 300000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_pack_unpack_roundtrip(flimo_lib):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(12, 12))
    H = A + A.T
    h = rng.normal(size=12)
    r = api.unpack96(pack96(H, h, 17, 3.5, 21))
    assert np.array_equal(r.HTH, H) and np.array_equal(r.HTh, h) and (r.n_rows, r.sum_sq_res, r.n_valid) == (17, 3.5, 21)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    case = synth.make_case("tiny")
    om = O.OracleMap()
    om.add(case.map_pts)
    n = case.scan.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    cfg = O.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20)

    def local_pass(state):
        r = om.match(cfg, state[:14], case.scan[lo:hi])
        ss = float((r["dist"][r["good"]].astype(np.float64) ** 2).sum())
        return torch.from_numpy(pack96(r["HTH"], r["HTh"], r["rows"], ss, r["n_valid"]))

    def all_reduce(t):
        dist.all_reduce(t)
        return t.numpy()

    m = api.Mapper(device=-1)           # host-only handle: filter algebra only
    x, P, passes = sharded_update(m, case.init, synth.default_P0(), 2, 0.0, local_pass, all_reduce)
    out[rank] = (x, P, passes)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_update_two_ranks(oracle, flimo_lib):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (x0, P0, p0), (x1, P1, p1) = out[0], out[1]
    assert p0 == p1 == 3
    assert np.array_equal(x0, x1) and np.array_equal(P0, P1)          # every rank ends in the same state
    case = synth.make_case("tiny")
    om = oracle.OracleMap()
    om.add(case.map_pts)
    xo, Po, tr = om.update(oracle.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20), case.init, synth.default_P0(), 2, 0.0, case.scan)
    assert len(tr) == 3
    assert np.abs(x0 - xo).max() < 1e-10 and np.allclose(P0, Po, rtol=1e-4, atol=1e-11)
