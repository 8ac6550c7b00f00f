"""N > 1 host logic on CPU: world_size-2 gloo processes shard the synthetic ensemble exactly like the
GPU ranks of bench.py do, and the time/count reductions behave (MAX / SUM).  Also checks the oracle gives
shard-invariant results (what each GPU rank computes does not depend on how the ensemble was split)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS
from ivp_b200.dist import reduce_time_and_count, shard_range, weak_offset


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    prob, y0, par, t0, tf = synth.ensemble("vdp", n_per_rank, offset=weak_offset(n_per_rank, rank))
    o = pyoracle.solve_batch(PROBLEMS[prob], t0, 5.0, y0, par, Options(method=Method.DOP853, rtol=1e-8, atol=1e-8))
    ms, total = reduce_time_and_count(10.0 * (rank + 1), float(o.naccpt.sum()), use_dist=True)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), y0=y0, yf=o.y_final, acc=o.naccpt, ms=ms, total=total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction(tmp_path):
    world, n_per = 2, 96
    mp.spawn(_worker, args=(world, _free_port(), n_per, str(tmp_path)), nprocs=world, join=True)
    from oracle import pyoracle
    prob, y0, par, t0, tf = synth.ensemble("vdp", world * n_per)
    full = pyoracle.solve_batch(PROBLEMS[prob], t0, 5.0, y0, par, Options(method=Method.DOP853, rtol=1e-8, atol=1e-8))
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["y0"] for p in parts]), y0)          # shards tile the ensemble
    assert np.array_equal(np.concatenate([p["yf"] for p in parts]), full.y_final)  # shard-invariant results
    for p in parts:
        assert float(p["ms"]) == 20.0                                             # MAX over ranks
        assert float(p["total"]) == float(full.naccpt.sum())                      # SUM over ranks


def test_static_split_covers_everything():
    for N in (0, 1, 7, 1000, 1 << 20):
        for G in (1, 2, 3, 4, 8):
            r = [shard_range(N, g, G) for g in range(G)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
