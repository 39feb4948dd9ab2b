"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank evaluates ITS pattern shard with the oracle (standing in for
the per-rank engine), the per-shard (lnL, d1, d2) vectors are combined with bpp_phyl_b200.shard.combine, and the
result must equal the single-process evaluation of the whole alignment."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from bpp_phyl_b200 import shard
from oracle import ref_models as rm


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(12, 90, rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25)), r, p, seed=77)
    lo, hi = shard.shard_range(c.N, rank, world)
    sub = cases.Case()
    sub.__dict__.update(c.__dict__)
    sub.N = hi - lo
    sub.weights = c.weights[lo:hi]
    sub.codes_by_leaf = {k: v[lo:hi] for k, v in c.codes_by_leaf.items()}
    res = cases.oracle_eval(sub, want_d1=True, want_d2=True)
    nn = c.flat.n_nodes
    out = torch.zeros(1 + 2 * nn, dtype=torch.float64)
    out[0] = res.lnl
    out[1:nn] = torch.from_numpy(res.d1)
    out[1 + nn:2 * nn] = torch.from_numpy(res.d2)
    shard.combine(out)
    if rank == 0:
        q.put(out.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(g, world, port, q)) for g in range(world)]
    for p_ in procs:
        p_.start()
    got = q.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(12, 90, rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25)), r, p, seed=77)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    nn = c.flat.n_nodes
    assert abs(got[0] - res.lnl) <= 1e-12 * abs(res.lnl)
    np.testing.assert_allclose(got[1:nn], res.d1, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(got[1 + nn:2 * nn], res.d2, rtol=1e-11, atol=1e-10)
