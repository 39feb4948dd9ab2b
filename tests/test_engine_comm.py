"""SURVEY 8e through the C ABI: pattern shards over several engines.

  * bppgpu_eval_multi       one process, one engine per device (both shards on device 0 when the box has one GPU);
  * bppgpu_comm_init        NCCL inside the engine: a one-rank job on any box, and a real two-rank job (two processes, two
                            GPUs, ids exchanged through a file) when the box has two devices;
  * bppgpu_eval_status      the numeric status of an asynchronous evaluation.
Every form must return the unsharded evaluation's lnL, d1 and d2."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import cases
from oracle import ref_models as rm

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(c, lo, hi):
    sub = cases.Case()
    sub.__dict__.update(c.__dict__)
    sub.N = hi - lo
    sub.weights = c.weights[lo:hi]
    sub.codes_by_leaf = {k: np.ascontiguousarray(v[lo:hi]) for k, v in c.codes_by_leaf.items()}
    return sub


def _case(kind):
    if kind == "dna":
        r, p = rm.gamma_rates(4, 0.5)
        return cases.make_case(40, 900, rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25)), r, p, seed=91, ambiguity=0.02)
    r, p = rm.gamma_rates(4, 0.7)
    return cases.make_case(16, 300, rm.lg08(), r, p, seed=92)


@pytest.mark.parametrize("kind,flags,want", [("dna", 0, 1), ("dna", 1, 7), ("protein", 1, 7)])
def test_eval_multi_over_pattern_shards_equals_the_unsharded_evaluation(kind, flags, want):
    from bpp_phyl_b200 import capi, shard
    import torch
    c = _case(kind)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    ndev = torch.cuda.device_count()
    world = 3
    engines = [cases.make_engine(_shard(c, *shard.shard_range(c.N, g, world)), flags=flags, device=g % ndev) for g in range(world)]
    try:
        lnl, d1, d2 = capi.eval_multi(engines, want)
    finally:
        for e in engines:
            e.close()
    assert abs(lnl[0] - res.lnl) <= 1e-9 * abs(res.lnl)
    if want & 6:
        nb = c.flat.n_nodes - 1
        np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)


def test_one_rank_nccl_job_and_async_status():
    from bpp_phyl_b200 import capi
    import torch
    c = _case("dna")
    res = cases.oracle_eval(c)
    with cases.make_engine(c) as e:
        e.comm_init(0, 1, capi.comm_unique_id())
        lnl, _, _ = e.eval(1)
        assert abs(lnl[0] - res.lnl) <= 1e-9 * abs(res.lnl)
        out = torch.zeros(1 + 2 * c.flat.n_nodes, dtype=torch.float64, device="cuda:0")
        st = torch.cuda.Stream()
        e.eval_device(1, out.data_ptr(), st.cuda_stream)
        assert e.eval_status() == 0
        assert abs(float(out[0]) - res.lnl) <= 1e-9 * abs(res.lnl)
        e.comm_finalize()
        assert abs(e.eval(1)[0][0] - res.lnl) <= 1e-9 * abs(res.lnl)
    with pytest.raises(capi.BppGpuError):
        with cases.make_engine(c) as e:
            bad = next(iter(c.codes_by_leaf.values())).copy()
            bad[3] = c.table.shape[0]                       # one code past the table
            e.set_tip_codes(next(iter(c.codes_by_leaf.keys())), bad)


_WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
rank, world, idfile = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
import cases
from bpp_phyl_b200 import capi, shard
from oracle import ref_models as rm
from test_engine_comm import _case, _shard
for kind, flags, want, wroot in (("dna", 0, 1, False), ("protein", 1, 7, False), ("chr", 0, 1, True)):
    if kind == "chr":
        m = rm.chromosome(1, 24, gain=0.7, loss=0.4, dupl=0.2, demi=0.1)
        c = cases.make_case(9, 40, m, np.ones(1), np.ones(1), seed=5, rooted=True, mean_brlen=0.3, compress=False)
        flags |= capi.FLAG_WEIGHTED_ROOT
    else:
        c = _case(kind)
    sub = _shard(c, *shard.shard_range(c.N, rank, world))
    e = cases.make_engine(sub, flags=flags, device=rank)
    if rank == 0:
        open(idfile + ".tmp", "wb").write(capi.comm_unique_id()); os.replace(idfile + ".tmp", idfile + kind)
    while not os.path.exists(idfile + kind):
        time.sleep(0.05)
    e.comm_init(rank, world, open(idfile + kind, "rb").read())
    lnl, d1, d2 = e.eval(want)
    res = cases.oracle_eval(c, want_d1=bool(want & 2), want_d2=bool(want & 4), weighted_root=wroot)
    assert abs(lnl[0] - res.lnl) <= 1e-9 * abs(res.lnl), (kind, rank, lnl[0], res.lnl)
    if want & 6:
        nb = c.flat.n_nodes - 1
        np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)
    e.close()
print("rank", rank, "ok")
"""


def test_two_rank_nccl_job_when_the_box_has_two_gpus():
    """Two processes, two GPUs, NCCL inside the engines (incl. the weighted-root record exchange).  On a one-GPU box NCCL
    refuses two ranks on one device, so the same shards go through bppgpu_eval_multi above and this test only checks the
    one-device precondition."""
    import torch
    if torch.cuda.device_count() < 2:
        assert torch.cuda.device_count() == 1
        return
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "id_")
        procs = [subprocess.Popen([sys.executable, "-c", _WORKER, ROOT, str(r), "2", idfile], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
