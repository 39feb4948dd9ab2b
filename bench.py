#!/usr/bin/env python
"""bench.py -- the hot path on the BASELINE.json workload, one JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one full log-likelihood evaluation of the workload (K1 P(t) tables for every branch x class,
K2b tip tables, K2+K3 pruning + root reduction [+ K4/K5 derivatives], and for N > 1 the NCCL all-reduce of the
per-shard scalars).  ``value`` = CLV updates per second (one update = one (node, pattern, class, state) element of
an internal node's conditional likelihood vector), whole job, inputs resident in HBM.
``e2e`` = the same through the C-ABI call sequence of one optimiser step with HOST buffers: branch lengths and the
model eigensystem go host->device, log L (and derivatives) come back device->host, inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's banner ("NCCL version ...") and any debug output on stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

WORKLOADS = {
    # BASELINE.json configs[1]: GTR+G4 DNA, 1024 taxa x 1M patterns
    "dna_gtr_g4_1024x1M": dict(S=4, C=4, taxa=1024, patterns=1_000_000, alpha=0.5, derivs=False, seed=20260102),
    # configs[2]: 20 states + G4, 500 taxa x 200k patterns with d1/d2
    "protein_g4_500x200k_d2": dict(S=20, C=4, taxa=500, patterns=200_000, alpha=0.7, derivs=True, seed=20260103),
    # configs[2] without the derivatives (value-only protein evaluation)
    "protein_g4_500x200k": dict(S=20, C=4, taxa=500, patterns=200_000, alpha=0.7, derivs=False, seed=20260103),
    # layout experiment: the same CLV bytes as protein_g4_500x200k with one class (rows contiguous)
    "protein_c1_500x800k": dict(S=20, C=1, taxa=500, patterns=800_000, alpha=None, derivs=False, seed=20260103),
    # configs[3]: 64-state codon (61 sense), 200 taxa x 100k patterns, C = 1
    "codon_200x100k": dict(S=64, C=1, taxa=200, patterns=100_000, alpha=None, derivs=False, seed=20260104),
    # configs[4]: ChromEvol-style chromosome-number model, 200 states, 500 taxa, one character, 4096 parameter points
    "chromosome_500x4096pts": dict(S=200, C=1, taxa=500, patterns=1, points=4096, alpha=None, derivs=False, seed=20260105),
}
DEFAULT = "dna_gtr_g4_1024x1M"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(tree, N, C, S):
    """per CLV update at a node with k sons of which k_t are tips: 8*(1 + k - k_t) bytes."""
    tot = 0
    for n in range(tree.nn):
        sons = tree.sons(n)
        if len(sons) == 0:
            continue
        k_int = int((~tree.is_leaf[sons]).sum())
        tot += 8 * (1 + k_int)
    return tot * N * C * S


def model_for(w, rng):
    from bpp_phyl_b200 import synth
    if w["S"] == 4:
        return synth.gtr()
    return synth.random_reversible(w["S"], rng)


def build_inputs(w, rank, world, device):
    """Tree + model + this rank's contiguous pattern shard (strong scaling: the workload's patterns are split)."""
    from bpp_phyl_b200 import synth
    rng = np.random.default_rng(w["seed"])
    tree = synth.random_tree(w["taxa"], rng, mean_brlen=0.05)
    es = model_for(w, rng)
    rates, probs = synth.gamma_rates(w["C"], w["alpha"]) if w["C"] > 1 else (np.ones(1), np.ones(1))
    N = w["patterns"]
    from bpp_phyl_b200.shard import shard_range
    lo, hi = shard_range(N, rank, world)
    t0 = time.time()
    nblk = 16
    if N % nblk == 0 and nblk % world == 0 and N >= nblk:
        # world-size invariant data: the job's patterns are 16 independently seeded blocks and a rank simulates the blocks
        # of its range, so strong-scaling runs at 1, 2, 4, 8 GPUs evaluate the SAME alignment (equal lnL is then a full-size
        # check of the sharding)
        B = N // nblk
        b0 = lo // B
        parts = [synth.simulate_tip_codes(tree, es, rates, B, seed=w["seed"] + 7919 * (b0 + k + 1), device=device)
                 for k in range((hi - lo) // B)]
        codes = np.concatenate(parts, axis=1) if len(parts) > 1 else parts[0]
    else:
        codes = synth.simulate_tip_codes(tree, es, rates, hi - lo, seed=w["seed"] + 7919 * rank, device=device)
    log("[rank %d] simulated %d patterns x %d tips in %.1fs" % (rank, hi - lo, tree.n_leaves, time.time() - t0))
    return tree, es, rates, probs, codes


def make_engine(w, tree, es, rates, probs, codes, dev_index, flags):
    from bpp_phyl_b200 import capi, synth
    S = w["S"]
    e = capi.Engine(S, w["C"], codes.shape[1], tree.child_off, tree.children, tree.root, np.eye(S), device=dev_index,
                    flags=flags)
    e.set_all_tip_codes(codes)
    e.set_pattern_weights(np.ones(codes.shape[1], np.uint32))
    e.set_rates(rates, probs)
    md = synth.model_desc(es)
    e.set_model(0, md)
    e.set_branch_lengths(0, tree.brlen)
    e.set_root_freqs(0, es["pi"])
    return e, md


def cpu_leg(w, tree, es, rates, probs, codes, threads, target_seconds=12.0, per_thread=512):
    """The reference's CPU algorithm (oracle/ref_cpu.cpp, a port: the reference cannot be built here) on a bounded
    sample of the same workload."""
    from oracle import ref_cpu
    ref_cpu.build()
    n = min(codes.shape[1], per_thread * threads)
    sub = np.ascontiguousarray(codes[:, :n])
    want = 7 if w["derivs"] else 1
    args = dict(S=w["S"], Ccat=w["C"], N=n, child_off=tree.child_off, children=tree.children, root=tree.root, codes=sub,
                code_table=np.eye(w["S"]), weights=np.ones(n, np.uint32), rates=rates, probs=probs, V=es["V"],
                Vinv=es["Vinv"], ev=es["ev"], model_rate=1.0, brlen=tree.brlen, rootfreq=es["pi"], scaled=True, want=want,
                nthreads=threads)
    r = ref_cpu.eval_raw(reps=1, **args)
    reps = int(max(1, min(20, target_seconds / max(r["seconds"], 1e-3))))
    if reps > 1:
        r = ref_cpu.eval_raw(reps=reps, **args)
    upd = tree.n_internal * n * w["C"] * w["S"]
    return {"value": upd / r["seconds"], "unit": "CLV updates/s", "cores": threads, "kind": "port",
            "sample": "%d of %d patterns, full tree, best of %d evals (%.2f s each), scaled arithmetic" %
                      (n, w["patterns"], reps, r["seconds"]),
            "evals_per_s_full_size_extrapolated": (upd / r["seconds"]) / (tree.n_internal * w["patterns"] * w["C"] * w["S"]),
            "lnl_sample": r["lnl"]}, sub


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu summaries (profiles/);
    None when no capture of that kernel on this workload is committed."""
    import csv
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.csv"))):
        rd = wr = None
        try:
            for row in csv.reader(open(f)):
                if len(row) == 4 and kernel.split("_kernel")[0] in row[0]:
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[2])
                    if row[1] == "dram__bytes_read.sum" and scale and rd is None:
                        rd = float(row[3]) * scale
                    if row[1] == "dram__bytes_write.sum" and scale and wr is None:
                        wr = float(row[3]) * scale
        except OSError:
            continue
        if rd is not None and wr is not None and kernel == "walk4_kernel":
            best = rd + wr          # the walk4 captures are of the default (1M-pattern) workload
    return best


def ncu_traffic_per_eval(workload, kernel):
    """DRAM bytes of all launches of `kernel` in one evaluation of `workload` (profiles/r1e_traffic.json, written by
    tools/traffic_from_launches.py from the committed ncu launch list), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1e_traffic.json")))
        return t[workload][kernel]["dram_bytes_per_eval"]
    except (OSError, ValueError, KeyError):
        return None


def family_dmma_flops(tree, rows):
    """FP64 tensor-core flops the per-father derivative kernel issues (S = 20, d1 + d2): per 8 rows and class, 15 DMMAs
    (m8n8k4, 512 flop) for the father contraction of every non-root father and 40 per internal son (P | dP | d2P stacked)."""
    dmma = 0
    for n in range(tree.nn):
        sons = tree.sons(n)
        if not len(sons):
            continue
        dmma += (15 if n != tree.root else 0) + 40 * sum(1 for s in sons if len(tree.sons(s)))
    return dmma * 512.0 * rows / 8


def flops_per_eval(tree, rows, S):
    """pruning contraction flops: 2 S^2 per (row, internal son) + S per extra son"""
    tot = 0
    for n in range(tree.nn):
        sons = tree.sons(n)
        if len(sons) == 0:
            continue
        k_int = int((~tree.is_leaf[sons]).sum())
        tot += k_int * 2 * S * S + (len(sons) - 1) * S
    return tot * rows


def run_points(a, w, rank, world, local, K, W, metric, config):
    """configs[4]: many parameter points on one character.  Points are an independent batch axis: each rank takes a
    contiguous block of them, no collective."""
    import torch
    import torch.distributed as dist
    from bpp_phyl_b200 import capi, synth
    from bpp_phyl_b200.shard import shard_range
    device = "cuda:%d" % local
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    S = w["S"]
    npts_total = (a.points or w["points"]) * (world if a.scaling == "weak" else 1)
    lo, hi = shard_range(npts_total, rank, world)
    npts = hi - lo
    rng = np.random.default_rng(w["seed"])
    tree = synth.random_tree(w["taxa"], rng, mean_brlen=0.02, rooted=True)
    t0 = time.time()
    pts = synth.chromosome_points(S, npts, seed=w["seed"] + 1000 * rank)
    log("[rank %d] %d parameter points (host eigendecompositions) in %.1fs" % (rank, npts, time.time() - t0))
    mds = [synth.chromosome_model_desc(es) for es in pts]
    P0, _, _ = capi.pt_batch(mds[0], tree.brlen, capi.WANT_P, device=local)
    codes = synth.simulate_single_character(tree, P0, root_state=23, seed=w["seed"])
    e = capi.Engine(S, 1, 1, tree.child_off, tree.children, tree.root, np.eye(S), n_points=npts, n_models=npts, device=local,
                    flags=capi.FLAG_WEIGHTED_ROOT)
    e.set_all_tip_codes(codes)
    e.set_pattern_weights(np.ones(1, np.uint32))
    e.set_rates(np.ones(1), np.ones(1))
    for k in range(npts):
        e.set_model(k, mds[k])
        e.set_branch_lengths(k, tree.brlen)
    nn = tree.nn
    out = torch.zeros(npts * (1 + 2 * nn), dtype=torch.float64, device=device)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log("[rank %d] engine ready, path %d" % (rank, e.stats()["path"]))
    for _ in range(W):
        e.eval_device(1, out.data_ptr(), stream.cuda_stream)
    barrier()
    e.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        e.eval_device(1, out.data_ptr(), stream.cuda_stream)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    st = e.stats()
    log("[rank %d] timed region done: %.3f ms/step, P(t) %.3f ms, pruning %.3f ms" % (rank, ms / K, st["pt_ms_sum"] / K, st["prune_ms_sum"] / max(1, st["prune_count"])))
    lnl0 = float(out[0].item())
    # e2e: every point's model (host eigensystem) and branch lengths go host -> device, log L of every point comes back
    h2d = npts * (4 * S * S * 8 + 2 * S * 8 + nn * 8)
    d2h = npts * 8

    def step_e2e():
        for k in range(npts):
            e.set_model(k, mds[k])
            e.set_branch_lengths(k, tree.brlen)
        return e.eval(1)[0]

    n_e2e = 0 if a.profile else max(1, min(K, 2))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        lnl_host = step_e2e()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / max(1, n_e2e)
    tms = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    clocks = sampler.stop() if rank == 0 else None
    upd_step = tree.n_internal * npts_total * S          # CLV updates of the whole job per step
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads(open(os.path.join(ROOT, "profiles", "r1_fp64_peaks.json")).readline())
        except (OSError, ValueError, AttributeError):
            pass
        dmma_peak = peaks.get("dmma_m8n8k4_tflops", 37.1)
        n_mat = npts * (nn - 1)
        flops = n_mat * (2.0 * S ** 3 + S * S)
        pt_ms = st["pt_ms_sum"] / max(1, K)              # all chunks of one evaluation
        ach = flops / (pt_ms * 1e-3) / 1e12 if pt_ms > 0 else None
        line = {"metric": metric, "value": upd_step * K / (ms * 1e-3), "unit": "CLV updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic (one chromosome count per taxon simulated down a random rooted tree; parameter points "
                                        "drawn uniformly, numerically defective generators redrawn)",
                "config": dict(config, points=npts_total, sharding="points/%d (replicas, no collective)" % world),
                "logl_evals_per_s": npts_total * K / (ms * 1e-3), "lnl_point0": lnl0,
                "e2e": {"value": upd_step / (e2e_ms * 1e-3) if e2e_ms > 0 else None, "unit": "CLV updates/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                        "logl_evals_per_s": npts_total / (e2e_ms * 1e-3) if e2e_ms > 0 else None,
                        "call": "bppgpu_set_model + bppgpu_set_branch_lengths per point, one bppgpu_eval (host buffers)"},
                "gpu_launches": int(st["kernel_launches"]) * K, "launches_per_step": int(st["kernel_launches"]),
                "roofline": {"bound": "tensor", "kernel": "pt_dmma_kernel", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                             "frac": ach / dmma_peak if ach else None, "traffic": None, "kernel_ms": pt_ms,
                             "flops_per_eval": flops,
                             "peak_source": "FP64 mma.sync m8n8k4 measured with tools/fp64_peak.cu (profiles/r1_fp64_peaks.json); "
                                            "MEASURED_PEAKS.json has no FP64 figure",
                             "pruning_ms": st["prune_ms_sum"] / max(1, st["prune_count"])},
                "clocks": clocks, "hbm_resident_bytes": int(st["hbm_bytes_resident"])}
        if world == 1 and not a.no_cpu:
            # the reference's CPU algorithm on a few points, one point per host thread
            from concurrent.futures import ThreadPoolExecutor
            from oracle import ref_cpu
            ref_cpu.build()
            threads = min(os.cpu_count() or 1, npts, 16)

            def one(k):
                es = pts[k]
                return ref_cpu.eval_raw(S, 1, 1, tree.child_off, tree.children, tree.root, codes, np.eye(S), np.ones(1, np.uint32),
                                        np.ones(1), np.ones(1), es["V"], es["Vinv"], es["ev"], 1.0, tree.brlen, es["pi"], scaled=True,
                                        want=1, nthreads=1, reps=1, ev_im=es["ev_im"], chr_clamp=True, weighted_root=True)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:
                res = list(ex.map(one, range(threads)))
            dt = time.perf_counter() - t0
            rel = max(abs(res[k]["lnl"] - lnl_host[k]) / abs(res[k]["lnl"]) for k in range(threads))
            line["cpu_baseline"] = {"value": tree.n_internal * threads * S / dt, "unit": "CLV updates/s", "cores": threads, "kind": "port",
                                    "sample": "%d of %d points, full tree, one point per thread, %.1f s" % (threads, npts_total, dt),
                                    "logl_evals_per_s": threads / dt, "rel_diff_vs_gpu": rel}
        print(json.dumps(line), flush=True)
        log("[rank 0] JSON line printed")
    e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--patterns", type=int, default=0, help="override the workload's pattern count (debug)")
    ap.add_argument("--points", type=int, default=0, help="override the number of parameter points (chromosome workload)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU gets the workload's pattern count (patterns are an independent axis, no data-path "
                         "collective); strong = the workload's patterns are split across the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + 1 timed step, no e2e / CPU legs (for ncu only)")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    if a.patterns:
        w["patterns"] = a.patterns
    W = max(a.warmup, 3) if a.impl == "native" else a.warmup
    K = a.steps
    if a.profile:
        W, K, a.no_cpu = 1, 1, True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    per_gpu_patterns = w["patterns"]
    if a.scaling == "weak" and "points" not in w:
        w["patterns"] = per_gpu_patterns * world        # whole job; each rank still owns a contiguous 1/world block
    metric = "CLV updates/s"
    config = {"workload": a.workload, "states": w["S"], "rate_classes": w["C"], "taxa": w["taxa"],
              "patterns": w["patterns"], "patterns_per_gpu": w["patterns"] // world, "derivatives": w["derivs"],
              "sharding": "contiguous pattern blocks, one per GPU; all-reduce of (lnL, d1, d2) only",
              "l2": "inputs larger than L2 (tip codes %.0f MB per rank; CLVs never re-read from a previous step)" %
                    (w["taxa"] * w["patterns"] / world / 1e6)}

    if "points" in w and a.impl == "native":
        import __graft_entry__ as g
        from bpp_phyl_b200 import capi
        if not capi.LIB_PATH.exists():
            g.build_lib()
        return run_points(a, w, rank, world, local, K, W, metric, config)
    if a.impl == "reference":
        if rank != 0:
            return 0
        from bpp_phyl_b200 import synth
        rng = np.random.default_rng(w["seed"])
        tree = synth.random_tree(w["taxa"], rng, mean_brlen=0.05)
        es = model_for(w, rng)
        rates, probs = synth.gamma_rates(w["C"], w["alpha"]) if w["C"] > 1 else (np.ones(1), np.ones(1))
        threads = os.cpu_count() or 1
        n = 512 * threads
        codes = np.random.default_rng(1).integers(0, w["S"], size=(tree.n_leaves, n), dtype=np.uint8)
        from oracle import ref_cpu
        ref_cpu.build()
        want = 7 if w["derivs"] else 1
        args = dict(S=w["S"], Ccat=w["C"], N=n, child_off=tree.child_off, children=tree.children, root=tree.root,
                    codes=codes, code_table=np.eye(w["S"]), weights=np.ones(n, np.uint32), rates=rates, probs=probs,
                    V=es["V"], Vinv=es["Vinv"], ev=es["ev"], model_rate=1.0, brlen=tree.brlen, rootfreq=es["pi"],
                    scaled=True, want=want, nthreads=threads)
        # one call: W + K evaluations on the same likelihood object (setData-style allocation is not part of a step);
        # the C side times every evaluation, the last K are the timed steps
        r_all = ref_cpu.eval_raw(reps=max(1, a.warmup), **args) if a.warmup else None
        r = ref_cpu.eval_raw(reps=K, **args)
        dt = r["total_seconds"]
        upd = tree.n_internal * n * w["C"] * w["S"]
        val = upd * K / dt
        sample = "%d of %d patterns per step (full tree), %d host threads as independent pattern shards" % (n, w["patterns"], threads)
        print(json.dumps({"impl": "reference", "metric": metric, "value": val, "unit": "CLV updates/s", "n_gpus": a.gpus,
                          "steps": K, "warmup": a.warmup, "ms_per_step": 1e3 * dt / K, "higher_is_better": True,
                          "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": val, "unit": "CLV updates/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "CLV updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---------------- native arm ----------------
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from bpp_phyl_b200 import capi, shard
    if not capi.LIB_PATH.exists():
        g.build_lib()
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    flags = capi.FLAG_KEEP_CLVS if w["derivs"] else 0
    want = 7 if w["derivs"] else 1
    tree, es, rates, probs, codes = build_inputs(w, rank, world, device)
    e, md = make_engine(w, tree, es, rates, probs, codes, local, flags)
    nn = tree.nn
    out = torch.zeros(1 + 2 * nn, dtype=torch.float64, device=device)
    # a real (non-NULL) stream: kernels, NCCL and the timing events all go on it
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)

    def step_device():
        e.eval_device(want, out.data_ptr(), stream.cuda_stream)
        shard.combine(out)          # NCCL all-reduce(sum) of (lnL, d1, d2) on the same stream; no-op at N = 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log("[rank %d] engine ready, path %d" % (rank, e.stats()["path"]))
    for _ in range(W):
        step_device()
    barrier()
    log("[rank %d] warm-up done" % rank)
    st0 = e.stats()             # clears the kernel-timing ring
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    log("[rank %d] timed region done: %.3f ms/step" % (rank, ms / K))
    st = e.stats()
    lnl_dev = float(out[0].item())
    tms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())

    # ---- e2e: one optimiser step through the C ABI with host buffers -----------------------------------------
    brl = tree.brlen.copy()
    host_out = torch.empty(1 + 2 * nn, dtype=torch.float64).pin_memory()
    h2d = brl.nbytes + es["V"].nbytes + es["Vinv"].nbytes + es["ev"].nbytes + es["Q"].nbytes
    d2h = 8 * (1 + (2 * nn if w["derivs"] else 0))

    def step_e2e(i):
        brl[0] = tree.brlen[0] * (1.0 + 1e-3 * ((i % 5) - 2))      # a branch-length probe, as an optimiser would send
        e.set_model(0, md)                                          # eigensystem host -> device
        e.set_branch_lengths(0, brl)                                # host -> device
        if world == 1:
            return e.eval(want)[0][0]                               # log L (d1, d2) device -> host
        e.eval_device(want, out.data_ptr(), stream.cuda_stream)
        shard.combine(out)
        host_out.copy_(out, non_blocking=False)
        return float(host_out[0])

    for i in range(0 if a.profile else W):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for i in range(0 if a.profile else K):
        step_e2e(i)
    ev1.record(stream)
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))
    tms = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
    log("[rank %d] e2e region done: %.3f ms/step" % (rank, e2e_ms / K))
    clocks = sampler.stop() if rank == 0 else None
    e.set_branch_lengths(0, tree.brlen)

    upd_step = tree.n_internal * w["patterns"] * w["C"] * w["S"]         # whole job (all shards)
    value = upd_step * K / (ms * 1e-3)
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        n_local = codes.shape[1]
        alg = algorithmic_bytes(tree, n_local, w["C"], w["S"])
        kms = st["prune_ms_sum"] / max(1, st["prune_count"])
        kname = {1: "walk4_kernel", 2: "walkS_kernel", 3: "generic_node_kernel", 4: "dmma_node_kernel"}.get(st["path"], "?")
        fam20 = st["path"] == 4 and w["S"] == 20 and w["C"] <= 4 and os.environ.get("BPPGPU_FAMILY", "1") != "0"
        if fam20 or (st["path"] == 4 and w["S"] == 64 and w["C"] == 1 and os.environ.get("BPPGPU_FAMILY", "1") != "0"):
            kname = "dmma_prune_kernel"
        if st["path"] == 4 and w["S"] >= 32:
            # dense contraction on the FP64 tensor cores (mma.sync DMMA; tcgen05 has no f64 kind)
            try:
                fp = json.loads(open(os.path.join(ROOT, "profiles", "r1_fp64_peaks.json")).readline())
            except (OSError, ValueError):
                fp = {}
            dmma_peak = fp.get("dmma_m8n8k4_tflops", 37.1)
            rows = n_local * w["C"]
            alg_flops = sum((2 * w["S"] * len(tree.sons(n)) + len(tree.sons(n)) - 1) * w["S"] for n in range(tree.nn)
                            if len(tree.sons(n))) * rows                  # every son contracted, tips included
            exe_flops = flops_per_eval(tree, rows, w["S"])                # DMMA flops issued: tip sons are table gathers
            ach = alg_flops / (kms * 1e-3) / 1e12 if kms > 0 else None
            roofline = {"bound": "tensor", "kernel": kname + " (one launch per node; all launches of an evaluation timed together)",
                        "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak if ach else None,
                        "traffic": ncu_traffic_per_eval(a.workload, kname) if (not a.patterns and world == 1) else None,
                        "kernel_ms": kms, "launches_timed": st["prune_count"], "algorithmic_flops_per_eval": alg_flops,
                        "executed_dmma_tflops": exe_flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                        "peak_source": "FP64 mma.sync m8n8k4 measured with tools/fp64_peak.cu (profiles/r1_fp64_peaks.json); "
                                       "MEASURED_PEAKS.json has no FP64 figure",
                        "note": "algorithmic flops = (2 S k + k - 1) per CLV element with k sons (SURVEY 8d, the reference contracts "
                                "tip sons too); tip sons are gathers here, so executed_dmma_tflops is the tensor-pipe figure"}
        else:
            achieved = alg / (kms * 1e-3) / 1e9 if kms > 0 else None
            roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak if achieved else None,
                        "traffic": (ncu_traffic(kname) if a.workload == DEFAULT else ncu_traffic_per_eval(a.workload, kname))
                        if (not a.patterns and world == 1) else None,
                        "kernel_ms": kms,
                        "launches_timed": st["prune_count"], "algorithmic_bytes_per_launch": alg,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                        "note": ("algorithmic bytes = 8*(1+internal sons) per CLV element (SURVEY 8d), the traffic a level-scheduled "
                                 "kernel must move; this kernel keeps CLVs on chip, so frac > 1 is expected and DRAM traffic is "
                                 "the tip codes only (traffic = dram bytes of one launch, ncu, profiles/)") if st["path"] == 1 else
                                "algorithmic bytes = 8*(1+internal sons) per CLV element (SURVEY 8d); one launch per node, all "
                                "launches of an evaluation timed together"}
        if fam20 and w["derivs"]:
            # second kernel family of this workload: the per-father upper + d1/d2 pass, bound by the FP64 tensor pipe
            try:
                fp = json.loads(open(os.path.join(ROOT, "profiles", "r1_fp64_peaks.json")).readline())
            except (OSError, ValueError):
                fp = {}
            dmma_peak = fp.get("dmma_m8n8k4_tflops", 37.1)
            dms = ms / K - kms - st["pt_ms_sum"] / max(1, K)
            fl = family_dmma_flops(tree, n_local * w["C"])
            roofline["derivative_pass"] = {
                "kernel": "dmma_family_kernel (one launch per father)", "bound": "tensor", "ms": dms,
                "executed_dmma_flops": fl, "achieved": fl / (dms * 1e-3) / 1e12 if dms > 0 else None, "peak": dmma_peak,
                "unit": "TFLOP/s", "frac": fl / (dms * 1e-3) / 1e12 / dmma_peak if dms > 0 else None,
                "traffic": ncu_traffic_per_eval(a.workload, "dmma_family_kernel") if (not a.patterns and world == 1) else None,
                "note": "ms = step - pruning - P(t) (event timed); flops = DMMAs issued x 512"}
        line = {"metric": metric, "value": value, "unit": "CLV updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (tips simulated down a random tree under the model; every site kept as a pattern, weight 1)",
                "config": config, "logl_evals_per_s": K / (ms * 1e-3), "lnl": lnl_dev,
                "e2e": {"value": upd_step * K / (e2e_ms * 1e-3), "unit": "CLV updates/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K,
                        "call": "bppgpu_set_model + bppgpu_set_branch_lengths + bppgpu_eval (host buffers)"},
                "gpu_launches": int(st["kernel_launches"]) * K, "launches_per_step": int(st["kernel_launches"]),
                "roofline": roofline, "clocks": clocks,
                "hbm_resident_bytes": int(st["hbm_bytes_resident"])}
        if world == 1 and not a.no_cpu:
            t0 = time.time()
            cb, sub = cpu_leg(w, tree, es, rates, probs, codes, threads=os.cpu_count() or 1)
            # the same sample through the CUDA path: the checker, not the thing measured
            e2, _ = make_engine(w, tree, es, rates, probs, sub, local, 0)
            l2 = e2.eval(1)[0][0]
            e2.close()
            cb["gpu_lnl_same_sample"] = l2
            cb["rel_diff_vs_gpu"] = abs(l2 - cb["lnl_sample"]) / abs(cb["lnl_sample"])
            line["cpu_baseline"] = cb
            log("cpu leg %.1fs" % (time.time() - t0))
        print(json.dumps(line), flush=True)
        log("[rank 0] JSON line printed")
    e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        raise
    sys.stdout.flush()
    sys.exit(rc)
