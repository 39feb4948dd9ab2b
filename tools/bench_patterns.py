"""Times the site-pattern compression (SitePatterns.cpp:52-106) of one synthetic alignment on the host routine
(bppgpu_site_patterns: std::sort + memcmp, one core -- what the reference does) and on the device routine
(bppgpu_site_patterns_device, H2D copy of the columns and D2H copy of every output included), checks that the
outputs are identical and prints one JSON line.   python tools/bench_patterns.py [n_sites] [n_taxa]"""
import json
import pathlib
import sys
import time

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from bpp_phyl_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
taxa = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rng = np.random.default_rng(20260104)
# columns that share long prefixes, like a real alignment: one ancestral character per site, 4 % of the cells mutated,
# and a pool of sites 4x smaller than the alignment so that most patterns repeat
pool = max(1, n // 4)
base = rng.integers(4, size=pool, dtype=np.uint8)
cols = np.repeat(base[:, None], taxa, axis=1)
mut = rng.random((pool, taxa), dtype=np.float32) < 0.04
cols[mut] = rng.integers(4, size=int(mut.sum()), dtype=np.uint8)
cols = np.frombuffer(b"ACGT", np.uint8)[cols][rng.integers(pool, size=n)]
cols = np.ascontiguousarray(cols)

capi.site_patterns_device(cols[:min(n, 100_000)])      # context, module load and the pinned staging pool (allocated on the first
                                                       # large copy, once per process) stay outside the timed region
t0 = time.perf_counter()
ps_d, w_d, ix_d, tips = capi.site_patterns_device(cols)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
ps_h, w_h, ix_h = capi.site_patterns(cols)
t_host = time.perf_counter() - t0
same = bool(np.array_equal(ps_d, ps_h) and np.array_equal(w_d, w_h) and np.array_equal(ix_d, ix_h)
            and np.array_equal(tips, cols[ps_h].T))

# the same call on PINNED host buffers (one direct DMA each way) and on DEVICE-resident buffers (nothing crosses PCIe): what a
# caller that keeps its alignment in cudaHostAlloc memory / on the GPU gets (include/bppgpu.h: bppgpu_site_patterns_device)
import torch  # noqa: E402  (device memory and pinned allocations only)


def run_raw(make):
    c = make(torch.from_numpy(cols))
    ps = make(torch.empty(n, dtype=torch.int64))
    wt = make(torch.empty(n, dtype=torch.int32))
    ix = make(torch.empty(n, dtype=torch.int64))
    tp = make(torch.empty(n * taxa, dtype=torch.uint8))
    torch.cuda.synchronize()
    best, k = 1e30, 0
    for _ in range(3):
        t0 = time.perf_counter()
        k = capi.site_patterns_device_raw(c.data_ptr(), n, taxa, ps.data_ptr(), wt.data_ptr(), ix.data_ptr(), tp.data_ptr())
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    ok = bool(k == len(ps_h) and np.array_equal(ps[:k].cpu().numpy(), ps_h) and np.array_equal(wt[:k].cpu().numpy().view(np.uint32), w_h)
              and np.array_equal(ix.cpu().numpy(), ix_h) and np.array_equal(tp[:k * taxa].cpu().numpy().reshape(taxa, k), tips))
    return best, ok


t_pin, ok_pin = run_raw(lambda t: t.pin_memory())
t_res, ok_res = run_raw(lambda t: t.cuda())
print(json.dumps({"workload": "site patterns %d sites x %d taxa" % (n, taxa), "n_patterns": int(len(ps_h)),
                  "host_s": round(t_host, 4), "device_s_incl_copies": round(t_dev, 4),
                  "device_s_pinned_buffers": round(t_pin, 4), "device_s_resident_buffers": round(t_res, 4),
                  "speedup_pageable": round(t_host / t_dev, 2), "speedup_pinned": round(t_host / t_pin, 2),
                  "speedup_resident": round(t_host / t_res, 2),
                  "alignment_bytes": int(cols.nbytes), "identical": bool(same and ok_pin and ok_res)}))
