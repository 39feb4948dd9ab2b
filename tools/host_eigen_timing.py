"""Times the shim's host-side model update (ChromosomeSubstitutionModel::updateMatrices through bppgpu_host_model) at S = 200:
ms per model on the calling thread, with the shim's internal worker threads (BPPGPU_SHIM_THREADS, default min(cores, 16))."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from bpp_phyl_b200 import capi
capi.lib()
rng = np.random.default_rng(1)
P = [(rng.uniform(0, 2), rng.uniform(0, 2), rng.uniform(0, 1), rng.uniform(0, 1)) for _ in range(8)]
capi.host_model("Chromosome", 1, 200, *P[0])
t = time.perf_counter()
for g in P:
    capi.host_model("Chromosome", 1, 200, *g)
print("ms per model", 1e3 * (time.perf_counter() - t) / len(P))
