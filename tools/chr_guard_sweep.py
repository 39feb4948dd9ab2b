#!/usr/bin/env python
"""How loose may the guard of the factored batched-points route be?  For the bench's parameter points (S = 200, 500 taxa): lnL of
every point through the table route (reference semantics: per-entry clamp) against the factored route at several guard
tolerances; prints the fraction of points on the table route and the worst relative lnL difference among the factored ones.

  python tools/chr_guard_sweep.py [npoints]      (one subprocess per tolerance: the tolerance is read once per process)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from bpp_phyl_b200 import capi, synth
npts = int(sys.argv[2])
S = 200
rng = np.random.default_rng(20260105)
tree = synth.random_tree(500, rng, mean_brlen=0.02, rooted=True)
pts = synth.chromosome_points(S, npts, seed=20260105, well_conditioned_only=True)
mds = [synth.chromosome_model_desc(es) for es in pts]
P0, _, _ = capi.pt_batch(mds[0], tree.brlen, capi.WANT_P)
codes = synth.simulate_single_character(tree, P0, root_state=23, seed=20260105)
e = capi.Engine(S, 1, 1, tree.child_off, tree.children, tree.root, np.eye(S), n_points=npts, n_models=npts, flags=capi.FLAG_WEIGHTED_ROOT)
e.set_all_tip_codes(codes); e.set_pattern_weights(np.ones(1, np.uint32)); e.set_rates(np.ones(1), np.ones(1))
for k in range(npts):
    e.set_model(k, mds[k]); e.set_branch_lengths(k, tree.brlen)
lnl = e.eval(1)[0]
st = e.stats()
print(json.dumps({"lnl": lnl.tolist(), "table_points": int(st["table_points"])}))
"""


def run(npts, env):
    out = subprocess.run([sys.executable, "-c", CHILD, ROOT, str(npts)], capture_output=True, text=True, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def main():
    import numpy as np
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    ref = np.array(run(npts, {"BPPGPU_POINTS_FACTORED": "0"})["lnl"])
    res = {}
    for tol in ("1e-12", "1e-10", "1e-9", "1e-8", "1e-7", "1e-6", "1e-4"):
        r = run(npts, {"BPPGPU_POINTS_GUARD_TOL": tol})
        rel = np.abs(np.array(r["lnl"]) - ref) / np.abs(ref)
        res[tol] = {"table_points": r["table_points"], "max_rel_diff": float(rel.max()), "n_above_1e-9": int((rel > 1e-9).sum())}
        print(tol, res[tol], flush=True)
    print(json.dumps({"npoints": npts, "sweep": res}))


if __name__ == "__main__":
    main()
