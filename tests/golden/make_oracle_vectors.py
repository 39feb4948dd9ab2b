"""Writes tests/golden/oracle_vectors.json: outputs of the oracle (oracle/, the CPU restatement of the reference's
algorithm) on small seeded cases, one per model family of BASELINE.json's configs.

These are REGRESSION fixtures, not reference outputs: the reference cannot be built or imported in this container
(needs bpp-core and the forked bpp-seq), and its own tests hold no known-answer value for LG08, YN98 or the
Chromosome model, so parity for those families is pinned only through the DNA golden values of reference_vectors.json
plus the structural checks of tests/test_oracle_golden.py.  The GPU tests compare the CUDA path with these numbers
without running the oracle; the CPU suite checks that the oracle still reproduces them.

    python tests/golden/make_oracle_vectors.py
"""
import json
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

import cases  # noqa: E402
from oracle import ref_models as rm  # noqa: E402


def build(spec):
    """spec -> (Case, oracle kwargs); shared with the tests so both sides construct identical inputs."""
    fam = spec["family"]
    if fam == "gtr":
        m = rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25))
    elif fam == "lg08":
        m = rm.lg08()
    elif fam == "yn98":
        m = rm.yn98(2.0, 0.3)
    elif fam == "chromosome":
        m = rm.chromosome(1, spec["max_chr"], gain=0.7, loss=0.4, dupl=0.2, demi=rm.DEMI_EQUAL_DUPL)
    else:
        raise ValueError(fam)
    r, p = rm.gamma_rates(spec["ncat"], spec["alpha"]) if spec["ncat"] > 1 else rm.constant_rate()
    c = cases.make_case(spec["taxa"], spec["sites"], m, r, p, seed=spec["seed"], rooted=spec.get("rooted", False),
                        mean_brlen=spec.get("mean_brlen", 0.05), ambiguity=spec.get("ambiguity", 0.0), compress=spec.get("compress", True))
    return c, dict(want_d1=spec["derivs"], want_d2=spec["derivs"], nh_form=spec.get("nh_form", False),
                   weighted_root=spec.get("weighted_root", False))


SPECS = [
    {"id": "gtr_g4_30x200", "family": "gtr", "ncat": 4, "alpha": 0.5, "taxa": 30, "sites": 200, "seed": 20260101, "derivs": True},
    {"id": "lg08_g4_16x80", "family": "lg08", "ncat": 4, "alpha": 0.7, "taxa": 16, "sites": 80, "seed": 20260102, "derivs": True,
     "ambiguity": 0.03, "mean_brlen": 0.1},
    {"id": "lg08_g4_rooted_nh_9x37", "family": "lg08", "ncat": 4, "alpha": 0.7, "taxa": 9, "sites": 37, "seed": 20260103, "derivs": True,
     "rooted": True, "nh_form": True, "compress": False},
    {"id": "yn98_c1_8x40", "family": "yn98", "ncat": 1, "alpha": None, "taxa": 8, "sites": 40, "seed": 20260104, "derivs": True},
    {"id": "chromosome_30_weighted_root_12x1", "family": "chromosome", "max_chr": 30, "ncat": 1, "alpha": None, "taxa": 12, "sites": 1,
     "seed": 20260105, "derivs": False, "rooted": True, "weighted_root": True, "mean_brlen": 0.3, "compress": False},
    {"id": "gtr_g4_underflow_700x12", "family": "gtr", "ncat": 4, "alpha": 0.5, "taxa": 700, "sites": 12, "seed": 20260106, "derivs": False,
     "mean_brlen": 0.5, "compress": False},
]


def main():
    out = {"_comment": "oracle outputs (regression fixtures; see make_oracle_vectors.py). minus-signs follow the reference: d1, d2 are "
                       "derivatives of -lnL per BrLen<i>.", "vectors": []}
    for spec in SPECS:
        c, kw = build(spec)
        res = cases.oracle_eval(c, **kw)
        v = {"spec": spec, "n_patterns": int(c.N), "lnl": float(res.lnl), "site_lnl": [float(x) for x in res.site_lnl[:16]]}
        if spec["derivs"]:
            v["d1"] = [float(x) for x in res.d1]
            v["d2"] = [float(x) for x in res.d2]
        if spec.get("weighted_root"):
            v["root_freqs"] = [float(x) for x in res.root_freqs]
        out["vectors"].append(v)
        print(spec["id"], v["lnl"])
    (HERE / "oracle_vectors.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
