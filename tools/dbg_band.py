"""GPU debug helper: the cases of tests/test_gpu_s16.py::test_pairs_all_stripe_widths[band], every mismatch listed.
   python tools/dbg_band.py <band> [single]   (single: also every case as a batch of its own)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen
from gam_ngs_b200 import capi
import gam_ngs_b200 as g
from test_gpu_parity import project, run_batch
from test_gpu_s16 import _shape
from util import oracle_expect

band = int(sys.argv[1])
rng = np.random.default_rng(8000 + band)
cases = []
for length in (70, 333, 700, 1500):
    for _ in range(7):
        a, b = gen.make_pair(rng, length + int(rng.integers(0, 60)), div=float(rng.choice([0.0, 0.02, 0.1])), p_n=0.0)
        b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
        cases.append(dict(a=a, b=b, band=band, gap=-8, **_shape(rng, a, b, int(rng.integers(0, 4)))))
cases = cases[:-1]
ctx = g.Context(devices=[0])
exps = [oracle_expect(c) for c in cases]
def diff(got, exp):
    return {k: (got.get(k), exp.get(k)) for k in set(exp) | set(got) if k != "ops" and got.get(k) != exp.get(k)} or ("ops" if got != exp else {})
for mode in (capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE):
    got = run_batch(ctx, cases, mode)
    for k in range(len(cases)):
        e = project(exps[k], mode)
        if got[k] != e:
            print("batch mode", mode, "case", k, {a: b for a, b in cases[k].items() if a not in "ab"}, "la", len(cases[k]["a"]), "lb", len(cases[k]["b"]), diff(got[k], e))
    if len(sys.argv) > 2:
        for k in range(len(cases)):
            got1 = run_batch(ctx, [cases[k]], mode)
            e = project(exps[k], mode)
            if got1[0] != e:
                print("single mode", mode, "case", k, diff(got1[0], e))
print("done")
