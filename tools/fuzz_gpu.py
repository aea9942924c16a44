"""One-off wide fuzz of the CUDA path against the oracle restatement (beyond the seeded cases of tests/):
python tools/fuzz_gpu.py [cases] [seed].  Random lengths 30-2500, bands 0-700 (every stripe width, all
lane-group sizes, K1 and K2), gaps, windows hanging off either contig, force flags, N content, reverse
complement views, all three modes.  Prints one JSON line; exits non-zero on the first mismatch."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen
import gam_ngs_b200 as g
from gam_ngs_b200 import capi
from util import oracle_expect
from test_gpu_parity import result_to_expect, project


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    ctx = g.Context(devices=[0])
    t0 = time.time()
    done = 0
    geos = {}
    while done < n:
        m = min(1500, n - done)
        cases, views = [], []
        for _ in range(m):
            length = int(np.exp(rng.uniform(np.log(30), np.log(2500))))
            band = int(rng.choice([0, 1, 5, 16, 33, 47, 64, 80, 100, 128, 150, 192, 200, 256, 271, 287, 300, 400, 700]))
            if band > 300 and length > 1200:
                length = 1200
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.01, 0.03, 0.1, 0.3])), p_n=float(rng.choice([0, 0, 0.003, 0.05])))
            if rng.random() < 0.5:
                b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
            la, lb = len(a), len(b)
            shape = int(rng.integers(0, 5))
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, max(1, la // 2))), end_a=la - 1, begin_b=0, end_b=lb - 1)
            elif shape == 2:
                w = dict(begin_a=int(rng.integers(0, 5)), end_a=la + int(rng.integers(0, 60)), begin_b=int(rng.integers(0, max(1, lb // 2))), end_b=lb + 5)
            elif shape == 3:
                ba = int(rng.integers(0, la)); bb = int(rng.integers(0, lb))
                w = dict(begin_a=ba, end_a=min(la - 1, ba + int(rng.integers(0, 400))), begin_b=bb, end_b=min(lb - 1, bb + int(rng.integers(0, 400))))
            else:
                w = dict(begin_a=int(rng.integers(0, la + band + 3)), end_a=int(rng.integers(0, la + 50)), begin_b=int(rng.integers(0, lb)), end_b=int(rng.integers(0, lb + 20)))
            job = dict(a=a, b=b, band=band, gap=int(rng.choice([-8, -8, -8, -5, -12, -29])),
                       force_start=bool(rng.random() < 0.2), force_end=bool(rng.random() < 0.2), **w)
            # reverse-complement view of b: the job refers to revcomp(revcomp(b)) = b through b_rc
            rc = bool(rng.random() < 0.3)
            cases.append(job); views.append(rc)
            geo = capi.band_geometry(band)
            geos[str(geo)] = geos.get(str(geo), 0) + 1
        exps = [oracle_expect(c) for c in cases]
        for mode in (capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE):
            ctx.clear_contigs()
            jobs = g.make_jobs(m)
            for k, c in enumerate(cases):
                jobs[k]["a_id"] = ctx.add_contig(c["a"])
                jobs[k]["b_id"] = ctx.add_contig(gen.revcomp(c["b"]) if views[k] else c["b"])
                jobs[k]["b_rc"] = int(views[k])
                for f in ("begin_a", "end_a", "begin_b", "end_b", "band", "gap"):
                    jobs[k][f] = c[f]
                jobs[k]["force_start"], jobs[k]["force_end"] = int(c["force_start"]), int(c["force_end"])
                jobs[k]["mode"] = mode
            res, ops = ctx.align_batch(jobs)
            for k in range(m):
                exp = exps[k]
                if exp.get("status") == 0 and exp.get("n_ops") == 0:
                    exp = {"status": 1}
                got = result_to_expect(ctx, res[k], ops, mode)
                if got != project(exp, mode):
                    print("MISMATCH", done + k, mode, {a: b for a, b in cases[k].items() if a not in "ab"}, views[k], got, project(exp, mode))
                    sys.exit(1)
        done += m
    print(json.dumps({"fuzz_cases": n, "seed": seed, "modes": 3, "mismatches": 0, "seconds": round(time.time() - t0, 1),
                      "geometries_hit": geos}))


if __name__ == "__main__":
    main()
