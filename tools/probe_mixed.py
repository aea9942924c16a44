"""Kernel-only run of a config-5 shaped batch (pairs of mixed length, log-uniform lo..hi) for launch lists:
probe_mixed.py pairs band mode [reps] [lo] [hi]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen
import gam_ngs_b200 as g


def main():
    npairs, band, mode = (int(x) for x in sys.argv[1:4])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    lo = int(sys.argv[5]) if len(sys.argv) > 5 else 256
    hi = int(sys.argv[6]) if len(sys.argv) > 6 else 16384
    ctx = g.Context(devices=[0])
    rng = np.random.default_rng(5)
    lengths = np.exp(rng.uniform(np.log(lo), np.log(hi), npairs)).astype(np.int64)
    a, al, b, bl = gen.bulk_pairs(rng, npairs, 0, div=0.02, lengths=lengths)
    ctx.add_contigs(np.concatenate([a, b]), np.concatenate([al, bl]))
    jobs = g.make_jobs(npairs)
    jobs["a_id"] = np.arange(npairs); jobs["b_id"] = np.arange(npairs, 2 * npairs)
    jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = band; jobs["mode"] = mode
    plan = ctx.plan(jobs)
    ms = []
    for _ in range(reps):
        plan.run(); plan.sync(); ms.append(plan.last_ms)
    res, _ = plan.fetch()
    print(json.dumps({"pairs": npairs, "lengths": [lo, hi], "band": band, "mode": mode, "ms": min(ms),
                      "gcups": plan.cells / (min(ms) * 1e-3) / 1e9, "launches": int(plan.kernel_launches),
                      "score_sum": int(res["score"].sum())}), flush=True)


if __name__ == "__main__":
    main()
