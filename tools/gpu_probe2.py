"""Compares lane-group geometries on config-2 shaped batches (run with GAMX_FORCE_LG set)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen
import gam_ngs_b200 as g

def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    length = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    band = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    ctx = g.Context(devices=[0])
    rng = np.random.default_rng(2)
    a, al, b, bl = gen.bulk_pairs(rng, npairs, length)
    ctx.add_contigs(np.concatenate([a, b]), np.concatenate([al, bl]))
    jobs = g.make_jobs(npairs)
    jobs["a_id"] = np.arange(npairs); jobs["b_id"] = np.arange(npairs, 2 * npairs)
    jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = band
    for mode, name in [(0, "score"), (1, "endpoints"), (2, "full")]:
        jobs["mode"] = mode
        plan = ctx.plan(jobs)
        for _ in range(2):
            plan.run(); plan.sync()
        ms = []
        for _ in range(3):
            plan.run(); plan.sync(); ms.append(plan.last_ms)
        res, _ = plan.fetch()
        print(json.dumps({"force_lg": os.environ.get("GAMX_FORCE_LG"), "pairs": npairs, "len": length, "band": band, "mode": name,
                          "ms": min(ms), "gcups": plan.cells / (min(ms) * 1e-3) / 1e9, "score_sum": int(res["score"].sum()),
                          "ok": int((res["status"] == 0).sum())}), flush=True)
        plan.close()

if __name__ == "__main__":
    main()
