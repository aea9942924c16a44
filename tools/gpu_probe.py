"""Quick GPU probe: integer/DPX issue peaks and first throughput numbers (writes JSON lines)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import gam_ngs_b200 as g  # noqa: E402


def main():
    ctx = g.Context(devices=[0])
    names = ["viaddmax_s32", "vimax3_s32", "viaddmax_s16x2", "lop3", "prmt", "imad"]
    peaks = {n: ctx.measure_int_peak(i) for i, n in enumerate(names)}
    print(json.dumps({"int_peaks_lane_ops_per_s": peaks}), flush=True)
    rng = np.random.default_rng(2)
    for (npairs, length, band) in [(20000, 1000, 64), (2000, 1000, 150), (300, 30000, 256)]:
        ctx.clear_contigs()
        jobs = g.make_jobs(npairs)
        base_a, base_b = gen.make_pair(rng, length, div=0.02)
        t0 = time.time()
        uniq = min(npairs, 512)
        ids = []
        for k in range(uniq):
            a, b = gen.make_pair(rng, length, div=0.02)
            ids.append((ctx.add_contig(a), ctx.add_contig(b), len(a), len(b)))
        for k in range(npairs):
            ia, ib, la, lb = ids[k % uniq]
            jobs[k]["a_id"], jobs[k]["b_id"] = ia, ib
            jobs[k]["end_a"], jobs[k]["end_b"] = la - 1, lb - 1
            jobs[k]["band"] = band
        for mode, mname in [(0, "score"), (1, "endpoints"), (2, "full")]:
            jobs["mode"] = mode
            plan = ctx.plan(jobs)
            for _ in range(2):
                plan.run(); plan.sync()
            ms = []
            for _ in range(3):
                plan.run(); plan.sync(); ms.append(plan.last_ms)
            res, ops = plan.fetch()
            t1 = time.time()
            r2, o2 = ctx.align_batch(jobs)
            t2 = time.time()
            cells = plan.cells
            print(json.dumps({"pairs": npairs, "len": length, "band": band, "mode": mname,
                              "kernel_ms": min(ms), "gcups_kernel": cells / (min(ms) * 1e-3) / 1e9,
                              "e2e_s": t2 - t1, "gcups_e2e": cells / (t2 - t1) / 1e9,
                              "status_ok": int((res["status"] == 0).sum()),
                              "score_sum": int(res["score"].sum())}), flush=True)
            plan.close()


if __name__ == "__main__":
    main()
