"""BASELINE config 5 (scaled down): divergence/indel sweep and band-width sweep on mixed-length pairs
(lengths log-uniform in [256, 16384]); kernel-only GCUPS per point, score+endpoints and full-ops.
Also a config-3 point (10-50 kb overlaps, band 256, full traceback + edit strings).
Writes JSON lines.  python tools/sweep.py [pairs] [all|bands]   (bands: only the band-width sweep, e.g. at a
larger pair count, where the last wave's traceback and the launch tails weigh less)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import gam_ngs_b200 as g  # noqa: E402


def run_point(ctx, a, al, b, bl, band, modes, label, int_peak):
    n = len(al)
    ctx.clear_contigs()
    ctx.add_contigs(np.concatenate([a, b]), np.concatenate([al, bl]))
    jobs = g.make_jobs(n)
    jobs["a_id"] = np.arange(n); jobs["b_id"] = np.arange(n, 2 * n)
    jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = band
    out = dict(label)
    for mode, name in modes:
        jobs["mode"] = mode
        plan = ctx.plan(jobs)
        for _ in range(2):
            plan.run(); plan.sync()
        ms = []
        for _ in range(3):
            plan.run(); plan.sync(); ms.append(plan.last_ms)
        res, _ = plan.fetch()
        gc = plan.cells / (min(ms) * 1e-3) / 1e9
        out[name] = {"gcups": round(gc, 1), "ms": round(min(ms), 3), "frac_int_peak_4ops": round(gc * 4e9 / int_peak, 3),
                     "ok": int((res["status"] == 0).sum()), "mean_homology": float(res["homology"].mean()) if mode else None}
        plan.close()
    out["pairs"] = n
    print(json.dumps(out), flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    ctx = g.Context(devices=[0])
    int_peak = ctx.measure_int_peak(0)
    rng = np.random.default_rng(5)
    lengths = np.exp(rng.uniform(np.log(256), np.log(16384), size=n)).astype(np.int64)
    modes = [(1, "endpoints"), (2, "full")]
    only_bands = len(sys.argv) > 2 and sys.argv[2] == "bands"
    for d in (() if only_bands else (0.0, 0.01, 0.02, 0.05, 0.10)):
        for indel in (0.0, 0.5):
            a, al, b, bl = gen.bulk_pairs(rng, n, 0, div=d, indel_share=indel, lengths=lengths)
            run_point(ctx, a, al, b, bl, 64, modes, {"sweep": "divergence", "div": d, "indel_share": indel, "band": 64}, int_peak)
    a, al, b, bl = gen.bulk_pairs(rng, n, 0, div=0.02, lengths=lengths)
    for band in (16, 32, 64, 128, 256, 512, 1024):
        run_point(ctx, a, al, b, bl, band, [(0, "score")] + modes, {"sweep": "band", "div": 0.02, "band": band}, int_peak)
    if only_bands:
        return
    # config 3 shape: long overlaps, band 256, full traceback + edit strings
    a, al, b, bl = gen.bulk_pairs(rng, 4000, 0, div=0.02, len_lo=10000, len_hi=50000)
    run_point(ctx, a, al, b, bl, 256, [(0, "score")] + modes, {"sweep": "config3", "div": 0.02, "band": 256, "len": "10-50kb"}, int_peak)


if __name__ == "__main__":
    main()
