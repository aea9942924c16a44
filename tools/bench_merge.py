"""BASELINE config 1 / 4: gam-merge's alignment stage on a synthetic master/slave assembly pair.

GPU arm: gamx_merge_align (batch collector: rounds of GPU batches over all merge blocks).
CPU arm: the same call pattern driven sequentially per merge block (oracle/merge_oracle.py) with every
alignment executed by the compiled reference aligner (oracle/_ref), threads = 1 and = all host cores
(merge blocks are distributed over a thread pool, the parallel shape of ThreadedBuildPctg.cc:159-169).
Prints one JSON line.   python tools/bench_merge.py [genome_len] [cpu_merge_block_sample]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import oracle  # noqa: E402
import gam_ngs_b200 as g  # noqa: E402
from merge_util import oracle_merge, result_dict, to_arrays  # noqa: E402


def main():
    genome = int(sys.argv[1]) if len(sys.argv) > 1 else 2_900_000
    sample = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    rng = np.random.default_rng(1)
    t0 = time.perf_counter()
    M, S, MB = gen.make_assembly(rng, genome_len=genome, master_mean=60_000, slave_mean=40_000, div=0.01,
                                 trim_prob=0.5, wrong_strand_prob=0.1)
    t_gen = time.perf_counter() - t0
    ctx = g.Context(devices=[0])
    mbs, blk = to_arrays(g, M, S, MB, ctx)
    ctx.merge_align(mbs[:4], blk)  # warm-up (contig upload, kernel load)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        res, stats = ctx.merge_align(mbs, blk)
        times.append(time.perf_counter() - t0)
    gpu_s = min(times)
    # CPU arms on a bounded sample of merge blocks (same call pattern, reference aligner)
    class RefChecker:
        def __init__(self):
            self.r = oracle.reference()
        def align(self, *a, **k):
            return self.r.align(*a, **k)
        def find_hits(self, *a):
            return self.r.find_hits(*a)
    idx = list(range(0, len(MB), max(1, len(MB) // sample)))[:sample]
    sub = [MB[i] for i in idx]
    t0 = time.perf_counter()
    want, ostats = oracle_merge(M, S, sub, RefChecker())
    cpu1_s = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    def one(mb):
        return oracle_merge(M, S, [mb], RefChecker())[0][0]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        par = list(ex.map(one, sub))
    cpuN_s = time.perf_counter() - t0
    mismatches = sum(result_dict(res[i]) != w for i, w in zip(idx, want)) + sum(p != w for p, w in zip(par, want))
    line = {"config": "cfg1_merge_alignment_stage", "genome_bp": genome, "master_contigs": len(M), "slave_contigs": len(S),
            "merge_blocks": len(MB), "gpu": {"seconds": gpu_s, "gcups": stats["cells"] / gpu_s / 1e9, **stats},
            "cpu_reference": {"sample_merge_blocks": len(sub), "cells": ostats.cells,
                              "threads_1": {"seconds": cpu1_s, "gcups": ostats.cells / cpu1_s / 1e9},
                              f"threads_{cores}": {"seconds": cpuN_s, "gcups": ostats.cells / cpuN_s / 1e9}, "cores": cores},
            "parity_mismatches_on_sample": int(mismatches), "align_ok": int(res["align_ok"].sum()),
            "exceptions": int((res["status"] != 0).sum()), "gen_seconds": t_gen}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
