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
    # one process per GPU (torchrun): the merge blocks shard independently - rank r takes the blocks the
    # cost-balanced split (gamx_shard_by_cost, cost = bases of the two contigs) assigns to shard r; no collective
    # on the data path, the timing is the slowest rank's (barrier + max over ranks)
    from gam_ngs_b200.dist import Ranks
    from gam_ngs_b200 import capi
    ranks = Ranks()
    all_mb = MB
    if ranks.world > 1:
        cost = np.array([sum(b["m_end"] - b["m_begin"] + 1 for b in m["blocks"]) for m in all_mb], dtype=np.uint64)
        shard = capi.shard_by_cost(cost, ranks.world)
        MB = [m for m, s_ in zip(all_mb, shard) if s_ == ranks.rank]
    ctx = g.Context(devices=[ranks.local_rank])
    mbs, blk = to_arrays(g, M, S, MB, ctx)
    ctx.merge_align(mbs[:4], blk)  # warm-up (contig upload, kernel load)
    times = []
    for _ in range(3):
        ranks.barrier()
        t0 = time.perf_counter()
        res, stats = ctx.merge_align(mbs, blk)
        times.append(ranks.max(time.perf_counter() - t0))
    gpu_s = min(times)
    if ranks.world > 1:
        tot = {k: int(ranks.sum(float(v))) for k, v in stats.items() if k != "rounds"}
        tot["rounds"] = int(ranks.max(float(stats["rounds"])))
        ok_all, exc_all = int(ranks.sum(float(res["align_ok"].sum()))), int(ranks.sum(float((res["status"] != 0).sum())))
        if ranks.rank == 0:
            print(json.dumps({"config": "cfg4_merge_alignment_stage_sharded", "genome_bp": genome, "n_gpus": ranks.world,
                              "merge_blocks": len(all_mb), "merge_blocks_rank0": len(MB), "scaling": "strong",
                              "gpu": {"seconds": gpu_s, "gcups": tot["cells"] / gpu_s / 1e9, **tot},
                              "align_ok": ok_all, "exceptions": exc_all, "gen_seconds": t_gen}), flush=True)
        ranks.close()
        return
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
