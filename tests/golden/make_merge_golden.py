#!/usr/bin/env python3
"""Generates tests/golden/merge_golden.json from the UNMODIFIED reference caller code: the bodies of
PctgBuilder::alignMergeBlock / findBestAlignment / alignBlocks / is_good (PctgBuilder.cc:726-844,1361-1730) as
compiled by oracle/pctg_shim.cc (`make -C oracle ref`).  Run in the build container (needs /root/reference):

    python tests/golden/make_merge_golden.py

Assemblies: master and slave contigs (strings over ATCGN).  Each case: one merge block of an assembly (contig
indices, blocks, tail flags) and what the reference wrote into the MergeBlock.  The -m "not gpu" tests pin oracle/merge_oracle.py to these, the
-m gpu tests pin gamx_merge_align (include/gamx.h)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import gen  # noqa: E402
import oracle  # noqa: E402

ALPHA = "ATCGN"


def main():
    ref = oracle.reference()
    assemblies, cases, classes = [], [], {}
    for seed in (101, 102, 103, 104):
        rng = np.random.default_rng(seed)
        M, S, MB = gen.make_assembly(rng, genome_len=24_000, master_mean=6_000, slave_mean=4_500, frame_lo=300, frame_hi=1500,
                                     trim_prob=0.5, wrong_strand_prob=0.3, p_n=0.002)
        S, MB = gen.perturb_merge_blocks(rng, M, S, MB, n_extra=4)
        assemblies.append(dict(masters=["".join(ALPHA[c] for c in m) for m in M], slaves=["".join(ALPHA[c] for c in s_) for s_ in S]))
        for mb in MB:
            want = ref.align_merge_block(M[mb["m"]], S[mb["s"]], mb["blocks"], mb["tails"])
            key = (want.get("status"), want.get("align_ok"), want.get("coords_set"))
            classes[key] = classes.get(key, 0) + 1
            cases.append(dict(assembly=len(assemblies) - 1, m=mb["m"], s=mb["s"], blocks=mb["blocks"], tails=list(mb["tails"]), expect=want))
    with open(os.path.join(HERE, "merge_golden.json"), "w") as f:
        json.dump(dict(source="PctgBuilder.cc:726-844,1361-1730 via oracle/pctg_shim.cc", assemblies=assemblies, cases=cases), f)
    print(len(cases), "cases;", {str(k): v for k, v in classes.items()})


if __name__ == "__main__":
    main()
