"""Shared helpers for the merge-stage (batch collector) tests."""
import numpy as np

import oracle
from oracle import merge_oracle as mo


def oracle_merge(masters, slaves, merge_blocks, checker=None):
    """Runs the sequential restatement of alignMergeBlock on every merge block.
    Returns ([result dicts like gamx_merge_result], stats)."""
    orc = mo.MergeOracle(checker or oracle.restatement())
    out = []
    for mbd in merge_blocks:
        blocks = [mo.Block(b["num_reads"], mo.Frame("+-"[b["m_strand"]], b["m_begin"], b["m_end"]),
                           mo.Frame("+-"[b["s_strand"]], b["s_begin"], b["s_end"])) for b in mbd["blocks"]]
        mb = mo.MergeBlock(mbd["m"], mbd["s"], blocks, *[bool(x) for x in mbd.get("tails", (1, 1, 1, 1))])
        try:
            orc.align_merge_block(mb, masters[mbd["m"]], slaves[mbd["s"]])
            d = dict(status=0, align_ok=int(mb.align_ok), coords_set=int(mb.coords_set))
            if mb.coords_set:
                d.update(align_rev=int(mb.align_rev), m_start=mb.m_start, m_end=mb.m_end, s_start=mb.s_start, s_end=mb.s_end)
        except (IndexError, ValueError):
            d = dict(status=2)
        out.append(d)
    return out, orc.stats


def to_arrays(g, masters, slaves, merge_blocks, ctx):
    """Uploads the contigs and builds the gamx_merge_block / gamx_block arrays."""
    from gam_ngs_b200 import capi
    ctx.clear_contigs()
    m_ids = [ctx.add_contig(c) for c in masters]
    s_ids = [ctx.add_contig(c) for c in slaves]
    nb = sum(len(m["blocks"]) for m in merge_blocks)
    blk = np.zeros(nb, dtype=capi.BLOCK_DTYPE)
    mbs = np.zeros(len(merge_blocks), dtype=capi.MERGE_BLOCK_DTYPE)
    k = 0
    for i, m in enumerate(merge_blocks):
        mbs[i]["m_id"], mbs[i]["s_id"] = m_ids[m["m"]], s_ids[m["s"]]
        mbs[i]["first_block"], mbs[i]["n_blocks"] = k, len(m["blocks"])
        t = m.get("tails", (1, 1, 1, 1))
        mbs[i]["m_ltail"], mbs[i]["m_rtail"], mbs[i]["s_ltail"], mbs[i]["s_rtail"] = t
        for b in m["blocks"]:
            for f in ("num_reads", "m_strand", "s_strand", "m_begin", "m_end", "s_begin", "s_end"):
                blk[k][f] = b[f]
            k += 1
    return mbs, blk


def result_dict(r):
    d = dict(status=int(r["status"]))
    if d["status"]:
        return d
    d.update(align_ok=int(r["align_ok"]), coords_set=int(r["coords_set"]))
    if d["coords_set"]:
        d.update(align_rev=int(r["align_rev"]), m_start=int(r["m_start"]), m_end=int(r["m_end"]),
                 s_start=int(r["s_start"]), s_end=int(r["s_end"]))
    return d
