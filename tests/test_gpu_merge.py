"""GPU parity of the batch collector (gamx_merge_align: alignMergeBlock / findBestAlignment /
alignBlocks as rounds of GPU batches, include/gamx.h) against the sequential oracle
(oracle/merge_oracle.py) on synthetic master/slave assemblies (BASELINE config 1 shape, scaled down)."""
import numpy as np
import pytest

import gen
import gam_ngs_b200 as g
from merge_util import oracle_merge, result_dict, to_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(devices=[0])
    yield c
    c.close()


@pytest.mark.parametrize("seed,kw", [
    (11, dict()),
    (12, dict(trim_prob=0.8, wrong_strand_prob=0.4, p_n=0.001)),
    (13, dict(trim_prob=0.5, wrong_strand_prob=0.5, div=0.03)),
    (14, dict(trim_prob=0.9, rc_frac=1.0, div=0.06)),      # 6 % divergence: homology below 95 -> merges refused
])
def test_merge_stage_matches_oracle(ctx, seed, kw):
    rng = np.random.default_rng(seed)
    M, S, MB = gen.make_assembly(rng, genome_len=150_000, master_mean=30_000, slave_mean=20_000, **kw)
    # unrelated pair, an empty merge block and restricted tail flags
    MB.append(dict(m=0, s=len(S) - 1, blocks=[dict(num_reads=10, m_strand=0, s_strand=0, m_begin=100, m_end=1500,
                                                   s_begin=50, s_end=1400)]))
    for k, mb in enumerate(MB):
        if k % 3 == 1:
            mb["tails"] = (int(rng.integers(0, 2)), 1, 1, int(rng.integers(0, 2)))
    want, ostats = oracle_merge(M, S, MB)
    mbs, blk = to_arrays(g, M, S, MB, ctx)
    res, stats = ctx.merge_align(mbs, blk)
    for k in range(len(MB)):
        assert result_dict(res[k]) == want[k], (k, MB[k]["m"], MB[k]["s"], len(MB[k]["blocks"]))
    assert stats["alignments"] == ostats.alignments and stats["hits_calls"] == ostats.hits_calls
    assert stats["cells"] == ostats.cells
    # rounds are bounded by the longest chain (two orientations) plus the two tail rounds
    assert stats["rounds"] <= 2 * max(len(m["blocks"]) for m in MB) + 2
