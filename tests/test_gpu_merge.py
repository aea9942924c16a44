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


def test_merge_stage_matches_golden_from_reference_caller(ctx):
    """gamx_merge_align against tests/golden/merge_golden.json: vectors generated from the UNMODIFIED bodies of
    PctgBuilder::alignMergeBlock & co. (tests/golden/make_merge_golden.py), all four outcome classes."""
    from util import load_merge_golden
    n = 0
    for M, S, mbs in load_merge_golden():
        arr, blk = to_arrays(g, M, S, mbs, ctx)
        res, _ = ctx.merge_align(arr, blk)
        for k, mb in enumerate(mbs):
            assert result_dict(res[k]) == mb["expect"], (k, mb["m"], mb["s"], len(mb["blocks"]), mb["tails"])
            n += 1
    assert n >= 50


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_merge_stage_matches_compiled_reference_caller(ctx, seed):
    """The same without any restatement in between: gamx_merge_align against oracle/_ref/libgamref.so
    (the reference's caller code compiled by oracle/pctg_shim.cc; the prebuilt library travels to the GPU box)."""
    import oracle
    if not oracle.reference_available():
        pytest.skip("reference build not present")
    ref = oracle.reference()
    rng = np.random.default_rng(seed)
    M, S, MB = gen.make_assembly(rng, genome_len=60_000, master_mean=12_000, slave_mean=9_000, trim_prob=0.5,
                                 wrong_strand_prob=0.3, p_n=0.002)
    S, MB = gen.perturb_merge_blocks(rng, M, S, MB)
    arr, blk = to_arrays(g, M, S, MB, ctx)
    res, _ = ctx.merge_align(arr, blk)
    for k, mb in enumerate(MB):
        want = ref.align_merge_block(M[mb["m"]], S[mb["s"]], mb["blocks"], mb["tails"])
        assert result_dict(res[k]) == want, (k, mb["m"], mb["s"], len(mb["blocks"]), mb["tails"])
