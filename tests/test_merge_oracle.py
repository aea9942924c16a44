"""CPU tests of the merge-stage oracle (oracle/merge_oracle.py): the control flow restated from
PctgBuilder.cc:726-844,1361-1730 is exercised with both pinned alignment checkers (the C restatement
and the compiled reference) and must not depend on which one is used; basic expectations on synthetic
assemblies guard the generator and the state machine."""
import numpy as np
import pytest

import gen
import oracle
from merge_util import oracle_merge


def _assembly(seed, **kw):
    rng = np.random.default_rng(seed)
    return gen.make_assembly(rng, genome_len=90_000, master_mean=25_000, slave_mean=18_000, **kw)


def test_merge_oracle_clean_assembly_merges_everything():
    M, S, MB = _assembly(1)
    res, stats = oracle_merge(M, S, MB)
    assert len(MB) >= 6 and stats.alignments >= len(MB)
    for mb, r in zip(MB, res):
        assert r["status"] == 0 and r["align_ok"] == 1 and r["coords_set"] == 1, (mb["m"], mb["s"], r)
        assert r["m_start"] <= r["m_end"] < len(M[mb["m"]]) and r["s_start"] <= r["s_end"] < len(S[mb["s"]])


def test_merge_oracle_tails_retries_and_bad_pairs():
    M, S, MB = _assembly(2, trim_prob=0.8, wrong_strand_prob=0.4, p_n=0.001)
    # an unrelated pair: the chained alignments fail in both orientations -> align_ok = false
    MB.append(dict(m=0, s=len(S) - 1, blocks=[dict(num_reads=10, m_strand=0, s_strand=0, m_begin=100, m_end=1500,
                                                   s_begin=50, s_end=1400)]))
    res, stats = oracle_merge(M, S, MB)
    assert stats.hits_calls > 0            # tail alignments were seeded by findHits
    assert res[-1]["status"] == 0 and res[-1]["align_ok"] == 0 and res[-1]["coords_set"] == 0
    assert sum(r.get("align_ok", 0) for r in res) >= len(MB) // 2


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
def test_merge_oracle_same_with_reference_aligner():
    M, S, MB = _assembly(3, trim_prob=0.7, wrong_strand_prob=0.3, p_n=0.002)
    a, _ = oracle_merge(M, S, MB, oracle.restatement())

    class RefChecker:  # the compiled reference for alignments and findHits
        def __init__(self):
            self.r = oracle.reference()

        def align(self, *args, **kw):
            res, ops = self.r.align(*args, **kw)
            return res, ops

        def find_hits(self, *args):
            return self.r.find_hits(*args)

    b, _ = oracle_merge(M, S, MB, RefChecker())
    assert a == b


# ---- the restated control flow pinned to the reference's own caller code ------------------------------------
# (PctgBuilder::alignMergeBlock / findBestAlignment / alignBlocks / is_good, PctgBuilder.cc:726-844,1361-1730:
#  the unmodified bodies compiled by oracle/pctg_shim.cc, or the golden vectors generated from them)

def test_merge_oracle_matches_golden_from_reference_caller():
    from util import load_merge_golden
    n, classes = 0, set()
    for M, S, mbs in load_merge_golden():
        got, _ = oracle_merge(M, S, mbs)
        for mb, r in zip(mbs, got):
            assert r == mb["expect"], (mb["m"], mb["s"], len(mb["blocks"]), mb["tails"])
            classes.add((r["status"], r.get("align_ok"), r.get("coords_set")))
            n += 1
    assert n >= 50
    # merged, refused by the chained alignments, refused by a tail alignment, reference throws
    assert {(0, 1, 1), (0, 0, 0), (0, 0, 1), (2, None, None)} <= classes


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
@pytest.mark.parametrize("seed", [10, 11, 12, 13, 14, 15])
def test_merge_oracle_pinned_to_reference_caller(seed):
    """Adversarial merge blocks (reverse block order, shifted / empty / out-of-contig frames, mixed strand
    evidence, damaged slaves, unrelated pairs, random tail flags): restatement == compiled reference."""
    ref = oracle.reference()
    rng = np.random.default_rng(seed)
    M, S, MB = gen.make_assembly(rng, genome_len=60_000, master_mean=12_000, slave_mean=9_000, trim_prob=0.5,
                                 wrong_strand_prob=0.3, p_n=0.002)
    S, MB = gen.perturb_merge_blocks(rng, M, S, MB)
    got, _ = oracle_merge(M, S, MB)
    for mb, r in zip(MB, got):
        want = ref.align_merge_block(M[mb["m"]], S[mb["s"]], mb["blocks"], mb["tails"])
        assert r == want, (mb["m"], mb["s"], len(mb["blocks"]), mb["tails"])
