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
