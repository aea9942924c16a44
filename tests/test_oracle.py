"""CPU tests: the C restatement (oracle/bsw_oracle.c) against (1) the committed golden
vectors generated from the unmodified reference and (2) the reference itself when its
build (oracle/_ref/libgamref.so) is present."""
import numpy as np
import pytest

import gen
import oracle
from util import load_golden, oracle_expect, x_size_of


def test_restatement_matches_golden():
    cases = load_golden()
    assert len(cases) >= 200
    seen = set()
    for job, exp in cases:
        got = oracle_expect(job)
        assert got == exp, (job, got, exp)
        seen.add(exp["status"])
    assert seen == {0, 1, 2}  # alignments, empty results and out_of_range all covered


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
def test_reference_matches_golden():
    ref = oracle.reference()
    for job, exp in load_golden():
        assert oracle_expect(job, ref) == exp


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_matches_reference_fuzz(seed):
    ref, rst = oracle.reference(), oracle.restatement()
    rng = np.random.default_rng(seed)
    n = 0
    while n < 3000:
        job = gen.fuzz_case(rng)
        x = x_size_of(job)
        if x is not None and (x == 0 or x > 5000):
            continue  # x_size == 0 is undefined behaviour in the reference (.cc:102-122)
        n += 1
        assert oracle_expect(job, rst) == oracle_expect(job, ref), job


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
def test_restatement_matches_reference_config_shapes():
    """BASELINE.json config 2 / config 3 shapes at sizes the CPU finishes in seconds."""
    ref, rst = oracle.reference(), oracle.restatement()
    rng = np.random.default_rng(7)
    for length, band, p_n in [(1000, 64, 0.0), (1000, 64, 0.001), (6000, 256, 0.0), (3000, 150, 0.001)]:
        for _ in range(3):
            a, b = gen.make_pair(rng, length, div=0.02, p_n=p_n, offset=int(rng.integers(0, band // 2)))
            job = dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1,
                       band=band, gap=-8, force_start=False, force_end=False)
            e1, e2 = oracle_expect(job, rst), oracle_expect(job, ref)
            assert e1 == e2
            assert e1["status"] == 0 and e1["homology"] > 90


def test_revcomp_matches_reference_semantics():
    rng = np.random.default_rng(3)
    s = gen.random_seq(rng, 101, p_n=0.1)
    rc = gen.revcomp(s)
    assert (gen.revcomp(rc) == s).all()
    if oracle.reference_available():
        assert (oracle.reference().revcomp(s) == rc).all()


def _hits_case(rng):
    la = int(rng.integers(0, 400))
    p_n = float(rng.choice([0.0, 0.0, 0.02, 0.2]))
    a = gen.random_seq(rng, la, p_n)
    kind = rng.integers(0, 4)
    if kind == 0 or la < 30:
        b = gen.random_seq(rng, int(rng.integers(0, 400)), p_n)
    elif kind == 1:  # b = mutated window of a: one dominant diagonal
        s = int(rng.integers(0, la // 2)); e = int(rng.integers(s + 1, la + 1))
        b = gen.mutate(rng, a[s:e], div=float(rng.choice([0.0, 0.02, 0.1])))
    elif kind == 2:  # low complexity: many equal k-mers, ties between diagonals
        unit = gen.random_seq(rng, int(rng.integers(1, 6)))
        a = np.resize(unit, la).astype(np.uint8)
        b = np.resize(unit, int(rng.integers(20, 200))).astype(np.uint8)
    else:            # b longer than a / prefix
        b = np.concatenate([gen.random_seq(rng, int(rng.integers(0, 50))), a, gen.random_seq(rng, int(rng.integers(0, 50)))])
    lb = len(b)
    if rng.random() < 0.5:
        w = (0, max(la - 1, 0), 0, max(lb - 1, 0))
    else:
        w = (int(rng.integers(0, la + 3)), int(rng.integers(0, la + 30)), int(rng.integers(0, lb + 3)), int(rng.integers(0, lb + 30)))
    return a, b, w


@pytest.mark.skipif(not oracle.reference_available(), reason="reference build not present")
def test_find_hits_restatement_matches_reference():
    """ABlast::findHits (ablast.cc:41-76) restated in oracle/bsw_oracle.c vs the compiled reference,
    including the radix-4 aliasing of N and ties between diagonals."""
    ref, rst = oracle.reference(), oracle.restatement()
    rng = np.random.default_rng(42)
    nonempty = ties = 0
    for _ in range(1500):
        a, b, w = _hits_case(rng)
        want = ref.find_hits(a, w[0], w[1], b, w[2], w[3])
        got, mc = rst.find_hits(a, w[0], w[1], b, w[2], w[3])
        assert list(got) == list(want), (w, len(a), len(b))
        nonempty += len(want) > 0
        ties += len(want) > 1
    assert nonempty > 300 and ties > 50
