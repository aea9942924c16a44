"""GPU parity tests of the 16x2 pair kernels (k1s_kernel, bsw_warp16.h) through the C ABI: N-free
inputs so that the half-word path runs (windows with an N fall back to the 32-bit body, which
test_gpu_parity.py covers), against the oracle, bit-exact in all three modes."""
import numpy as np
import pytest

import gen
from gam_ngs_b200 import capi
from test_gpu_parity import ctx, project, run_batch  # noqa: F401  (ctx is a fixture)
from util import oracle_expect

pytestmark = pytest.mark.gpu


def _shape(rng, a, b, shape):
    la, lb = len(a), len(b)
    if shape == 0:
        return dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
    if shape == 1:
        return dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                    force_start=False, force_end=True)
    if shape == 2:
        return dict(begin_a=3, end_a=la + 40, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 5,
                    force_start=True, force_end=False)
    return dict(begin_a=int(rng.integers(0, la)), end_a=int(rng.integers(0, la + 30)),
                begin_b=int(rng.integers(0, lb // 2)), end_b=int(rng.integers(lb // 2, lb + 9)),
                force_start=bool(rng.integers(0, 2)), force_end=bool(rng.integers(0, 2)))


def _compare(ctx, cases, modes=(capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE)):
    exps = [oracle_expect(c) for c in cases]
    for mode in modes:
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode, {a: b for a, b in cases[k].items() if a not in "ab"})


@pytest.mark.parametrize("band", [0, 1, 7, 16, 33, 47, 64, 100, 130, 150, 200, 256, 271, 287])
def test_pairs_all_stripe_widths(ctx, band):
    """Every geometry the host picks (C = 2..18, LG = 4..32), mixed lengths (so the halves of a pair and
    the pairs of a warp finish at different steps), every window shape, odd job counts."""
    rng = np.random.default_rng(8000 + band)
    cases = []
    for length in (70, 333, 700, 1500):
        for _ in range(7):
            a, b = gen.make_pair(rng, length + int(rng.integers(0, 60)), div=float(rng.choice([0.0, 0.02, 0.1])), p_n=0.0)
            b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
            cases.append(dict(a=a, b=b, band=band, gap=-8, **_shape(rng, a, b, int(rng.integers(0, 4)))))
    cases = cases[:-1]  # odd count: the last pair has an idle half
    _compare(ctx, cases)


@pytest.mark.parametrize("gap", [-5, -13, -29])
def test_pairs_gap_values_and_extreme_inputs(ctx, gap):
    """The 16-bit range argument must hold for any input: homopolymers, unrelated sequences, long
    insertions, period-2 repeats; mildest and harshest gap of the fast kernels; many rebases."""
    rng = np.random.default_rng(8100 - gap)
    n = 3000
    homo = np.zeros(n, dtype=np.uint8)
    r1 = rng.integers(0, 4, n).astype(np.uint8)
    r2 = rng.integers(0, 4, n).astype(np.uint8)
    ins = np.concatenate([r1[:1000], rng.integers(0, 4, 100).astype(np.uint8), r1[1000:]])
    alt = np.tile(np.array([0, 1], dtype=np.uint8), n // 2)
    seqs = [(homo, homo.copy()), (r1, r2), (r1, ins), (ins, r1), (alt, np.roll(alt, 1)), (homo, alt), (r1, r1.copy()), (alt, alt.copy())]
    for band in (20, 64, 150, 280):
        cases = [dict(a=a, b=b, band=band, gap=gap, **_shape(rng, a, b, 0)) for a, b in seqs]
        _compare(ctx, cases, modes=(capi.MODE_FULL, capi.MODE_SCORE))


def test_pairs_and_fallback_in_one_batch(ctx):
    """Jobs with and without N, two bands and two gaps in one batch: pairs form inside a (band, gap) group,
    a warp whose jobs hold an N runs the 32-bit body - results do not depend on the route."""
    rng = np.random.default_rng(8200)
    cases = []
    for n in range(300):
        a, b = gen.make_pair(rng, int(rng.integers(100, 900)), div=0.03, p_n=0.002 if n % 5 == 0 else 0.0)
        cases.append(dict(a=a, b=b, band=int(rng.choice([64, 150])), gap=int(rng.choice([-8, -6])),
                          **_shape(rng, a, b, int(rng.integers(0, 4)))))
    _compare(ctx, cases)


def test_pairs_long_jobs_many_rebases(ctx):
    """Rows in the thousands (the 32-bit score range of row 6553 onwards does not fit a half-word; the
    per-lane bases carry it), the warp-per-job traceback of long jobs on pair regions."""
    rng = np.random.default_rng(8300)
    cases = []
    for length, band in [(5000, 64), (9000, 64), (7000, 150), (6000, 256), (4200, 16)]:
        for _ in range(3):
            a, b = gen.make_pair(rng, length + int(rng.integers(0, 500)), div=0.03, p_n=0.0)
            cases.append(dict(a=a, b=b, band=band, gap=-8, **_shape(rng, a, b, 0)))
    _compare(ctx, cases, modes=(capi.MODE_FULL, capi.MODE_SCORE))


def test_pairs_views(ctx):
    """Reverse-complement views and offsets into stored contigs, as the merge stage uses them."""
    rng = np.random.default_rng(8400)
    cases, views, exps = [], [], []
    for _ in range(40):
        a, b = gen.make_pair(rng, int(rng.integers(150, 900)), div=0.03, p_n=0.0)
        a_off = int(rng.integers(0, 40)); b_off = int(rng.integers(0, 40))
        a_len = len(a) - a_off - int(rng.integers(0, 20)); b_len = len(b) - b_off - int(rng.integers(0, 20))
        va, vb = a[a_off:a_off + a_len], b[b_off:b_off + b_len]
        mat = dict(a=va, b=vb, begin_a=0, end_a=len(va) - 1, begin_b=0, end_b=len(vb) - 1, band=64, gap=-8,
                   force_start=False, force_end=False)
        exps.append(oracle_expect(mat))
        cases.append(dict(mat, a=gen.revcomp(a), b=gen.revcomp(b)))
        views.append(dict(a_rc=1, a_off=a_off, a_len=a_len, b_rc=1, b_off=b_off, b_len=b_len))
    for mode in (capi.MODE_FULL, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode, views=views)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode)
