"""CPU tests of the kernel bodies through the lane simulator (tests/sim/warp_sim.cc): the same
bsw_warp.h / bsw_generic.h / bsw_host.h code the CUDA library compiles, executed by 32
cooperative fibers, compared with the oracle.  The -m gpu tests repeat this on the device."""
import numpy as np
import pytest

import gen
import simlib
from util import load_golden, oracle_expect, x_size_of


def _check(job, exp, modes=(2, 1, 0), lane_order=0, force_class=0, **view):
    for mode in modes:
        cls, r, ops = simlib.sim_align(job, mode=mode, lane_order=lane_order, force_class=force_class, **view)
        got = simlib.result_to_expect(r, ops if mode == 2 else None, mode)
        assert got == simlib.project(exp, mode), (mode, cls, {k: v for k, v in job.items() if k not in "ab"})
    return cls


def test_sim_matches_golden():
    classes = set()
    for n, (job, exp) in enumerate(load_golden()):
        if x_size_of(job) == 0:
            exp = {"status": 3}
        classes.add(_check(job, exp, lane_order=n & 1))
    assert classes == {0, 1, 2}  # early-out, warp kernel and generic kernel all exercised


def test_generic_body_matches_golden():
    for job, exp in load_golden():
        if x_size_of(job) == 0:
            continue
        _check(job, exp, modes=(2,), force_class=2)


@pytest.mark.parametrize("seed", [11, 12])
def test_sim_fuzz_small(seed):
    rng = np.random.default_rng(seed)
    n = 0
    while n < 400:
        job = gen.fuzz_case(rng)
        x = x_size_of(job)
        if x is not None and x > 3000:
            continue
        n += 1
        exp = oracle_expect(job) if x != 0 else {"status": 3}
        _check(job, exp, modes=(2, 0), lane_order=n & 1)


@pytest.mark.parametrize("band", [0, 1, 16, 47, 64, 100, 150, 256, 271])
def test_sim_warp_kernel_bands(band):
    """Every lane-stripe width C = 2..17, fast path, tile refills (rows > 256), force flags."""
    rng = np.random.default_rng(1000 + band)
    for length in (70, 333, 700):
        a, b = gen.make_pair(rng, length, div=0.05, p_n=0.004)
        b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
        la, lb = len(a), len(b)
        for shape in range(3):
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                         force_start=False, force_end=True)
            else:
                w = dict(begin_a=3, end_a=la + 40, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 5,
                         force_start=True, force_end=False)
            job = dict(a=a, b=b, band=band, gap=-8, **w)
            cls = _check(job, oracle_expect(job), modes=(2, 0), lane_order=shape & 1)
            assert cls == 1


def test_sim_views_reverse_complement_and_offsets():
    """Jobs address views (rc, offset, length) of stored contigs: must equal aligning the
    materialised reverse-complement / chopped contig (PctgBuilder.cc:1443, :1577)."""
    rng = np.random.default_rng(5)
    for _ in range(12):
        a, b = gen.make_pair(rng, int(rng.integers(150, 500)), div=0.03, p_n=0.01)
        band = int(rng.choice([20, 64, 150]))
        a_off = int(rng.integers(0, 40)); b_off = int(rng.integers(0, 40))
        a_len = len(a) - a_off - int(rng.integers(0, 20)); b_len = len(b) - b_off - int(rng.integers(0, 20))
        # the stored contigs hold the reverse complement of what the job must see
        a_store, b_store = gen.revcomp(a), gen.revcomp(b)
        va, vb = a[a_off:a_off + a_len], b[b_off:b_off + b_len]
        job_mat = dict(a=va, b=vb, begin_a=0, end_a=len(va) - 1, begin_b=0, end_b=len(vb) - 1, band=band,
                       gap=-8, force_start=False, force_end=False)
        exp = oracle_expect(job_mat)
        job_view = dict(job_mat, a=a_store, b=b_store)
        _check(job_view, exp, modes=(2,), a_rc=1, a_off=a_off, a_len=a_len, b_rc=1, b_off=b_off, b_len=b_len)
        _check(dict(job_mat, a=a, b=b_store), exp, modes=(2,), a_rc=0, a_off=a_off, a_len=a_len,
               b_rc=1, b_off=b_off, b_len=b_len, force_class=2)


@pytest.mark.parametrize("band", [20, 40, 64, 100, 130])
def test_sim_lane_groups_share_a_step_loop(band):
    """Bands that run 2 or 4 pairs per warp (LG = 16 / 8): pairs of different length and shape in the
    same warp must not disturb each other (masked drain, per-group capture windows)."""
    rng = np.random.default_rng(3000 + band)
    for rep in range(4):
        jobs = []
        for g in range(4):
            length = int(rng.integers(40, 420))
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.03, 0.2])), p_n=0.01)
            la, lb = len(a), len(b)
            shape = int(rng.integers(0, 3))
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                         force_start=False, force_end=True)
            else:
                w = dict(begin_a=2, end_a=la + 9, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 3,
                         force_start=True, force_end=False)
            jobs.append(dict(a=a, b=b, band=band, gap=-8, **w))
        for mode in (2, 0):
            out = simlib.sim_align_multi(jobs, mode=mode, lane_order=rep)
            ran = 0
            for job, o in zip(jobs, out):
                if o is None:
                    continue
                ran += 1
                r, ops = o
                got = simlib.result_to_expect(r, ops if mode == 2 else None, mode)
                assert got == simlib.project(oracle_expect(job), mode), (band, rep, mode)
            assert ran >= 1


@pytest.mark.parametrize("band,force", [(300, 0), (512, 0), (1024, 0), (2303, 0), (64, 3), (150, 3), (256, 3)])
def test_sim_cta_per_pair_kernel(band, force):
    """K2: one pair per CTA (64/128/256 lanes, neighbour exchange through shared memory).  Bands
    wider than a warp can hold, and warp-sized bands in latency mode (force=3)."""
    rng = np.random.default_rng(4000 + band)
    for length in (90, 400):
        a, b = gen.make_pair(rng, length, div=0.05, p_n=0.004)
        b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
        la, lb = len(a), len(b)
        for shape in range(3):
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                         force_start=False, force_end=True)
            else:
                w = dict(begin_a=3, end_a=la + 40, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 5,
                         force_start=True, force_end=False)
            job = dict(a=a, b=b, band=band, gap=-8, **w)
            cls = _check(job, oracle_expect(job), modes=(2, 0), lane_order=shape, force_class=force)
            assert cls == 3


# ---- 16x2 pairs (bsw_warp16.h): two jobs per lane group, half-word cells, per-lane rebasing --------------
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
    return dict(begin_a=int(rng.integers(0, la)), end_a=int(rng.integers(0, la + 30)), begin_b=int(rng.integers(0, lb // 2)),
                end_b=int(rng.integers(lb // 2, lb + 9)), force_start=bool(rng.integers(0, 2)), force_end=bool(rng.integers(0, 2)))


def _check_pairs(jobs, modes=(2, 0), lane_order=0, first_group=0):
    for mode in modes:
        rc, out = simlib.sim_align_pairs(jobs, mode=mode, lane_order=lane_order, first_group=first_group)
        assert rc == len(jobs), rc
        for n, (job, (r, ops)) in enumerate(zip(jobs, out)):
            got = simlib.result_to_expect(r, ops if mode == 2 else None, mode)
            assert got == simlib.project(oracle_expect(job), mode), \
                (mode, n, {k: v for k, v in job.items() if k not in "ab"})


@pytest.mark.parametrize("band", [0, 1, 7, 16, 33, 47, 64, 100, 130, 150, 200, 256, 271, 287])
def test_sim_pairs_every_stripe_width(band):
    """Every geometry the host picks (C = 2..18, LG = 4..32): pairs of different length and shape in one
    group (masked drain per half, per-half capture windows, frozen pos < 0 cells), odd job counts,
    idle groups in front."""
    rng = np.random.default_rng(7000 + band)
    c, lg = simlib.band_geometry(band)
    G = 32 // lg
    for rep in range(3):
        nj = int(rng.integers(1, min(2 * G, 6) + 1))
        first = int(rng.integers(0, G - (nj + 1) // 2 + 1))
        jobs = []
        for g in range(nj):
            length = int(rng.integers(30, 300))
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.03, 0.2])), p_n=0.0)
            b = b[int(rng.integers(0, min(band // 2, len(b) // 4) + 1)):]
            jobs.append(dict(a=a, b=b, band=band, gap=-8, **_shape(rng, a, b, int(rng.integers(0, 4)))))
        _check_pairs(jobs, lane_order=rep, first_group=first)


@pytest.mark.parametrize("band,length", [(64, 1000), (16, 900), (150, 1300), (256, 700)])
def test_sim_pairs_rebase_and_tiles(band, length):
    """Jobs longer than the rebase interval (256 steps) and several sequence tiles; the partner is much
    shorter, so one half idles (keeps its registers) through most rebases."""
    rng = np.random.default_rng(7100 + band)
    a, b = gen.make_pair(rng, length, div=0.03, p_n=0.0)
    a2, b2 = gen.make_pair(rng, 300, div=0.1, p_n=0.0)
    jobs = [dict(a=a, b=b, band=band, gap=-8, **_shape(rng, a, b, 0)),
            dict(a=a2, b=b2, band=band, gap=-8, **_shape(rng, a2, b2, 1))]
    _check_pairs(jobs)
    _check_pairs(jobs[::-1], lane_order=1)


@pytest.mark.parametrize("gap", [-5, -13, -29])
def test_sim_pairs_gap_values_and_extreme_inputs(gap):
    """The range argument of bsw_warp16.h must hold for any input: homopolymers (every cell a match:
    steepest rise), unrelated sequences (steepest fall), a long insertion (scores run along the band
    edge), with the mildest and the harshest gap the fast kernels take."""
    rng = np.random.default_rng(7200 - gap)
    n = 400
    homo = np.zeros(n, dtype=np.uint8)
    r1 = rng.integers(0, 4, n).astype(np.uint8)
    r2 = rng.integers(0, 4, n).astype(np.uint8)
    ins = np.concatenate([r1[:150], rng.integers(0, 4, 60).astype(np.uint8), r1[150:]])
    alt = np.tile(np.array([0, 1], dtype=np.uint8), n // 2)
    cases = [(homo, homo.copy()), (r1, r2), (r1, ins), (ins, r1), (alt, np.roll(alt, 1)), (homo, alt)]
    for band in (20, 64):
        jobs = [dict(a=a, b=b, band=band, gap=gap, **_shape(rng, a, b, 0)) for a, b in cases[:4]]
        _check_pairs(jobs[:2 * (32 // simlib.band_geometry(band)[1])])
        jobs = [dict(a=a, b=b, band=band, gap=gap, **_shape(rng, a, b, 0)) for a, b in cases[2:]]
        _check_pairs(jobs[:2 * (32 // simlib.band_geometry(band)[1])], lane_order=1)


def test_sim_pairs_refuse_n():
    """A window with an N is reported by the pre-scan (the kernel then falls back to the 32-bit body)."""
    rng = np.random.default_rng(7300)
    a, b = gen.make_pair(rng, 200, div=0.02, p_n=0.0)
    a2, b2 = gen.make_pair(rng, 200, div=0.02, p_n=0.0)
    b2 = b2.copy(); b2[57] = 4
    jobs = [dict(a=a, b=b, band=64, gap=-8, **_shape(rng, a, b, 0)), dict(a=a2, b=b2, band=64, gap=-8, **_shape(rng, a2, b2, 0))]
    rc, _ = simlib.sim_align_pairs(jobs, mode=1)
    assert rc == -2
    # an N outside the job's windows does not count
    a3 = np.concatenate([a2, np.array([4, 4, 4], dtype=np.uint8)])
    jobs[1] = dict(a=a3, b=b, band=64, gap=-8, begin_a=0, end_a=100, begin_b=0, end_b=30, force_start=False, force_end=False)
    rc, out = simlib.sim_align_pairs(jobs, mode=1)
    assert rc == 2


@pytest.mark.parametrize("rc", [1, 2, 3])
def test_sim_pairs_reverse_complement_views(rc):
    """The stored contigs hold the reverse complement of what the jobs must see (PctgBuilder.cc:1443):
    the tile staging of bsw_warp16.h reads such views backwards and complements them."""
    rng = np.random.default_rng(7400 + rc)
    jobs, stored = [], []
    for _ in range(4):
        a, b = gen.make_pair(rng, int(rng.integers(100, 500)), div=0.04, p_n=0.0)
        job = dict(a=a, b=b, band=64, gap=-8, **_shape(rng, a, b, int(rng.integers(0, 3))))
        jobs.append(job)
        stored.append(dict(job, a=gen.revcomp(a) if rc & 1 else a, b=gen.revcomp(b) if rc & 2 else b))
    for mode in (2, 0):
        n, out = simlib.sim_align_pairs(stored, mode=mode, rc=rc)
        assert n == len(jobs)
        for job, (r, ops) in zip(jobs, out):
            got = simlib.result_to_expect(r, ops if mode == 2 else None, mode)
            assert got == simlib.project(oracle_expect(job), mode)


@pytest.mark.parametrize("gap", [-3, -8])
def test_generic_body_end_a_wraps(gap):
    """end_a near 2^64 (m_at + mlen - 1 with mlen = 0): the reference has no "last column" cell then
    (int_type(end_a) < 0, banded_smith_waterman.cc:197-212) - the generic body must not invent one out of
    wrapped sums.  Sequences chosen so that every last-row score is negative."""
    rng = np.random.default_rng(77)
    a = rng.integers(0, 4, 90).astype(np.uint8)
    b = ((a[:60] + 1 + rng.integers(0, 3, 60)) % 4).astype(np.uint8)  # mismatch at every diagonal position
    for end_a in (2**64 - 1, 2**64 - 5, 2**63 + 7):
        for begin_a in (0, 3, 40):
            job = dict(a=a, b=b, band=20, gap=gap, begin_a=begin_a, end_a=end_a, begin_b=0, end_b=len(b) - 1,
                       force_start=False, force_end=False)
            exp = oracle_expect(job)
            _check(job, exp, modes=(2, 0), force_class=2)
            _check(job, exp, modes=(2, 0))
