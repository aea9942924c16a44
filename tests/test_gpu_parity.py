"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(include/gamx.h via gam_ngs_b200.capi), against the oracle on the same seeded inputs -
bit-exact scores, coordinates, statuses and every edit op.  Nothing here reads /root/reference:
the checker is the C restatement (oracle/bsw_oracle.c) plus the committed golden vectors."""
import os

import numpy as np
import pytest

import gen
import gam_ngs_b200 as g
from gam_ngs_b200 import capi
from util import load_golden, oracle_expect, x_size_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(devices=[0])
    yield c
    c.close()


def result_to_expect(ctx, r, ops, mode):
    if r["status"] != 0:
        return {"status": int(r["status"])}
    d = {"status": 0, "score": int(r["score"])}
    if mode == capi.MODE_SCORE:
        return d
    hm = int(r["has_match"])
    d.update(begin_a=int(r["begin_a"]), begin_b=int(r["begin_b"]), a_size=int(r["a_size"]),
             b_size=int(r["b_size"]), n_ops=int(r["n_ops"]), homology=float(r["homology"]),
             has_first_match=hm, first_match_a=int(r["first_match_a"]), first_match_b=int(r["first_match_b"]),
             has_last_match=hm, last_match_a=int(r["last_match_a"]), last_match_b=int(r["last_match_b"]),
             has_last_pos=hm, last_pos_a=int(r["last_pos_a"]), last_pos_b=int(r["last_pos_b"]),
             has_gaps=hm, gaps_a=int(r["gaps_a"]), gaps_b=int(r["gaps_b"]))
    if mode == capi.MODE_FULL:
        d["ops"] = bytes(ctx.unpack_ops(ops, int(r["ops_offset"]), int(r["n_ops"])))
    return d


def project(exp, mode):
    if exp["status"] != 0:
        return dict(exp)
    if mode == capi.MODE_SCORE:
        return {"status": 0, "score": exp["score"]}
    d = dict(exp)
    if mode != capi.MODE_FULL:
        d.pop("ops", None)
    return d


def run_batch(ctx, cases, mode, views=None):
    """cases: list of job dicts with raw sequences; returns normalised results."""
    ctx.clear_contigs()
    jobs = g.make_jobs(len(cases))
    for k, c in enumerate(cases):
        jobs[k]["a_id"] = ctx.add_contig(c["a"])
        jobs[k]["b_id"] = ctx.add_contig(c["b"])
        for f in ("begin_a", "end_a", "begin_b", "end_b", "band", "gap"):
            jobs[k][f] = c[f]
        jobs[k]["force_start"], jobs[k]["force_end"] = int(c["force_start"]), int(c["force_end"])
        jobs[k]["mode"] = mode
        if views is not None:
            for f, v in views[k].items():
                jobs[k][f] = v
    res, ops = ctx.align_batch(jobs)
    return [result_to_expect(ctx, res[k], ops, mode) for k in range(len(cases))]


def test_golden_vectors(ctx):
    cases = load_golden()
    jobs = [j for j, _ in cases]
    for mode in (capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE):
        got = run_batch(ctx, jobs, mode)
        for k, (job, exp) in enumerate(cases):
            if x_size_of(job) == 0:
                exp = {"status": 3}
            assert got[k] == project(exp, mode), (k, mode, {a: b for a, b in job.items() if a not in "ab"})


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_fuzz_every_clamp(ctx, seed):
    rng = np.random.default_rng(seed)
    cases, exps = [], []
    while len(cases) < 1500:
        job = gen.fuzz_case(rng)
        x = x_size_of(job)
        if x is not None and x > 3000:
            continue
        cases.append(job)
        exps.append(oracle_expect(job) if x != 0 else {"status": 3})
    for mode in (capi.MODE_FULL, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode, {a: b for a, b in cases[k].items() if a not in "ab"})


@pytest.mark.parametrize("band", [0, 1, 16, 47, 64, 100, 150, 256, 271])
def test_warp_kernel_all_stripe_widths(ctx, band):
    rng = np.random.default_rng(2000 + band)
    cases = []
    for length in (70, 333, 700, 1500):
        for _ in range(6):
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.02, 0.1])), p_n=0.003)
            b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
            la, lb = len(a), len(b)
            shape = int(rng.integers(0, 3))
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                         force_start=False, force_end=True)
            else:
                w = dict(begin_a=3, end_a=la + 40, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 5,
                         force_start=True, force_end=False)
            cases.append(dict(a=a, b=b, band=band, gap=int(rng.choice([-8, -8, -5, -29])), **w))
    exps = [oracle_expect(c) for c in cases]
    for mode in (capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode)


def test_config2_shape_1kb_band64(ctx):
    """BASELINE.json configs[1] shape: 1 kb pairs, band 64, ~2% divergence (+ an N variant)."""
    rng = np.random.default_rng(2)
    cases = []
    for n in range(400):
        a, b = gen.make_pair(rng, 1000, div=0.02, p_n=0.001 if n % 4 == 0 else 0.0)
        cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=64,
                          gap=-8, force_start=False, force_end=False))
    exps = [oracle_expect(c) for c in cases]
    for mode in (capi.MODE_ENDPOINTS, capi.MODE_FULL, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode)


def test_config3_shape_long_band256(ctx):
    """BASELINE.json configs[2] shape at a size the CPU oracle finishes in seconds."""
    rng = np.random.default_rng(3)
    cases = []
    for length in (10000, 17000, 25000):
        a, b = gen.make_pair(rng, length, div=0.02, offset=int(rng.integers(0, 128)))
        cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=256,
                          gap=-8, force_start=False, force_end=False))
    exps = [oracle_expect(c) for c in cases]
    got = run_batch(ctx, cases, capi.MODE_FULL)
    for k in range(len(cases)):
        assert got[k] == exps[k], k
        assert exps[k]["n_ops"] > 9000


def test_views_reverse_complement_and_offsets(ctx):
    rng = np.random.default_rng(5)
    cases, views, exps = [], [], []
    for _ in range(40):
        a, b = gen.make_pair(rng, int(rng.integers(150, 900)), div=0.03, p_n=0.01)
        band = int(rng.choice([20, 64, 150]))
        a_off, b_off = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        a_len = len(a) - a_off - int(rng.integers(0, 20))
        b_len = len(b) - b_off - int(rng.integers(0, 20))
        va, vb = a[a_off:a_off + a_len], b[b_off:b_off + b_len]
        mat = dict(a=va, b=vb, begin_a=0, end_a=len(va) - 1, begin_b=0, end_b=len(vb) - 1, band=band, gap=-8,
                   force_start=False, force_end=False)
        exps.append(oracle_expect(mat))
        cases.append(dict(mat, a=gen.revcomp(a), b=gen.revcomp(b)))  # store holds the reverse complement
        views.append(dict(a_rc=1, b_rc=1, a_off=a_off, a_len=a_len, b_off=b_off, b_len=b_len))
    got = run_batch(ctx, cases, capi.MODE_FULL, views)
    for k in range(len(cases)):
        assert got[k] == exps[k], k


def test_traceback_round_trip_properties_at_scale(ctx):
    """Size-independent properties on a larger batch than the oracle is asked to check:
    the edit string must consume exactly the aligned spans and re-score to the DP score."""
    rng = np.random.default_rng(9)
    ctx.clear_contigs()
    n = 3000
    jobs = g.make_jobs(n)
    seqs = []
    for k in range(n):
        a, b = gen.make_pair(rng, 1000, div=0.02)
        seqs.append((a, b))
        jobs[k]["a_id"], jobs[k]["b_id"] = ctx.add_contig(a), ctx.add_contig(b)
        jobs[k]["end_a"], jobs[k]["end_b"] = len(a) - 1, len(b) - 1
        jobs[k]["band"], jobs[k]["mode"] = 64, capi.MODE_FULL
    res, ops = ctx.align_batch(jobs)
    S = np.array([[5, -4, -4, -4, 0], [-4, 5, -4, -4, 0], [-4, -4, 5, -4, 0], [-4, -4, -4, 5, 0], [0, 0, 0, 0, 5]])
    assert (res["status"] == 0).all()
    for k in range(0, n, 7):
        r = res[k]
        o = ctx.unpack_ops(ops, int(r["ops_offset"]), int(r["n_ops"]))
        a, b = seqs[k]
        n_diag = int(((o == 2) | (o == 3)).sum())
        n_ga, n_gb = int((o == 0).sum()), int((o == 1).sum())
        assert n_ga == r["n_gap_a"] and n_gb == r["n_gap_b"] and int((o == 2).sum()) == r["n_match"]
        # spans: a advances on diag+GAP_B, b on diag+GAP_A; the path ends at the selected end cell
        end_a_pos = int(r["end_i"]) + int(r["end_j"]) - 64
        assert int(r["begin_a"]) + n_diag + n_gb == end_a_pos + 1
        assert int(r["begin_b"]) + n_diag + n_ga == int(r["end_i"]) + 1
        # re-score: every op costs its face value, except that GAP_B moves inside DP row 0 are
        # free (the first row takes its left neighbour without the gap penalty, .cc:120)
        pa, pb, sc = int(r["begin_a"]), int(r["begin_b"]), 0
        for op in o:
            if op >= 2:
                sc += S[a[pa], b[pb]]; pa += 1; pb += 1
            elif op == 0:
                sc -= 8; pb += 1
            else:
                sc -= 0 if pb == 1 and int(r["begin_b"]) == 0 else 8
                pa += 1
        if o[0] >= 2:
            assert sc == r["score"], k


def test_python_mirror_interface(ctx):
    """The Python mirror keeps the reference's call shape and error behaviour."""
    rng = np.random.default_rng(4)
    a, b = gen.make_pair(rng, 400, div=0.02)
    A, B = g.Contig(a), g.Contig(b)
    sw = g.BandedSmithWaterman(ctx=ctx)  # band 150
    al = sw.find_alignment(A, 0, A.size() - 1, B, 0, B.size() - 1)
    exp = oracle_expect(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=150,
                             gap=-8, force_start=False, force_end=False))
    assert (al.begin_a(), al.begin_b(), al.score(), al.length()) == (exp["begin_a"], exp["begin_b"], exp["score"], exp["n_ops"])
    assert bytes(al.sequence()) == exp["ops"] and al.homology() == exp["homology"]
    # default MyAlignment for an inverted b window (.cc:90)
    e = sw.find_alignment(A, 0, 10, B, 5, 4)
    assert (e.begin_a(), e.a_size(), e.length(), e.score()) == (0, 0, 0, 0)
    # std::out_of_range when the chosen end cell lies beyond |a|
    all_a, all_c = g.Contig("A" * 50), g.Contig("C" * 50)
    with pytest.raises(IndexError):
        g.BandedSmithWaterman(10, ctx=ctx).find_alignment(all_a, 0, 60, all_c, 0, 49)


def test_cpp_dropin_runs_on_gpu(tmp_path):
    """The C++ drop-in class (gam_ngs_b200/cpp/gamx_dropin.hpp) called like PctgBuilder.cc:1669
    returns the oracle's answer (checked inside tests/cpp/test_dropin.cc)."""
    import subprocess
    from test_abi import _build_dropin
    exe = _build_dropin(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


@pytest.mark.parametrize("band", [288, 400, 512, 1024, 1500])
def test_cta_per_pair_kernel_wide_bands(ctx, band):
    """K2 (one pair per CTA) handles bands wider than a warp's stripes (BASELINE config 5 sweeps
    the band up to 1024)."""
    rng = np.random.default_rng(5000 + band)
    cases = []
    for length in (150, 700, 2500):
        for _ in range(4):
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.02, 0.1])), p_n=0.003)
            b = b[int(rng.integers(0, min(band // 2, length // 4) + 1)):]
            la, lb = len(a), len(b)
            shape = int(rng.integers(0, 3))
            if shape == 0:
                w = dict(begin_a=0, end_a=la - 1, begin_b=0, end_b=lb - 1, force_start=False, force_end=False)
            elif shape == 1:
                w = dict(begin_a=int(rng.integers(0, la // 2)), end_a=la - 1, begin_b=0, end_b=lb - 1,
                         force_start=False, force_end=True)
            else:
                w = dict(begin_a=3, end_a=la + 40, begin_b=int(rng.integers(0, lb // 2)), end_b=lb + 5,
                         force_start=True, force_end=False)
            cases.append(dict(a=a, b=b, band=band, gap=-8, **w))
    exps = [oracle_expect(c) for c in cases]
    for mode in (capi.MODE_FULL, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode)


def test_cta_per_pair_latency_mode_for_few_long_pairs(ctx):
    """A handful of long overlaps (too few to fill 148 SMs with one warp each) is routed to the
    CTA-per-pair kernel; results must not depend on which kernel ran."""
    rng = np.random.default_rng(77)
    cases = []
    for length, band in [(6000, 150), (9000, 256), (4000, 64), (12000, 150)]:
        a, b = gen.make_pair(rng, length, div=0.02, offset=int(rng.integers(0, band // 2)))
        cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=band, gap=-8,
                          force_start=False, force_end=False))
    exps = [oracle_expect(c) for c in cases]
    got = run_batch(ctx, cases, capi.MODE_FULL)
    for k in range(len(cases)):
        assert got[k] == exps[k], k


def test_find_hits_matches_oracle(ctx):
    """f3: ABlast::findHits on the GPU (k-mer diagonal voting) vs the restatement pinned to the
    reference: list length, front() and back() of the hit list, incl. N aliasing and tied diagonals."""
    import oracle
    from test_oracle import _hits_case
    rst = oracle.restatement()
    rng = np.random.default_rng(43)
    ctx.clear_contigs()
    cases = [_hits_case(rng) for _ in range(600)]
    # a few larger, realistic tail windows
    for _ in range(6):
        a = gen.random_seq(rng, int(rng.integers(3000, 9000)), 0.001)
        s = int(rng.integers(0, len(a) // 2))
        b = gen.mutate(rng, a[s:s + int(rng.integers(500, 2500))], div=0.02)
        cases.append((a, b, (0, len(a) - 1, 0, len(b) - 1)))
    jobs = g.make_hits_jobs(len(cases))
    for k, (a, b, w) in enumerate(cases):
        jobs[k]["a_id"], jobs[k]["b_id"] = ctx.add_contig(a), ctx.add_contig(b)
        jobs[k]["a_start"], jobs[k]["a_end"], jobs[k]["b_start"], jobs[k]["b_end"] = w
    res = ctx.find_hits_batch(jobs)
    nonempty = ties = 0
    for k, (a, b, w) in enumerate(cases):
        want, mc = rst.find_hits(a, w[0], w[1], b, w[2], w[3])
        assert res[k]["n_hits"] == len(want), k
        if len(want):
            assert (res[k]["first_hit"], res[k]["last_hit"], res[k]["max_count"]) == (want[0], want[-1], mc), k
            nonempty += 1
            ties += len(want) > 1
    assert nonempty > 150 and ties > 20
    # reverse-complement views: the store holds rc(a); the job asks for view rc -> a
    a, b, w = cases[-1]
    ctx.clear_contigs()
    j = g.make_hits_jobs(1)
    j[0]["a_id"], j[0]["b_id"] = ctx.add_contig(gen.revcomp(a)), ctx.add_contig(b)
    j[0]["a_rc"] = 1
    j[0]["a_start"], j[0]["a_end"], j[0]["b_start"], j[0]["b_end"] = w
    r = ctx.find_hits_batch(j)[0]
    want, mc = rst.find_hits(a, w[0], w[1], b, w[2], w[3])
    assert (r["n_hits"], r["first_hit"], r["last_hit"]) == (len(want), want[0], want[-1])


def test_pipelined_batch_matches_single_launch_and_oracle():
    """The chunked, pipelined gamx_align_batch (asynchronous piece-wise contig upload from pinned
    memory, two slots/streams, helper-thread planning) returns exactly what the single-launch path
    returns, and both agree with the oracle; includes early-exit and generic-kernel jobs."""
    import torch
    rng = np.random.default_rng(11)
    n = 3000
    a, al, b, bl = gen.bulk_pairs(rng, n, 0, div=0.03, len_lo=150, len_hi=700)
    # chunk-interleaved contig order [A chunk][B chunk]... as bench.py's end-to-end leg uploads it
    m = 256
    ao = np.concatenate([[0], np.cumsum(al)]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.int64)
    host = torch.empty(int(ao[-1] + bo[-1]), dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    lengths, a_id, b_id = [], np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    pos = cid = 0
    for lo in range(0, n, m):
        hi = min(n, lo + m)
        for src, off, ln, ids in ((a, ao, al, a_id), (b, bo, bl, b_id)):
            seg = src[off[lo]:off[hi]]
            hv[pos:pos + len(seg)] = seg
            pos += len(seg)
            lengths.extend(ln[lo:hi].tolist())
            ids[lo:hi] = np.arange(cid, cid + hi - lo)
            cid += hi - lo
    jobs = g.make_jobs(n)
    jobs["a_id"], jobs["b_id"] = a_id, b_id
    jobs["end_a"], jobs["end_b"] = al - 1, bl - 1
    jobs["band"] = rng.choice([16, 64, 150], size=n)
    jobs["force_end"] = rng.integers(0, 8, size=n) == 0
    jobs["mode"] = np.where(rng.integers(0, 4, size=n) == 0, capi.MODE_SCORE, capi.MODE_ENDPOINTS)
    jobs["end_b"][5] = 0; jobs["begin_b"][5] = 3     # empty (end_b < begin_b)
    jobs["gap"][9] = -3                               # generic kernel
    import os
    os.environ["GAMX_UPLOAD_PIECE_BYTES"] = "200000"  # several upload pieces even at this size (read by gamx_create)
    c = g.Context(devices=[0])
    del os.environ["GAMX_UPLOAD_PIECE_BYTES"]
    try:
        c.set_pipeline_chunk(0)
        c.add_contigs(host.data_ptr(), np.array(lengths, dtype=np.uint64))
        ref, _ = c.align_batch(jobs)
        for chunk in (128, 1000):
            c.clear_contigs()
            c.set_pipeline_chunk(chunk)
            c.add_contigs(host.data_ptr(), np.array(lengths, dtype=np.uint64), async_upload=True)
            got, _ = c.align_batch(jobs)
            assert got.tobytes() == ref.tobytes(), chunk
        # an edit string requested somewhere in the batch: the pipelined path hands the whole batch
        # back to the single-plan path
        jf = jobs.copy()
        jf["mode"][2000] = capi.MODE_FULL
        c.set_pipeline_chunk(0)
        want, wops = c.align_batch(jf)
        c.set_pipeline_chunk(128)
        got, gops = c.align_batch(jf)
        assert got.tobytes() == want.tobytes()
        assert bytes(c.unpack_ops(gops, int(got[2000]["ops_offset"]), int(got[2000]["n_ops"]))) == \
            bytes(c.unpack_ops(wops, int(want[2000]["ops_offset"]), int(want[2000]["n_ops"])))
        assert int(got[2000]["n_ops"]) > 100
    finally:
        c.close()
    assert int(ref["status"][5]) == capi.JOB_EMPTY
    for k in list(range(0, n, 97)) + [9]:
        A, B = a[ao[k]:ao[k + 1]], b[bo[k]:bo[k + 1]]
        case = dict(a=A, b=B, begin_a=0, end_a=len(A) - 1, begin_b=int(jobs["begin_b"][k]), end_b=int(jobs["end_b"][k]),
                    band=int(jobs["band"][k]), gap=int(jobs["gap"][k]), force_start=False, force_end=bool(jobs["force_end"][k]))
        mode = int(jobs["mode"][k])
        assert result_to_expect(None, ref[k], None, mode) == project(oracle_expect(case), mode), k


def test_many_waves_odd_wave_sizes():
    """Direction scratch forced down to a few hundred KB per half (GAMX_DIRS_HALF_BYTES): a group then
    takes many waves with odd job counts, alternating between the two halves and the two fill streams,
    with the traceback kernel of one wave running beside the fill of the next.  An idle lane group of a
    wave's last warp must not touch the neighbouring half."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import gen, gam_ngs_b200 as g
from gam_ngs_b200 import capi
rng = np.random.default_rng(5)
n = 1501
a, al, b, bl = gen.bulk_pairs(rng, n, 0, div=0.03, len_lo=300, len_hi=900)
ctx = g.Context(devices=[0])
ctx.add_contigs(np.concatenate([a, b]), np.concatenate([al, bl]))
jobs = g.make_jobs(n)
jobs["a_id"] = np.arange(n); jobs["b_id"] = np.arange(n, 2 * n)
jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = 64; jobs["mode"] = capi.MODE_FULL
res, ops = ctx.align_batch(jobs)
np.save(sys.argv[1], res)
np.save(sys.argv[1] + ".ops", np.array([bytes(ctx.unpack_ops(ops, int(r["ops_offset"]), int(r["n_ops"]))).hex() for r in res]))
'''
    outs = []
    for tag, env in (("big", {}), ("small", {"GAMX_DIRS_HALF_BYTES": str(37 * 33 * 1024)})):
        path = f"/tmp/gamx_waves_{tag}.npy"
        e = dict(os.environ, **env)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=e,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        outs.append((np.load(path), np.load(path + ".ops.npy")))
    big, small = outs
    cols = [f for f in big[0].dtype.names if f != "ops_offset"]
    for f in cols:
        assert (big[0][f] == small[0][f]).all(), f
    assert (big[1] == small[1]).all()
    assert int((big[0]["status"] == 0).sum()) == len(big[0])


def test_multi_device_context_matches_single_device():
    """One context over two GPUs (the library's own cost-balanced sharding, host gather, no
    collective): same results as one device, for a mixed batch (several kernel families, all modes)
    and for a pipelined batch.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(21)
    n = 2400
    lens = np.exp(rng.uniform(np.log(200), np.log(6000), size=n)).astype(np.int64)
    a, al, b, bl = gen.bulk_pairs(rng, n, 0, div=0.03, lengths=lens)
    jobs = g.make_jobs(n)
    jobs["a_id"] = np.arange(n); jobs["b_id"] = np.arange(n, 2 * n)
    jobs["end_a"], jobs["end_b"] = al - 1, bl - 1
    jobs["band"] = rng.choice([16, 64, 150, 300], size=n)
    jobs["mode"] = rng.choice([capi.MODE_SCORE, capi.MODE_ENDPOINTS, capi.MODE_FULL], size=n)
    outs = []
    for devs in ([0], [0, 1]):
        c = g.Context(devices=devs)
        try:
            c.add_contigs(np.concatenate([a, b]), np.concatenate([al, bl]))
            res, ops = c.align_batch(jobs)
            strings = [bytes(c.unpack_ops(ops, int(r["ops_offset"]), int(r["n_ops"]))) if m == capi.MODE_FULL else b""
                       for r, m in zip(res, jobs["mode"])]
            j2 = jobs.copy()
            j2["mode"] = np.where(j2["mode"] == capi.MODE_FULL, capi.MODE_ENDPOINTS, j2["mode"])
            c.set_pipeline_chunk(300)
            piped, _ = c.align_batch(j2)
            outs.append((res, strings, piped))
        finally:
            c.close()
    (r1, s1, p1), (r2, s2, p2) = outs
    for f in r1.dtype.names:
        if f != "ops_offset":
            assert (r1[f] == r2[f]).all(), f
    assert s1 == s2
    assert p1.tobytes() == p2.tobytes()
    ep = jobs["mode"] == capi.MODE_ENDPOINTS
    for f in r1.dtype.names:
        if f != "ops_offset":  # (meaningless without an edit string)
            assert (r1[ep][f] == p1[ep][f]).all(), f


def test_config2_full_size_properties():
    """BASELINE.json configs[1] at its full size (1,000,000 pairs of 1 kb, band 64, score+endpoints):
    three independent product paths - resident plan (two waves), single-launch gamx_align_batch and the
    pipelined gamx_align_batch behind an asynchronous piece-wise upload - must agree bit for bit; every
    alignment must satisfy the size-independent invariants of a banded overlap alignment; a seeded
    sample is checked against the oracle."""
    import torch
    n = 1_000_000
    rng = np.random.default_rng(2024)
    a, al, b, bl = gen.bulk_pairs(rng, n, 1000, div=0.02)
    ao = np.concatenate([[0], np.cumsum(al)]).astype(np.int64)
    bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.int64)
    host = torch.empty(len(a) + len(b), dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    lengths = np.empty(2 * n, dtype=np.uint64)
    a_id = np.empty(n, np.uint32); b_id = np.empty(n, np.uint32)
    pos = cid = 0
    for lo in range(0, n, 16384):
        hi = min(n, lo + 16384)
        for src, off, ln, ids in ((a, ao, al, a_id), (b, bo, bl, b_id)):
            seg = src[off[lo]:off[hi]]
            hv[pos:pos + len(seg)] = seg; pos += len(seg)
            lengths[cid:cid + hi - lo] = ln[lo:hi]
            ids[lo:hi] = np.arange(cid, cid + hi - lo); cid += hi - lo
    jobs = g.make_jobs(n)
    jobs["a_id"], jobs["b_id"] = a_id, b_id
    jobs["end_a"], jobs["end_b"] = al - 1, bl - 1
    jobs["band"] = 64
    jobs["mode"] = capi.MODE_ENDPOINTS
    c = g.Context(devices=[0])
    try:
        c.add_contigs(host.data_ptr(), lengths)
        plan = c.plan(jobs); plan.run(); plan.sync()
        ref, _ = plan.fetch()
        plan.close()
        c.set_pipeline_chunk(0)
        single, _ = c.align_batch(jobs)
        assert single.tobytes() == ref.tobytes()
        c.set_pipeline_chunk(65536)
        c.clear_contigs()
        c.add_contigs(host.data_ptr(), lengths, async_upload=True)
        piped, _ = c.align_batch(jobs)
        assert piped.tobytes() == ref.tobytes()
    finally:
        c.close()
    r = ref
    assert (r["status"] == 0).all()
    m, x, ga, gb = (r[f].astype(np.int64) for f in ("n_match", "n_mismatch", "n_gap_a", "n_gap_b"))
    assert (r["n_ops"].astype(np.int64) == m + x + ga + gb).all()
    end_a_pos = r["end_i"] + r["end_j"] - 64                     # a-position of the end cell
    assert (r["begin_a"].astype(np.int64) + m + x + gb == end_a_pos + 1).all()
    assert (r["begin_b"].astype(np.int64) + m + x + ga == r["end_i"] + 1).all()
    face = 5 * m - 4 * x - 8 * (ga + gb)                          # no N in this workload
    assert (r["score"] >= face).all() and (r["score"] <= face + 8 * gb).all()  # (row-0 GAP_B moves are free, .cc:120)
    # overlap alignments end in the last row or in the last column
    assert ((r["end_i"] == bl.astype(np.int64) - 1) | (end_a_pos == al.astype(np.int64) - 1)).all()
    assert (r["homology"] > 90.0).mean() > 0.999                 # 2 % divergence
    for k in rng.choice(n, size=120, replace=False):
        A, B = a[ao[k]:ao[k + 1]], b[bo[k]:bo[k + 1]]
        case = dict(a=A, b=B, begin_a=0, end_a=len(A) - 1, begin_b=0, end_b=len(B) - 1, band=64, gap=-8,
                    force_start=False, force_end=False)
        assert result_to_expect(None, r[k], None, capi.MODE_ENDPOINTS) == project(oracle_expect(case), capi.MODE_ENDPOINTS), k


def test_gpu_matches_compiled_reference_directly(ctx):
    """No restatement in between: the CUDA path against oracle/_ref/libgamref.so (the UNMODIFIED reference aligner;
    the prebuilt library travels to the GPU box) on fuzz cases of every clamp plus config-2 / config-3 shaped
    pairs, all three modes."""
    import oracle
    if not oracle.reference_available():
        pytest.skip("reference build not present")
    ref = oracle.reference()
    rng = np.random.default_rng(4242)
    cases = []
    while len(cases) < 240:
        job = gen.fuzz_case(rng)
        x = x_size_of(job)
        if x is None or x == 0 or x > 3000:
            continue  # (x == 0: the reference's behaviour is undefined, .cc:102-122)
        cases.append(job)
    for length, band in ((1000, 64), (700, 16), (2500, 256), (1500, 150)):
        for _ in range(8):
            a, b = gen.make_pair(rng, length, div=float(rng.choice([0.0, 0.02, 0.08])), p_n=float(rng.choice([0.0, 0.002])))
            cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=band, gap=-8,
                              force_start=False, force_end=False))
    exps = [oracle_expect(c, ref) for c in cases]
    for mode in (capi.MODE_FULL, capi.MODE_ENDPOINTS, capi.MODE_SCORE):
        got = run_batch(ctx, cases, mode)
        for k in range(len(cases)):
            assert got[k] == project(exps[k], mode), (k, mode, {a: b for a, b in cases[k].items() if a not in "ab"})


def test_device_cigar_matches_host_rle(ctx):
    """gamx_align_batch_cigar: the run-length CIGARs built on the device (cigar_count_kernel / scan /
    cigar_emit_kernel) equal gamx_cigar_rle of the packed edit strings of gamx_align_batch and the runs of the
    oracle's edit strings; jobs of other modes and empty alignments have no runs."""
    rng = np.random.default_rng(515)
    cases = []
    for length, band in ((60, 7), (333, 16), (1000, 64), (1500, 150), (2500, 256), (5000, 512)):
        for _ in range(6):
            a, b = gen.make_pair(rng, length + int(rng.integers(0, 50)), div=float(rng.choice([0.0, 0.02, 0.1])), p_n=0.002)
            cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=band, gap=-8,
                              force_start=False, force_end=False))
    for _ in range(40):
        cases.append(gen.fuzz_case(rng))
    ctx.clear_contigs()
    jobs = g.make_jobs(len(cases))
    for k, c in enumerate(cases):
        jobs[k]["a_id"] = ctx.add_contig(c["a"]); jobs[k]["b_id"] = ctx.add_contig(c["b"])
        for f in ("begin_a", "end_a", "begin_b", "end_b", "band", "gap"):
            jobs[k][f] = c[f]
        jobs[k]["force_start"], jobs[k]["force_end"] = int(c["force_start"]), int(c["force_end"])
        jobs[k]["mode"] = capi.MODE_FULL if k % 7 != 3 else capi.MODE_ENDPOINTS
    res, ops = ctx.align_batch(jobs)
    res2, offs, runs = ctx.align_batch_cigar(jobs)
    n_runs = 0
    for k, c in enumerate(cases):
        assert res2[k]["status"] == res[k]["status"] and res2[k]["score"] == res[k]["score"] and res2[k]["n_ops"] == res[k]["n_ops"]
        mine = [(int(r & 3), int(r >> 2)) for r in runs[int(offs[k]):int(offs[k + 1])]]
        if res[k]["status"] != 0 or jobs[k]["mode"] != capi.MODE_FULL:
            assert mine == []
            continue
        assert mine == ctx.cigar_rle(ops, int(res[k]["ops_offset"]), int(res[k]["n_ops"])), k
        exp = oracle_expect(c) if x_size_of(c) != 0 else {"status": 3}
        if exp["status"] == 0:
            want, prev = [], None
            for op in exp["ops"]:
                if prev == op:
                    want[-1][1] += 1
                else:
                    want.append([op, 1]); prev = op
            assert mine == [tuple(w) for w in want], k
        assert sum(l for _, l in mine) == int(res[k]["n_ops"])
        n_runs += len(mine)
    assert n_runs > 500


def test_fasta_loader(ctx, tmp_path):
    """gamx_add_fasta reads records like the reference's reader (io_contig.code.hpp:540-565: everything but
    newline, blank and '>' is a base; nucleotide.code.hpp:47-75: unknown characters are N)."""
    rng = np.random.default_rng(99)
    seqs = [gen.random_seq(rng, n, p_n=0.01) for n in (1, 70, 1000, 4097)]
    txt = ""
    for k, s_ in enumerate(seqs):
        letters = "".join("ATCGN"[c] for c in s_)
        if k == 1:
            letters = letters.lower()
        if k == 2:
            letters = letters[:500] + "RYK" + letters[500:]     # IUPAC codes -> N
        txt += f">ctg{k} some description\n"
        txt += "\n".join(letters[i:i + 60] for i in range(0, len(letters), 60)) + "\n"
    path = tmp_path / "t.fa"
    path.write_text(txt)
    ctx.clear_contigs()
    first, n = ctx.add_fasta(str(path))
    assert n == 4 and [ctx.contig_name(first + k) for k in range(4)] == ["ctg0", "ctg1", "ctg2", "ctg3"]
    want = [s_.copy() for s_ in seqs]
    want[2] = np.concatenate([seqs[2][:500], np.array([4, 4, 4], dtype=np.uint8), seqs[2][500:]])
    assert [ctx.contig_length(first + k) for k in range(4)] == [len(w) for w in want]
    # the stored bases: align every contig against an uploaded copy of what it should be
    jobs = g.make_jobs(4)
    for k in range(4):
        jobs[k]["a_id"] = first + k
        jobs[k]["b_id"] = ctx.add_contig(want[k])
        jobs[k]["end_a"] = jobs[k]["end_b"] = len(want[k]) - 1
        jobs[k]["band"] = 8; jobs[k]["mode"] = capi.MODE_ENDPOINTS
    res, _ = ctx.align_batch(jobs)
    for k in range(4):
        nn = int((want[k] == 4).sum())
        assert res[k]["status"] == 0 and res[k]["n_ops"] == len(want[k]) and res[k]["n_match"] == len(want[k]), k
        assert res[k]["score"] == 5 * len(want[k])  # (N against N scores 5 like a match, .cc:86)


def test_plan_invalidated_by_a_later_batch_and_async_upload_contract(ctx):
    """(1) A plan whose slot buffers were taken over by a later plan or batch refuses to run / fetch instead of
    returning another batch's results.  (2) gamx_add_contigs_async: the caller's buffer may be reused as soon as
    the next batch call returns - also when that batch is small, single-plan and touches only the first contigs."""
    import torch
    rng = np.random.default_rng(31)
    cases = []
    for _ in range(6):
        a, b = gen.make_pair(rng, 400, div=0.02)
        cases.append(dict(a=a, b=b, begin_a=0, end_a=len(a) - 1, begin_b=0, end_b=len(b) - 1, band=32, gap=-8,
                          force_start=False, force_end=False))
    ctx.clear_contigs()
    jobs = g.make_jobs(len(cases))
    for k, c in enumerate(cases):
        jobs[k]["a_id"] = ctx.add_contig(c["a"]); jobs[k]["b_id"] = ctx.add_contig(c["b"])
        jobs[k]["end_a"], jobs[k]["end_b"], jobs[k]["band"], jobs[k]["mode"] = len(c["a"]) - 1, len(c["b"]) - 1, 32, capi.MODE_ENDPOINTS
    p1 = ctx.plan(jobs[:3])
    p2 = ctx.plan(jobs[3:])          # takes the slot over
    with pytest.raises(g.GamxError):
        p1.run()
    with pytest.raises(g.GamxError):
        p1.fetch()
    p2.run(); p2.sync()
    r2, _ = p2.fetch()
    want = [project(oracle_expect(c), capi.MODE_ENDPOINTS) for c in cases[3:]]
    assert [result_to_expect(ctx, r2[k], None, capi.MODE_ENDPOINTS) for k in range(3)] == want
    ctx.align_batch(jobs)            # a batch invalidates the plan as well
    with pytest.raises(g.GamxError):
        p2.run()
    p1.close(); p2.close()

    # (2) many contigs (several upload pieces with a small piece size are not needed: the contract is about
    # pieces that no job of the batch refers to), a batch that touches only the first pair, then the buffer dies
    n = 3000
    a, al, b, bl = gen.bulk_pairs(rng, n, 700)
    host = torch.empty(len(a) + len(b), dtype=torch.uint8, pin_memory=True)
    host.numpy()[:len(a)] = a; host.numpy()[len(a):] = b
    lengths = np.concatenate([al, bl])
    ctx.clear_contigs()
    first = ctx.add_contigs(host.data_ptr(), lengths, async_upload=True)
    j = g.make_jobs(1)
    j["a_id"], j["b_id"], j["end_a"], j["end_b"], j["band"], j["mode"] = first, first + n, al[0] - 1, bl[0] - 1, 64, capi.MODE_SCORE
    ctx.align_batch(j)
    host.numpy()[:] = 4              # the caller reuses its buffer: everything must have been copied by now
    jj = g.make_jobs(n)
    jj["a_id"] = first + np.arange(n); jj["b_id"] = first + n + np.arange(n)
    jj["end_a"] = al - 1; jj["end_b"] = bl - 1; jj["band"] = 64; jj["mode"] = capi.MODE_SCORE
    res, _ = ctx.align_batch(jj)
    ao = np.concatenate([[0], np.cumsum(al)]).astype(np.int64); bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.int64)
    for k in (0, 1, n // 2, n - 2, n - 1):
        c = dict(a=a[ao[k]:ao[k + 1]], b=b[bo[k]:bo[k + 1]], begin_a=0, end_a=int(al[k]) - 1, begin_b=0, end_b=int(bl[k]) - 1,
                 band=64, gap=-8, force_start=False, force_end=False)
        assert int(res[k]["score"]) == oracle_expect(c)["score"], k


def test_throughput_paths_on_small_batches():
    """Small launches are routed for latency: a group of fewer than 2 x SM-count pairs whose longest pair has
    1024+ rows runs on the CTA-per-pair kernel, and a launch of at most 8192 jobs walks every traceback on a warp.
    The tests above therefore see the warp-level pair kernels of 32-lane stripes and the thread-per-job traceback
    mostly through their big batches.  Here the golden vectors, every stripe width and one fuzz seed run again in a
    process that has both routings switched off (the switches are read once per process), so the throughput
    kernels face the same small adversarial cases."""
    import subprocess
    import sys
    if os.environ.get("GAMX_TEST_INNER"):
        pytest.skip("inner run")
    env = dict(os.environ, GAMX_NO_LATENCY_MODE="1", GAMX_NO_TB_ALL_WARP="1", GAMX_TEST_INNER="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "golden_vectors or all_stripe_widths or fuzz_every_clamp and 21 or reference_directly"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(here)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1000:]
