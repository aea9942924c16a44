"""ctypes wrapper of the CPU lane simulator (tests/sim/warp_sim.cc) - test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class GamxResult(C.Structure):
    """include/gamx.h: gamx_result"""
    _fields_ = [
        ("status", C.c_int32), ("has_match", C.c_int32), ("score", C.c_int64),
        ("begin_a", C.c_uint64), ("begin_b", C.c_uint64), ("a_size", C.c_uint64), ("b_size", C.c_uint64),
        ("n_ops", C.c_uint64), ("n_match", C.c_uint64),
        ("n_mismatch", C.c_uint64), ("n_gap_a", C.c_uint64), ("n_gap_b", C.c_uint64),
        ("homology", C.c_double),
        ("first_match_a", C.c_uint64), ("first_match_b", C.c_uint64),
        ("last_match_a", C.c_uint64), ("last_match_b", C.c_uint64),
        ("last_pos_a", C.c_uint64), ("last_pos_b", C.c_uint64),
        ("gaps_a", C.c_uint64), ("gaps_b", C.c_uint64),
        ("end_i", C.c_int64), ("end_j", C.c_int64),
        ("x_size", C.c_uint64), ("ops_offset", C.c_uint64),
    ]


def result_to_expect(r, ops=None, mode=2):
    """Normalise a gamx_result like util.oracle_expect does for the oracle."""
    if r.status != 0:
        return {"status": r.status}
    d = {"status": 0, "score": r.score}
    if mode == 0:
        return d
    d.update(begin_a=r.begin_a, begin_b=r.begin_b, a_size=r.a_size, b_size=r.b_size, n_ops=r.n_ops,
             homology=r.homology,
             has_first_match=int(r.has_match), first_match_a=r.first_match_a, first_match_b=r.first_match_b,
             has_last_match=int(r.has_match), last_match_a=r.last_match_a, last_match_b=r.last_match_b,
             has_last_pos=int(r.has_match), last_pos_a=r.last_pos_a, last_pos_b=r.last_pos_b,
             has_gaps=int(r.has_match), gaps_a=r.gaps_a, gaps_b=r.gaps_b)
    if ops is not None:
        d["ops"] = bytes(ops[: r.n_ops])
    return d


def project(exp, mode):
    """Reduce an oracle expectation to what a mode reports."""
    if exp["status"] != 0:
        return dict(exp)
    if mode == 0:
        return {"status": 0, "score": exp["score"]}
    d = dict(exp)
    if mode == 1:
        d.pop("ops", None)
    return d


_lib = None


def lib():
    global _lib
    if _lib is None:
        out = os.path.join(HERE, "_build", "libwarpsim.so")
        src = os.path.join(HERE, "sim", "warp_sim.cc")
        deps = [src] + [os.path.join(ROOT, "gam_ngs_b200", "csrc", f) for f in
                        ("bsw_common.h", "bsw_warp.h", "bsw_warp16.h", "bsw_generic.h", "bsw_traceback.h", "bsw_host.h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            # twelve translation units in parallel (warp_sim.cc: SIM_PART), then one link
            flags = ["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-fPIC"]
            objs = [os.path.join(HERE, "_build", f"warp_sim_{p}.o") for p in range(12)]
            procs = [subprocess.Popen(flags + [f"-DSIM_PART={p}", "-c", "-o", o, src]) for p, o in enumerate(objs)]
            if any(pr.wait() != 0 for pr in procs):
                raise RuntimeError("simulator build failed")
            subprocess.run(["g++", "-shared", "-o", out] + objs, check=True)
        _lib = C.CDLL(out)
        u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
        _lib.sim_align.argtypes = [u8p, u64, C.c_int, u64, u64, u8p, u64, C.c_int, u64, u64,
                                   u64, u64, u64, u64, u64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(GamxResult), u8p, u64]
        _lib.sim_align.restype = C.c_int
    return _lib


U64_MAX = 2**64 - 1


def sim_align(job, mode=2, force_class=0, lane_order=0, a_rc=0, a_off=0, a_len=U64_MAX, b_rc=0,
              b_off=0, b_len=U64_MAX):
    a = np.ascontiguousarray(job["a"], dtype=np.uint8)
    b = np.ascontiguousarray(job["b"], dtype=np.uint8)
    r = GamxResult()
    cap = len(a) + len(b) + 2 * job["band"] + 64
    ops = np.zeros(cap, dtype=np.uint8)
    u8p = C.POINTER(C.c_uint8)
    cls = lib().sim_align(a.ctypes.data_as(u8p), len(a), a_rc, a_off, a_len,
                          b.ctypes.data_as(u8p), len(b), b_rc, b_off, b_len,
                          job["begin_a"], job["end_a"], job["begin_b"], job["end_b"], job["band"],
                          job["gap"], int(job["force_start"]), int(job["force_end"]), mode,
                          force_class, lane_order, C.byref(r), ops.ctypes.data_as(u8p), cap)
    return cls, r, ops


def sim_align_multi(jobs, mode=2, lane_order=0):
    """Runs up to 4 regular jobs (same band and gap) in ONE simulated warp, one per lane group.
    Returns [(result, ops) or None for jobs that did not run on the warp kernel]."""
    L = lib()
    if not hasattr(L, "_multi_ready"):
        L.sim_align_multi.restype = C.c_int
        L._multi_ready = True
    n = len(jobs)
    u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
    keep = []
    ap, bp, op = (u8p * n)(), (u8p * n)(), (u8p * n)()
    la, lb, ba, ea, bb, eb = [(u64 * n)() for _ in range(6)]
    fs, fe = (C.c_int * n)(), (C.c_int * n)()
    cap = 0
    for k, j in enumerate(jobs):
        a = np.ascontiguousarray(j["a"], dtype=np.uint8); b = np.ascontiguousarray(j["b"], dtype=np.uint8)
        keep += [a, b]
        ap[k], bp[k] = a.ctypes.data_as(u8p), b.ctypes.data_as(u8p)
        la[k], lb[k], ba[k], ea[k], bb[k], eb[k] = len(a), len(b), j["begin_a"], j["end_a"], j["begin_b"], j["end_b"]
        fs[k], fe[k] = int(j["force_start"]), int(j["force_end"])
        cap = max(cap, len(a) + len(b) + 2 * j["band"] + 64)
    outs = [np.zeros(cap, dtype=np.uint8) for _ in range(n)]
    for k in range(n):
        op[k] = outs[k].ctypes.data_as(u8p)
    res = (GamxResult * n)()
    L.sim_align_multi(C.c_int(n), ap, la, bp, lb, ba, ea, bb, eb, u64(jobs[0]["band"]), C.c_int64(jobs[0]["gap"]),
                      fs, fe, C.c_int(mode), C.c_int(lane_order), res, op, u64(cap))
    return [(res[k], outs[k]) if res[k].status >= 0 else None for k in range(n)]


def sim_align_pairs(jobs, mode=2, lane_order=0, first_group=0, rc=0):
    """Runs jobs (same band and gap, no N) as 16x2 PAIRS in ONE simulated warp: jobs 2g, 2g+1 on lane
    group first_group+g (bsw_warp16.h + PairFetch traceback).  Returns (rc, [(result, ops)])."""
    L = lib()
    n = len(jobs)
    u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
    keep = []
    ap, bp, op = (u8p * n)(), (u8p * n)(), (u8p * n)()
    la, lb, ba, ea, bb, eb = [(u64 * n)() for _ in range(6)]
    fs, fe = (C.c_int * n)(), (C.c_int * n)()
    cap = 0
    for k, j in enumerate(jobs):
        a = np.ascontiguousarray(j["a"], dtype=np.uint8); b = np.ascontiguousarray(j["b"], dtype=np.uint8)
        keep += [a, b]
        ap[k], bp[k] = a.ctypes.data_as(u8p), b.ctypes.data_as(u8p)
        la[k], lb[k], ba[k], ea[k], bb[k], eb[k] = len(a), len(b), j["begin_a"], j["end_a"], j["begin_b"], j["end_b"]
        fs[k], fe[k] = int(j["force_start"]), int(j["force_end"])
        cap = max(cap, len(a) + len(b) + 2 * j["band"] + 64)
    outs = [np.zeros(cap, dtype=np.uint8) for _ in range(n)]
    for k in range(n):
        op[k] = outs[k].ctypes.data_as(u8p)
    res = (GamxResult * n)()
    L.sim_align_pairs.restype = C.c_int
    rc = L.sim_align_pairs(C.c_int(n), ap, la, bp, lb, ba, ea, bb, eb, u64(jobs[0]["band"]), C.c_int64(jobs[0]["gap"]),
                           fs, fe, C.c_int(mode), C.c_int(lane_order), C.c_int(first_group), C.c_int(rc), res, op, u64(cap))
    return rc, [(res[k], outs[k]) for k in range(n)]


def band_geometry(band):
    """(stripe width, lanes per pair) the host picks for a band (bsw_host.h: geometry_for_band)."""
    L = lib()
    c, lg = C.c_int(0), C.c_int(0)
    L.sim_band_geometry(C.c_uint64(band), C.byref(c), C.byref(lg))
    return c.value, lg.value
