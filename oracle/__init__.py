"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracles.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package.  The product path (``gam_ngs_b200``) never does.

Two checkers live here:

* ``restatement()`` -> ``oracle/_build/libbsw_oracle.so``: the plain-C restatement
  (``oracle/bsw_oracle.c``) of ``BandedSmithWaterman::find_alignment``
  (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323).
* ``reference()`` -> ``oracle/_ref/libgamref.so``: the UNMODIFIED reference aligner
  compiled from /root/reference by ``oracle/Makefile`` (``make ref``).  The built
  ``.so`` travels to the GPU box; ``/root/reference`` itself does not.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

STATUS_OK, STATUS_EMPTY, STATUS_OUT_OF_RANGE, STATUS_UNDEFINED = 0, 1, 2, 3
GAP_DEFAULT = -8  # my_alignment.hpp:46
INT64_MIN = -(2**63)


class RefResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("has_first_match", C.c_int32),
        ("has_last_match", C.c_int32),
        ("has_last_pos", C.c_int32),
        ("score", C.c_int64),
        ("begin_a", C.c_uint64),
        ("begin_b", C.c_uint64),
        ("a_size", C.c_uint64),
        ("b_size", C.c_uint64),
        ("n_ops", C.c_uint64),
        ("homology", C.c_double),
        ("first_match_a", C.c_uint64),
        ("first_match_b", C.c_uint64),
        ("last_match_a", C.c_uint64),
        ("last_match_b", C.c_uint64),
        ("last_pos_a", C.c_uint64),
        ("last_pos_b", C.c_uint64),
        ("gaps_a", C.c_uint64),
        ("gaps_b", C.c_uint64),
        ("has_gaps", C.c_int32),
        ("pad_", C.c_int32),
    ]


class MergeBlockRec(C.Structure):
    """oracle/pctg_shim.cc: gamref_block (= include/gamx.h gamx_block)"""
    _fields_ = [("num_reads", C.c_int32), ("m_strand", C.c_uint8), ("s_strand", C.c_uint8), ("reserved_", C.c_uint8 * 2),
                ("m_begin", C.c_int32), ("m_end", C.c_int32), ("s_begin", C.c_int32), ("s_end", C.c_int32)]


class MergeRefResult(C.Structure):
    """oracle/pctg_shim.cc: gamref_merge_result"""
    _fields_ = [("status", C.c_int32), ("align_ok", C.c_int32), ("align_rev", C.c_int32), ("coords_set", C.c_int32),
                ("m_start", C.c_int32), ("m_end", C.c_int32), ("s_start", C.c_int32), ("s_end", C.c_int32)]


class OracleResult(C.Structure):
    _fields_ = RefResult._fields_ + [
        ("n_match", C.c_uint64),
        ("x_size", C.c_uint64),
        ("end_i", C.c_int64),
        ("end_j", C.c_int64),
    ]


COMPARE_FIELDS = [
    "status", "score", "begin_a", "begin_b", "a_size", "b_size", "n_ops", "homology",
    "has_first_match", "first_match_a", "first_match_b",
    "has_last_match", "last_match_a", "last_match_b",
    "has_last_pos", "last_pos_a", "last_pos_b",
    "has_gaps", "gaps_a", "gaps_b",
]


def result_dict(r, ops=None):
    d = {k: getattr(r, k) for k in COMPARE_FIELDS}
    if d["status"] != STATUS_OK:
        d = {"status": d["status"]}
    elif ops is not None:
        d["ops"] = bytes(ops[: r.n_ops])
    return d


def build(which=("oracle", "ref"), quiet=True):
    """Compile the checkers.  Building the checker is not using it."""
    for target in which:
        if target == "ref" and not os.path.isdir(REF_ROOT):
            continue  # GPU box: only the prebuilt oracle/_ref/libgamref.so is used
        subprocess.run(["make", "-C", HERE, target], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


def _u8(x):
    x = np.ascontiguousarray(x, dtype=np.uint8)
    return x, x.ctypes.data_as(C.POINTER(C.c_uint8))


class Restatement:
    def __init__(self):
        path = os.path.join(HERE, "_build", "libbsw_oracle.so")
        if not os.path.exists(path):
            build(("oracle",))
        self.lib = C.CDLL(path)
        u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
        self.lib.bswo_align.argtypes = [u8p, u64, u64, u64, u8p, u64, u64, u64, u64, C.c_int64,
                                        C.c_int, C.c_int, C.POINTER(OracleResult), u8p, u64]
        self.lib.bswo_align.restype = C.c_int
        self.lib.bswo_cells.argtypes = [u64] * 6
        self.lib.bswo_cells.restype = u64
        self.lib.bswo_find_hits.argtypes = [u8p, u64, u64, u64, u8p, u64, u64, u64, C.POINTER(C.c_uint32), u64,
                                            C.POINTER(u64)]
        self.lib.bswo_find_hits.restype = u64

    def align(self, a, begin_a, end_a, b, begin_b, end_b, band=150, gap=GAP_DEFAULT,
              force_start=False, force_end=False, want_ops=True):
        a, pa = _u8(a)
        b, pb = _u8(b)
        r = OracleResult()
        cap = int(len(a) + len(b) + 2 * band + 64) if want_ops else 0
        ops = np.zeros(max(cap, 1), dtype=np.uint8)
        self.lib.bswo_align(pa, len(a), begin_a, end_a, pb, len(b), begin_b, end_b, band, gap,
                            int(force_start), int(force_end), C.byref(r),
                            ops.ctypes.data_as(C.POINTER(C.c_uint8)) if want_ops else None, cap)
        return r, (ops if want_ops else None)

    def find_hits(self, a, a_start, a_end, b, b_start, b_end, cap=1 << 16):
        """ABlast::findHits restated: returns (hits ascending, max_count)."""
        a, pa = _u8(a)
        b, pb = _u8(b)
        hits = np.zeros(cap, dtype=np.uint32)
        mc = C.c_uint64(0)
        n = self.lib.bswo_find_hits(pa, len(a), a_start, a_end, pb, len(b), b_start, b_end,
                                    hits.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(mc))
        return hits[: min(n, cap)].copy(), int(mc.value)

    def cells(self, la, begin_a, lb, begin_b, end_b, band):
        return self.lib.bswo_cells(la, begin_a, lb, begin_b, end_b, band)


class Reference:
    def __init__(self):
        path = os.path.join(HERE, "_ref", "libgamref.so")
        if not os.path.exists(path):
            build(("ref",))
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        u8p, u64, vp = C.POINTER(C.c_uint8), C.c_uint64, C.c_void_p
        L = self.lib
        L.gamref_contig_new.argtypes = [u8p, u64]
        L.gamref_contig_new.restype = vp
        L.gamref_contig_free.argtypes = [vp]
        L.gamref_contig_size.argtypes = [vp]
        L.gamref_contig_size.restype = u64
        L.gamref_contig_codes.argtypes = [vp, u8p]
        L.gamref_contig_revcomp.argtypes = [vp]
        L.gamref_contig_revcomp.restype = vp
        L.gamref_contig_chop_begin.argtypes = [vp, u64]
        L.gamref_contig_chop_begin.restype = vp
        L.gamref_align.argtypes = [vp, u64, u64, vp, u64, u64, u64, C.c_int64, C.c_int, C.c_int,
                                   C.POINTER(RefResult), u8p, u64]
        L.gamref_align.restype = C.c_int
        L.gamref_align_codes.argtypes = [u8p, u64, u64, u64, u8p, u64, u64, u64, u64, C.c_int64,
                                         C.c_int, C.c_int, C.POINTER(RefResult), u8p, u64]
        L.gamref_align_codes.restype = C.c_int
        L.gamref_find_hits.argtypes = [vp, u64, u64, vp, u64, u64, C.POINTER(C.c_uint32), u64]
        L.gamref_find_hits.restype = u64
        L.gamref_bench.argtypes = [C.POINTER(u8p), C.POINTER(u64), C.POINTER(u8p), C.POINTER(u64),
                                   u64, u64, C.c_int, C.POINTER(u64), C.POINTER(C.c_int64)]
        L.gamref_bench.restype = C.c_double

    def align(self, a, begin_a, end_a, b, begin_b, end_b, band=150, gap=None,
              force_start=False, force_end=False, want_ops=True):
        """gap=None uses the reference's 1-arg ctor (band only)."""
        a, pa = _u8(a)
        b, pb = _u8(b)
        r = RefResult()
        cap = int(len(a) + len(b) + 2 * band + 64) if want_ops else 0
        ops = np.zeros(max(cap, 1), dtype=np.uint8)
        self.lib.gamref_align_codes(pa, len(a), begin_a, end_a, pb, len(b), begin_b, end_b, band,
                                    INT64_MIN if gap is None else gap,
                                    int(force_start), int(force_end), C.byref(r),
                                    ops.ctypes.data_as(C.POINTER(C.c_uint8)) if want_ops else None,
                                    cap)
        return r, (ops if want_ops else None)

    def revcomp(self, a):
        a, pa = _u8(a)
        h = self.lib.gamref_contig_new(pa, len(a))
        h2 = self.lib.gamref_contig_revcomp(h)
        out = np.zeros(len(a), dtype=np.uint8)
        self.lib.gamref_contig_codes(h2, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        self.lib.gamref_contig_free(h)
        self.lib.gamref_contig_free(h2)
        return out

    def find_hits(self, a, a_start, a_end, b, b_start, b_end, cap=1 << 16):
        a, pa = _u8(a)
        b, pb = _u8(b)
        ha = self.lib.gamref_contig_new(pa, len(a))
        hb = self.lib.gamref_contig_new(pb, len(b))
        hits = np.zeros(cap, dtype=np.uint32)
        n = self.lib.gamref_find_hits(ha, a_start, a_end, hb, b_start, b_end,
                                      hits.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
        self.lib.gamref_contig_free(ha)
        self.lib.gamref_contig_free(hb)
        return hits[: min(n, cap)].copy()

    def align_merge_block(self, master, slave, blocks, tails=(1, 1, 1, 1)):
        """PctgBuilder::alignMergeBlock (PctgBuilder.cc:726-844, the unmodified bodies compiled by pctg_shim.cc) on
        one merge block.  blocks: dicts with num_reads, m_strand, s_strand (0 '+', 1 '-'), m_begin, m_end, s_begin,
        s_end.  Returns a dict shaped like tests/merge_util.result_dict."""
        if not hasattr(self.lib, "gamref_align_merge_block"):
            raise RuntimeError("libgamref.so predates the PctgBuilder shim: rebuild with `make -C oracle ref`")
        m, pm = _u8(master)
        s_, ps = _u8(slave)
        blk = (MergeBlockRec * len(blocks))()
        for k, b in enumerate(blocks):
            blk[k].num_reads = b["num_reads"]; blk[k].m_strand = b["m_strand"]; blk[k].s_strand = b["s_strand"]
            blk[k].m_begin, blk[k].m_end, blk[k].s_begin, blk[k].s_end = b["m_begin"], b["m_end"], b["s_begin"], b["s_end"]
        r = MergeRefResult()
        f = self.lib.gamref_align_merge_block
        f.restype = C.c_int
        rc = f(pm, C.c_uint64(len(m)), ps, C.c_uint64(len(s_)), blk, C.c_uint32(len(blocks)), C.c_int(int(tails[0])),
               C.c_int(int(tails[1])), C.c_int(int(tails[2])), C.c_int(int(tails[3])), C.byref(r))
        if rc != 0:
            raise ValueError("merge block without blocks")
        d = dict(status=int(r.status))
        if d["status"]:
            return d
        d.update(align_ok=int(r.align_ok), coords_set=int(r.coords_set))
        if d["coords_set"]:
            d.update(align_rev=int(r.align_rev), m_start=int(r.m_start), m_end=int(r.m_end), s_start=int(r.s_start), s_end=int(r.s_end))
        return d

    def bench(self, a_list, b_list, band, n_threads):
        """Times full-window alignments of the pairs on n_threads host threads.
        Returns (seconds, cells, score_sum)."""
        n = len(a_list)
        u8p = C.POINTER(C.c_uint8)
        keep = []
        ap, bp = (u8p * n)(), (u8p * n)()
        al, bl = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        for i in range(n):
            x, px = _u8(a_list[i])
            y, py = _u8(b_list[i])
            keep += [x, y]
            ap[i], bp[i], al[i], bl[i] = px, py, len(x), len(y)
        cells, ssum = C.c_uint64(0), C.c_int64(0)
        sec = self.lib.gamref_bench(ap, al, bp, bl, n, band, n_threads, C.byref(cells),
                                    C.byref(ssum))
        return sec, cells.value, ssum.value


_rest = None
_ref = None


def restatement() -> Restatement:
    global _rest
    if _rest is None:
        _rest = Restatement()
    return _rest


def reference() -> Reference:
    global _ref
    if _ref is None:
        _ref = Reference()
    return _ref


def reference_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libgamref.so")) or os.path.isdir(REF_ROOT)
