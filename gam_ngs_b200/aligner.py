"""Python mirror of the reference's aligner interface, for tests and scripting.

Same names, argument meaning and error behaviour as
    BandedSmithWaterman   lib/include/alignment/banded_smith_waterman.hpp:41-72
    MyAlignment           lib/include/alignment/my_alignment.hpp:65-126
    Contig                lib/include/assembly/contig.hpp:50-108
The C++ drop-in with the identical signatures lives in gam_ngs_b200/cpp/.  Everything is
computed by the CUDA kernels behind the C ABI (include/gamx.h); a single find_alignment call
is a batch of one (correct, not fast) - use BandedSmithWaterman.find_alignments or
Context.align_batch for throughput.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

from . import capi

GAP_A, GAP_B, MATCH, MISMATCH = 0, 1, 2, 3  # AlignmentAlphabet, my_alignment.hpp:57-62
_CHAR2CODE = np.full(256, 4, dtype=np.uint8)  # nucleotide.code.hpp:47-75: unknown -> N
for _i, _ch in enumerate("ATCG"):
    _CHAR2CODE[ord(_ch)] = _i
    _CHAR2CODE[ord(_ch.lower())] = _i

_default_ctx = None


def default_context() -> capi.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = capi.Context()
    return _default_ctx


class Contig:
    """Sequence container: base codes A=0,T=1,C=2,G=3,N=4 (nucleotide.hpp:35-43)."""

    def __init__(self, seq, name: str = ""):
        self.name = name
        if isinstance(seq, (str, bytes)):
            raw = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
            self.codes = _CHAR2CODE[raw]
        else:
            self.codes = np.minimum(np.ascontiguousarray(seq, dtype=np.uint8), 4)
        self._ids = {}  # context -> contig id in its store

    def size(self) -> int:
        return len(self.codes)

    __len__ = size

    def at(self, i: int) -> int:
        if not 0 <= i < len(self.codes):
            raise IndexError("Contig::at")  # std::out_of_range, contig.code.hpp:141-151
        return int(self.codes[i])

    def store_id(self, ctx: capi.Context) -> int:
        key = id(ctx)
        if key not in self._ids:
            self._ids[key] = ctx.add_contig(self.codes)
        return self._ids[key]


class MyAlignment:
    """Result value type; default-constructed = all zero, empty edit string (my_alignment.cc:39-47)."""

    def __init__(self, begin_a=0, begin_b=0, a_size=0, b_size=0, score=0, homology=0.0,
                 sequence: np.ndarray | None = None, extra: dict | None = None):
        self._begin_a, self._begin_b = int(begin_a), int(begin_b)
        self._a_size, self._b_size = int(a_size), int(b_size)
        self._score, self._homology = int(score), float(homology)
        self._sequence = np.zeros(0, dtype=np.uint8) if sequence is None else sequence
        self._extra = extra or {}

    def begin_a(self): return self._begin_a
    def begin_b(self): return self._begin_b
    def a_size(self): return self._a_size
    def b_size(self): return self._b_size
    def score(self): return self._score
    def homology(self): return self._homology
    def sequence(self): return self._sequence
    def length(self): return int(self._extra.get("n_ops", len(self._sequence)))


def _walk(al: MyAlignment):
    a, b = al.begin_a(), al.begin_b()
    for op in al.sequence():
        yield int(op), a, b
        if op == GAP_A:
            b += 1
        elif op == GAP_B:
            a += 1
        else:
            a += 1
            b += 1


def first_match_pos(al: MyAlignment):
    """my_alignment.cc:167-193 -> (found, (a, b)); served from the device reduction when present."""
    e = al._extra
    if e:
        return bool(e["has_match"]), (e["first_match_a"], e["first_match_b"])
    a, b = al.begin_a(), al.begin_b()
    for op, pa, pb in _walk(al):
        a, b = pa, pb
        if op == MATCH:
            return True, (pa, pb)
        a, b = pa + (op != GAP_A), pb + (op != GAP_B)
    return False, (a, b)


def last_match_pos(al: MyAlignment):
    """my_alignment.cc:228-262"""
    e = al._extra
    if e:
        return bool(e["has_match"]), (e["last_match_a"], e["last_match_b"])
    found, pos = False, (al.begin_a(), al.begin_b())
    for op, pa, pb in _walk(al):
        if op == MATCH:
            found, pos = True, (pa, pb)
    return found, pos


class BandedSmithWaterman:
    """Same constructors as banded_smith_waterman.cc:40-67: only gap_score and band_size take
    effect in the reference; match/mismatch/gap_ext are accepted and ignored."""

    def __init__(self, *args, ctx: capi.Context | None = None):
        self._gap, self._band = capi.DEFAULT_GAP, capi.DEFAULT_BAND
        if len(args) == 1:
            self._band = int(args[0])
        elif len(args) == 5:
            self._gap, self._band = int(args[2]), int(args[4])
        elif len(args) != 0:
            raise TypeError("BandedSmithWaterman(), (band_size) or (match, mismatch, gap, gap_ext, band_size)")
        self._ctx = ctx

    @property
    def ctx(self) -> capi.Context:
        return self._ctx or default_context()

    def _job(self, jobs, k, a, begin_a, end_a, b, begin_b, end_b, force_start, force_end, mode):
        j = jobs[k]
        j["a_id"], j["b_id"] = a.store_id(self.ctx), b.store_id(self.ctx)
        j["begin_a"], j["end_a"], j["begin_b"], j["end_b"] = begin_a, end_a, begin_b, end_b
        j["force_start"], j["force_end"] = int(force_start), int(force_end)
        j["band"], j["gap"], j["mode"] = self._band, self._gap, mode

    @staticmethod
    def _to_alignment(ctx, r, ops, mode) -> MyAlignment:
        st = int(r["status"])
        if st == capi.JOB_EMPTY:
            return MyAlignment()  # banded_smith_waterman.cc:90, :215
        if st == capi.JOB_OUT_OF_RANGE:
            raise IndexError("Contig::at: out of range")  # std::out_of_range via Contig::at
        if st == capi.JOB_UNDEFINED:
            raise ValueError("reference behaviour undefined for these arguments (x_size == 0)")
        seq = None
        if mode == capi.MODE_FULL:
            seq = ctx.unpack_ops(ops, int(r["ops_offset"]), int(r["n_ops"]))
        extra = {k: int(r[k]) for k in ("n_ops", "n_match", "has_match", "first_match_a", "first_match_b",
                                        "last_match_a", "last_match_b", "end_i", "end_j")}
        return MyAlignment(int(r["begin_a"]), int(r["begin_b"]), int(r["a_size"]), int(r["b_size"]),
                           int(r["score"]), float(r["homology"]), seq, extra)

    def find_alignment(self, a: Contig, begin_a: int, end_a: int, b: Contig, begin_b: int, end_b: int,
                       force_start: bool = False, force_end: bool = False,
                       mode: int = capi.MODE_FULL) -> MyAlignment:
        jobs = capi.make_jobs(1)
        self._job(jobs, 0, a, begin_a, end_a, b, begin_b, end_b, force_start, force_end, mode)
        res, ops = self.ctx.align_batch(jobs)
        return self._to_alignment(self.ctx, res[0], ops, mode)

    def find_alignments(self, calls: Iterable[Sequence], mode: int = capi.MODE_FULL):
        """Batched form: calls = [(a, begin_a, end_a, b, begin_b, end_b[, force_start, force_end]), ...].
        Returns a list with a MyAlignment or the exception instance the reference would raise."""
        calls = list(calls)
        jobs = capi.make_jobs(len(calls))
        for k, c in enumerate(calls):
            fs = c[6] if len(c) > 6 else False
            fe = c[7] if len(c) > 7 else False
            self._job(jobs, k, c[0], c[1], c[2], c[3], c[4], c[5], fs, fe, mode)
        res, ops = self.ctx.align_batch(jobs)
        out = []
        for r in res:
            try:
                out.append(self._to_alignment(self.ctx, r, ops, mode))
            except (IndexError, ValueError) as e:
                out.append(e)
        return out
