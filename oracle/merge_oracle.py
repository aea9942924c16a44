"""TEST INFRASTRUCTURE ONLY - sequential CPU restatement of gam-merge's alignment stage.

Follows, statement by statement and on plain Python structures,
    PctgBuilder::alignMergeBlock     /root/reference/lib/src/pctg/PctgBuilder.cc:726-844
    PctgBuilder::findBestAlignment   /root/reference/lib/src/pctg/PctgBuilder.cc:1361-1614
    PctgBuilder::alignBlocks         /root/reference/lib/src/pctg/PctgBuilder.cc:1617-1708
    PctgBuilder::is_good             /root/reference/lib/src/pctg/PctgBuilder.cc:1711-1730
    BestCtgAlignment::main_homology  /root/reference/lib/src/pctg/BestCtgAlignment.cc:109-126
with Frame / Block / MergeBlock reduced to the fields those functions read
(lib/include/assembly/Frame.hpp:53-60, Block.hpp:68-71, pctg/MergeDescriptor.hpp:40-69).

Parity status: PINNED.  Every alignment and every findHits call goes through a checker that is pinned to the
reference (oracle.restatement() / oracle.reference()), and the control flow above them is pinned to the reference's
own caller code: PctgBuilder.cc cannot be compiled whole (Boost.Graph, sparsehash, BamTools), but the bodies of
the four functions above are copied out of it at build time (oracle/pctg_extract.py) and compiled unmodified
against small stand-ins for Block / Frame / the graph (oracle/pctg_shim.cc -> oracle/_ref/libgamref.so:
gamref_align_merge_block).  tests/test_merge_oracle.py compares this file with that build on adversarial merge
blocks and with the golden vectors generated from it (tests/golden/merge_golden.json).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MIN_HOMOLOGY = 95.0  # PctgBuilder.hpp:63-65
U64 = 1 << 64


@dataclass
class Frame:  # Frame.hpp:53-60 (0-based inclusive coordinates)
    strand: str
    begin: int
    end: int

    def length(self) -> int:  # Frame.cc:124-127
        return 0 if self.end < self.begin else self.end - self.begin + 1


@dataclass
class Block:  # Block.hpp:68-71
    num_reads: int
    mf: Frame
    sf: Frame


@dataclass
class MergeBlock:  # MergeDescriptor.hpp:40-69 (fields the alignment stage reads / writes)
    m_id: int
    s_id: int
    blocks: list
    m_ltail: bool = True
    m_rtail: bool = True
    s_ltail: bool = True
    s_rtail: bool = True
    # outputs
    align_ok: bool = False
    align_rev: bool = False
    m_start: int = 0
    m_end: int = 0
    s_start: int = 0
    s_end: int = 0
    coords_set: bool = False


@dataclass
class Aln:
    """The fields of MyAlignment gam-merge reads (my_alignment.hpp:65-126 + my_alignment.cc:167-262)."""
    homology: float = 0.0
    length: int = 0
    begin_a: int = 0
    begin_b: int = 0
    has_match: bool = False
    first_match: tuple = (0, 0)
    last_match: tuple = (0, 0)


def _revcomp(s):
    comp = np.array([1, 0, 3, 2, 4], dtype=np.uint8)
    return comp[s[::-1]].copy()


class Stats:
    def __init__(self):
        self.alignments = 0
        self.cells = 0
        self.hits_calls = 0
        self.exceptions = 0


class MergeOracle:
    """aligner: oracle.restatement() or oracle.reference() (both expose align / find_hits)."""

    def __init__(self, checker, hits_checker=None, band=150):
        self.chk = checker
        self.hits_chk = hits_checker or checker
        self.band = band  # BandedSmithWaterman() default, banded_smith_waterman.hpp:38
        self.stats = Stats()

    # -- BandedSmithWaterman::find_alignment + the reductions; raises IndexError for std::out_of_range
    def find_alignment(self, a, begin_a, end_a, b, begin_b, end_b, force_start=False, force_end=False) -> Aln:
        begin_a, end_a, begin_b, end_b = int(begin_a) % U64, int(end_a) % U64, int(begin_b) % U64, int(end_b) % U64
        r, _ = self.chk.align(a, begin_a, end_a, b, begin_b, end_b, self.band, -8, force_start, force_end,
                              want_ops=False)
        self.stats.alignments += 1
        # DP cells of this call: x_size * (2*band+1), banded_smith_waterman.cc:90-97
        la, lb = len(a), len(b)
        if end_b >= begin_b:
            eb = end_b if end_b < lb else (lb - 1) % U64
            x = min((eb - begin_b + 1) % U64, (la + self.band - begin_a) % U64, 500000)
            self.stats.cells += x * (2 * self.band + 1)
        if r.status == 2:
            raise IndexError("Contig::at")
        if r.status != 0 or r.n_ops == 0:
            return Aln()  # default MyAlignment(): homology 0, length 0, begin (0,0)
        return Aln(float(r.homology), int(r.n_ops), int(r.begin_a), int(r.begin_b), bool(r.has_last_match),
                   (int(r.first_match_a), int(r.first_match_b)), (int(r.last_match_a), int(r.last_match_b)))

    def find_hits(self, a, a_start, a_end, b, b_start, b_end):
        self.stats.hits_calls += 1
        out = self.hits_chk.find_hits(a, int(a_start) % U64, int(a_end) % U64, b, int(b_start) % U64, int(b_end) % U64)
        return [int(h) for h in (out[0] if isinstance(out, tuple) else out)]

    @staticmethod
    def is_good_list(aligns, min_len):  # PctgBuilder.cc:1711-1723
        total = 0
        for al in aligns:
            if al.homology < MIN_HOMOLOGY:
                return False
            total += al.length
        return total >= min_len

    @staticmethod
    def is_good(al, min_len):  # PctgBuilder.cc:1726-1729
        return al.homology >= MIN_HOMOLOGY and al.length >= min_len

    def align_blocks(self, master, master_start, slave, slave_start, blocks):  # PctgBuilder.cc:1617-1708
        aligns = []
        first, last = blocks[0], blocks[-1]
        order = blocks if first.mf.begin <= last.mf.begin else blocks[::-1]
        m_start, s_start = master_start, slave_start
        last_match = (0, 0)
        prev = None
        for idx, b in enumerate(order):
            mf, sf = b.mf, b.sf
            mlen, slen = mf.length(), sf.length()
            if idx > 0:
                pmf, psf = prev.mf, prev.sf
                mgap = (mf.begin - pmf.end - 1) if pmf.begin <= mf.begin else (pmf.begin - mf.end - 1)
                sgap = (sf.begin - psf.end - 1) if psf.begin <= sf.begin else (psf.begin - sf.end - 1)
                m_start = max(last_match[0] + mgap, 0)
                s_start = max(last_match[1] + sgap, 0)
            al = self.find_alignment(master, m_start, m_start + mlen - 1, slave, s_start, s_start + slen - 1)
            aligns.append(al)
            last_match = al.last_match if al.length else (al.begin_a, al.begin_b)  # last_match_pos()
            prev = b
        return aligns

    def find_best_alignment(self, master, master_start, master_end, slave, slave_start, slave_end, blocks):
        """PctgBuilder.cc:1361-1614.  Returns dict(main=[Aln] or None for the bad alignment, rev, left, right,
        left_rev, right_rev, slave=<slave in the orientation it was left in>)."""
        con = dis = 0
        min_frame_len = 100
        for n, b in enumerate(blocks):
            ml = min(b.mf.length(), b.sf.length())
            if n == 0 or min_frame_len > ml:
                min_frame_len = ml
            if b.mf.strand != b.sf.strand:
                dis += b.num_reads
            else:
                con += b.num_reads
        con_prob = (con / (con + dis)) if (con + dis) else float("nan")
        mt, st = int(0.3 * len(master)), int(0.3 * len(slave))
        align_threshold = int(0.7 * min_frame_len)
        threshold = min(200, mt, st)
        good, rev = False, False
        aligns = []
        size = len(slave)
        if con_prob >= 0.5:
            aligns = self.align_blocks(master, master_start, slave, slave_start, blocks)
            if self.is_good_list(aligns, align_threshold):
                good, rev = True, False
            else:
                slave = _revcomp(slave)
                slave_start, slave_end = size - slave_end - 1, size - slave_start - 1
                aligns = self.align_blocks(master, master_start, slave, slave_start, blocks)
                if self.is_good_list(aligns, align_threshold):
                    good, rev = True, True
        if con_prob < 0.5:
            slave = _revcomp(slave)
            slave_start, slave_end = size - slave_end - 1, size - slave_start - 1
            aligns = self.align_blocks(master, master_start, slave, slave_start, blocks)
            if self.is_good_list(aligns, align_threshold):
                good, rev = True, True
            else:
                slave = _revcomp(slave)
                slave_start, slave_end = size - slave_end - 1, size - slave_start - 1
                aligns = self.align_blocks(master, master_start, slave, slave_start, blocks)
                if self.is_good_list(aligns, align_threshold):
                    good, rev = True, False
        nb = len(blocks)
        if (not good) or len(aligns) != nb or nb == 0:
            return dict(main=None, rev=rev, left=None, right=None, left_rev=False, right_rev=False, slave=slave)
        a_start = aligns[0].first_match if aligns[0].length else (aligns[0].begin_a, aligns[0].begin_b)
        a_end = aligns[-1].last_match if aligns[-1].length else (aligns[-1].begin_a, aligns[-1].begin_b)
        i1, i2 = a_start[0], (len(master) - a_end[0] - 1) % U64
        j1, j2 = a_start[1], (len(slave) - a_end[1] - 1) % U64
        out = dict(main=aligns, rev=rev, left=None, right=None, left_rev=False, right_rev=False, slave=slave)
        if min(i1, j1) < threshold and min(i2, j2) < threshold:
            return out
        if min(i1, j1) >= threshold:  # left tails
            if i1 < j1:
                hits = self.find_hits(slave, 0, a_start[1] - 1, master, 0, a_start[0] - 1)
                ba = hits[-1] if hits else a_start[1] - a_start[0]
                out["left"] = self.find_alignment(slave, ba, a_start[1] - 1, master, 0, a_start[0] - 1, False, True)
                out["left_rev"] = True
            else:
                hits = self.find_hits(master, 0, a_start[0] - 1, slave, 0, a_start[1] - 1)
                ba = hits[-1] if hits else a_start[0] - a_start[1]
                out["left"] = self.find_alignment(master, ba, a_start[0] - 1, slave, 0, a_start[1] - 1, False, True)
                out["left_rev"] = False
        if min(i2, j2) >= threshold:  # right tails
            if i2 < j2:
                if len(slave) <= a_end[1] + 1:
                    raise ValueError("chop_borders: std::domain_error")  # contig.code.hpp:235-237
                tail = slave[a_end[1] + 1:]
                hits = self.find_hits(tail, 0, len(tail) - 1, master, a_end[0] + 1, len(master) - 1)
                ba = hits[0] if hits else 0
                out["right"] = self.find_alignment(tail, ba, len(tail) - 1, master, a_end[0] + 1, len(master) - 1, True, False)
                out["right_rev"] = True
            else:
                if len(master) <= a_end[0] + 1:
                    raise ValueError("chop_borders: std::domain_error")
                tail = master[a_end[0] + 1:]
                hits = self.find_hits(tail, 0, len(tail) - 1, slave, a_end[1] + 1, len(slave) - 1)
                ba = hits[0] if hits else 0
                out["right"] = self.find_alignment(tail, ba, len(tail) - 1, slave, a_end[1] + 1, len(slave) - 1, True, False)
                out["right_rev"] = False
        return out

    def align_merge_block(self, mb: MergeBlock, master, slave):
        """PctgBuilder.cc:726-844: fills mb.align_ok / align_rev / m_start.. ; exceptions propagate like in the
        reference (caught per graph at ThreadedBuildPctg.cc:322-329)."""
        blocks = mb.blocks
        f, l = blocks[0], blocks[-1]
        m_start, m_end = min(f.mf.begin, l.mf.begin), max(f.mf.end, l.mf.end)
        s_start, s_end = min(f.sf.begin, l.sf.begin), max(f.sf.end, l.sf.end)
        best = self.find_best_alignment(master, m_start, m_end, slave, s_start, s_end, blocks)
        mb.align_ok = True
        main = best["main"]
        main_hom = min(al.homology for al in main) if main else 0.0  # bad_align(0): homology 0
        if main_hom < MIN_HOMOLOGY:
            mb.align_ok = False
            return mb
        first, last = main[0], main[-1]
        a_start = list(first.first_match if first.length else (first.begin_a, first.begin_b))
        a_end = list(last.last_match if last.length else (last.begin_a, last.begin_b))
        msz, ssz = len(master), len(slave)
        i1, i2 = a_start[0], (msz - a_end[0] - 1) % U64
        j1, j2 = a_start[1], (ssz - a_end[1] - 1) % U64
        left = best["left"] or Aln(homology=100.0)   # MyAlignment(100): homology 100, empty
        right = best["right"] or Aln(homology=100.0)
        mt, st = int(0.3 * msz), int(0.3 * ssz)
        left_min, right_min = int(0.7 * min(i1, j1)), int(0.7 * min(i2, j2))
        threshold = min(100, mt, st)
        s_lt = mb.s_rtail if best["rev"] else mb.s_ltail
        s_rt = mb.s_ltail if best["rev"] else mb.s_rtail
        if mb.m_ltail and s_lt and min(i1, j1) >= threshold:
            if self.is_good(left, left_min):
                a_start = list(left.first_match if left.length else (left.begin_a, left.begin_b))
                if best["left_rev"]:
                    a_start.reverse()
            else:
                mb.align_ok = False
        if mb.m_rtail and s_rt and min(i2, j2) >= threshold:
            if self.is_good(right, right_min):
                tmp = list(right.last_match if right.length else (right.begin_a, right.begin_b))
                if best["right_rev"]:
                    tmp.reverse()
                    a_end[0] = tmp[0]
                    a_end[1] += tmp[1] + 1
                else:
                    a_end[0] += tmp[0] + 1
                    a_end[1] = tmp[1]
            else:
                mb.align_ok = False
        if best["rev"]:
            t = a_start[1]
            a_start[1] = (ssz - a_end[1] - 1) % U64
            a_end[1] = (ssz - t - 1) % U64
        mb.align_rev = best["rev"]
        # MergeBlock fields are int32_t (MergeDescriptor.hpp:44-49)
        wrap = lambda v: ((int(v) + (1 << 31)) % (1 << 32)) - (1 << 31)  # noqa: E731
        mb.m_start, mb.m_end, mb.s_start, mb.s_end = wrap(a_start[0]), wrap(a_end[0]), wrap(a_start[1]), wrap(a_end[1])
        mb.coords_set = True
        return mb
