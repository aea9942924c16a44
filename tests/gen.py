"""Seeded synthetic inputs shared by the tests, golden-vector script and bench.py.

Base codes follow the reference's BaseType (lib/include/assembly/nucleotide.hpp:35-43):
A=0, T=1, C=2, G=3, N=4.
"""
from __future__ import annotations

import numpy as np

A, T, C, G, N = 0, 1, 2, 3, 4


def random_seq(rng, n, p_n=0.0):
    s = rng.integers(0, 4, size=n, dtype=np.uint8)
    if p_n > 0 and n:
        s[rng.random(n) < p_n] = N
    return s


def mutate(rng, s, div=0.02, indel_share=0.5, p_n=0.0):
    """Return a copy of s with a fraction `div` of edited positions; edits split
    sub : ins : del = (1-indel_share) : indel_share/2 : indel_share/2."""
    n = len(s)
    if n == 0:
        return s.copy()
    u = rng.random(n)
    p_sub = div * (1.0 - indel_share)
    p_ins = div * indel_share / 2
    p_del = div * indel_share / 2
    is_sub = u < p_sub
    is_ins = (u >= p_sub) & (u < p_sub + p_ins)
    is_del = (u >= p_sub + p_ins) & (u < p_sub + p_ins + p_del)
    out = s.copy()
    # substitution: pick a different ACGT base
    shift = rng.integers(1, 4, size=n, dtype=np.uint8)
    sub_val = np.where(s < 4, (s + shift) % 4, rng.integers(0, 4, size=n, dtype=np.uint8))
    out = np.where(is_sub, sub_val, out).astype(np.uint8)
    keep = ~is_del
    reps = np.where(is_ins, 2, 1) * keep
    res = np.repeat(out, reps)
    # inserted copies get a fresh random base: positions that are the 2nd copy
    idx = np.cumsum(reps) - 1  # index of last copy of each source position
    ins_pos = idx[is_ins & keep]
    res[ins_pos] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
    if p_n > 0 and len(res):
        res[rng.random(len(res)) < p_n] = N
    return res


def revcomp(s):
    """lib/include/assembly/contig.code.hpp:187-229: A<->T, C<->G, N stays."""
    comp = np.array([1, 0, 3, 2, 4], dtype=np.uint8)
    return comp[s[::-1]].copy()


def make_pair(rng, length, div=0.02, indel_share=0.5, p_n=0.0, offset=0):
    a = random_seq(rng, length, p_n)
    b = mutate(rng, a, div, indel_share, p_n)
    if offset:
        b = b[offset:]
    return a, b


def fuzz_case(rng, max_len=80, max_band=40):
    """A small random job exercising every clamp of SURVEY Appendix A."""
    la = int(rng.integers(0, max_len + 1))
    related = rng.random() < 0.7
    p_n = float(rng.choice([0.0, 0.0, 0.05, 0.3]))
    a = random_seq(rng, la, p_n)
    if related and la > 0:
        b = mutate(rng, a, div=float(rng.choice([0.0, 0.02, 0.1, 0.3])), p_n=p_n)
        cut = int(rng.integers(0, min(len(b), 12) + 1))
        if rng.random() < 0.5:
            b = b[cut:]
        else:
            b = np.concatenate([random_seq(rng, cut), b])
    else:
        b = random_seq(rng, int(rng.integers(0, max_len + 1)), p_n)
    lb = len(b)
    band = int(rng.integers(0, max_band + 1))
    mode = rng.integers(0, 4)
    if mode == 0:  # the common call shape: whole windows
        ba, ea, bb, eb = 0, max(la - 1, 0), 0, max(lb - 1, 0)
    elif mode == 1:  # in-range random windows
        ba = int(rng.integers(0, max(la, 1)))
        ea = int(rng.integers(ba, max(la, ba + 1)))
        bb = int(rng.integers(0, max(lb, 1)))
        eb = int(rng.integers(bb, max(lb, bb + 1)))
    else:  # anything goes, including out-of-range and inverted windows
        ba = int(rng.integers(0, la + band + 6))
        ea = int(rng.integers(0, la + band + 12))
        bb = int(rng.integers(0, lb + 4))
        eb = int(rng.integers(0, lb + 8))
    fs = bool(rng.random() < 0.25)
    fe = bool(rng.random() < 0.25)
    gap = int(rng.choice([-8, -8, -8, -8, -5, -12, -3, -1, 0, 2, -29, -30]))
    return dict(a=a, b=b, begin_a=ba, end_a=ea, begin_b=bb, end_b=eb, band=band, gap=gap,
                force_start=fs, force_end=fe)


def bulk_pairs(rng, n, length, div=0.02, indel_share=0.5, len_lo=None, len_hi=None, chunk=32768, lengths=None):
    """Vectorised generator for large batches: n pairs (a_k, b_k), b_k = mutate(a_k).
    Lengths: fixed `length`, or uniform in [len_lo, len_hi].  Returns
    (a_codes, a_lens, b_codes, b_lens): concatenated uint8 codes and per-pair lengths."""
    a_parts, b_parts, a_lens_all, b_lens_all = [], [], [], []
    if lengths is not None:
        lengths = np.asarray(lengths, dtype=np.int64)
        n = len(lengths)
        per_chunk = max(1, int(chunk * 1000 // max(1, int(lengths.mean()))))
    elif len_lo is not None:
        per_chunk = max(1, int(chunk * 1000 // max(1, (len_lo + len_hi) // 2)))
    else:
        per_chunk = max(1, int(chunk * 1000 // max(1, length)))
    done = 0
    while done < n:
        m = min(per_chunk, n - done)
        if lengths is not None:
            la = lengths[done:done + m]
        elif len_lo is None:
            la = np.full(m, length, dtype=np.int64)
        else:
            la = rng.integers(len_lo, len_hi + 1, size=m, dtype=np.int64)
        total = int(la.sum())
        a = rng.integers(0, 4, size=total, dtype=np.uint8)
        starts = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(la, out=starts[1:])
        n_edit = rng.binomial(total, div) if div > 0 else 0
        pos = np.unique(rng.integers(0, total, size=n_edit)) if n_edit else np.zeros(0, dtype=np.int64)
        kind = rng.random(len(pos))
        p_sub = 1.0 - indel_share
        sub_pos = pos[kind < p_sub]
        ins_pos = pos[(kind >= p_sub) & (kind < p_sub + indel_share / 2)]
        del_pos = pos[kind >= p_sub + indel_share / 2]
        src = a.copy()
        src[sub_pos] = (src[sub_pos] + rng.integers(1, 4, size=len(sub_pos), dtype=np.uint8)) % 4
        reps = np.ones(total, dtype=np.int8)
        reps[del_pos] = 0
        reps[ins_pos] = 2
        b = np.repeat(src, reps)
        # output index of the extra copy after each insertion position
        shift = np.searchsorted(ins_pos, ins_pos, side="left") - np.searchsorted(del_pos, ins_pos, side="left")
        b[ins_pos + shift + 1] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
        pair_of = lambda p: np.searchsorted(starts, p, side="right") - 1  # noqa: E731
        lb = la + np.bincount(pair_of(ins_pos), minlength=m) - np.bincount(pair_of(del_pos), minlength=m)
        a_parts.append(a); b_parts.append(b); a_lens_all.append(la); b_lens_all.append(lb)
        done += m
    return (np.concatenate(a_parts), np.concatenate(a_lens_all).astype(np.uint64),
            np.concatenate(b_parts), np.concatenate(b_lens_all).astype(np.uint64))


def mutate_with_map(rng, s, div=0.01, indel_share=0.5):
    """Like mutate(), but also returns pos_map with pos_map[p] = index in the result of source
    position p (for a deleted base: the index of the next surviving base); pos_map[len(s)] = len(result)."""
    n = len(s)
    u = rng.random(n)
    p_sub = div * (1.0 - indel_share)
    p_ins = div * indel_share / 2
    p_del = div * indel_share / 2
    is_sub = u < p_sub
    is_ins = (u >= p_sub) & (u < p_sub + p_ins)
    is_del = (u >= p_sub + p_ins) & (u < p_sub + p_ins + p_del)
    out = s.copy()
    shift = rng.integers(1, 4, size=n, dtype=np.uint8)
    out = np.where(is_sub & (s < 4), (s + shift) % 4, out).astype(np.uint8)
    reps = np.where(is_del, 0, np.where(is_ins, 2, 1)).astype(np.int64)
    res = np.repeat(out, reps)
    pos_map = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(reps, out=pos_map[1:])
    ins_pos = np.nonzero(is_ins)[0]
    res[pos_map[ins_pos] + 1] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
    return res, pos_map


def make_assembly(rng, genome_len=200_000, master_mean=60_000, slave_mean=40_000, div=0.01, rc_frac=0.5,
                  frame_lo=500, frame_hi=5000, min_overlap=500, p_n=0.0, trim_prob=0.0, wrong_strand_prob=0.0):
    """Synthetic master/slave assembly pair: a random genome cut at Poisson
    breakpoints into master contigs; a mutated copy cut independently into slave contigs, a fraction of
    them reverse-complemented; for every overlapping (master, slave) pair one merge block whose blocks
    are consecutive frames tiling the overlap.  Returns (masters, slaves, merge_blocks) with
    merge_blocks = [dict(m=idx, s=idx, blocks=[dict(num_reads, m_strand, s_strand, m_begin, m_end, s_begin, s_end)])]."""
    G = random_seq(rng, genome_len, p_n)
    Gs, pmap = mutate_with_map(rng, G, div)

    def cuts(total, mean):
        pts = [0]
        while pts[-1] < total:
            pts.append(pts[-1] + max(2000, int(rng.exponential(mean))))
        pts[-1] = total
        if len(pts) > 2 and pts[-1] - pts[-2] < 2000:
            pts.pop(-2)
        return pts

    mc = cuts(genome_len, master_mean)          # master contig k = G[mc[k]:mc[k+1]]
    sc_g = cuts(genome_len, slave_mean)         # slave cut points in genome coordinates
    masters = [G[mc[k]:mc[k + 1]].copy() for k in range(len(mc) - 1)]
    slaves, s_rc = [], []
    for k in range(len(sc_g) - 1):
        seq = Gs[pmap[sc_g[k]]:pmap[sc_g[k + 1]]].copy()
        rc = bool(rng.random() < rc_frac)
        slaves.append(revcomp(seq) if rc else seq)
        s_rc.append(rc)
    merge_blocks = []
    for mi in range(len(masters)):
        for si in range(len(slaves)):
            lo, hi = max(mc[mi], sc_g[si]), min(mc[mi + 1], sc_g[si + 1])   # overlap in genome coordinates
            if hi - lo < min_overlap:
                continue
            blocks = []
            p = lo
            while p < hi:
                q = min(hi, p + int(rng.integers(frame_lo, frame_hi + 1)))
                if hi - q < frame_lo // 2:
                    q = hi
                m_b, m_e = p - mc[mi], q - 1 - mc[mi]
                sb, se = int(pmap[p] - pmap[sc_g[si]]), int(pmap[q] - 1 - pmap[sc_g[si]])
                if se < sb:
                    p = q
                    continue
                if s_rc[si]:
                    L = len(slaves[si])
                    sb, se = L - 1 - se, L - 1 - sb
                blocks.append(dict(num_reads=max(1, (q - p) // 50), m_strand=0, s_strand=1 if s_rc[si] else 0,
                                   m_begin=int(m_b), m_end=int(m_e), s_begin=int(sb), s_end=int(se)))
                p = q
            # real blocks rarely tile the whole overlap: dropping leading / trailing frames leaves contig
            # tails beyond the aligned blocks, which triggers the findHits-seeded tail alignments
            if len(blocks) >= 3 and rng.random() < trim_prob:
                blocks = blocks[int(rng.integers(0, 2)):len(blocks) - int(rng.integers(0, 2))]
                if len(blocks) >= 3 and rng.random() < 0.5:
                    blocks = blocks[1:]
            # wrong strand evidence makes findBestAlignment try the other orientation first
            if rng.random() < wrong_strand_prob:
                for b in blocks:
                    b["s_strand"] ^= 1
            if blocks:
                merge_blocks.append(dict(m=mi, s=si, blocks=blocks))
    return masters, slaves, merge_blocks


def perturb_merge_blocks(rng, masters, slaves, merge_blocks, n_extra=12):
    """Adversarial variants of an assembly's merge blocks for the caller-side parity tests (the control flow of
    PctgBuilder.cc:726-844,1361-1730 beyond the happy path): random tail flags, blocks listed in reverse order,
    shifted / empty / out-of-contig frames, wrong strand evidence on some blocks only, unrelated contig pairs and
    pairs whose slave was damaged after the blocks were laid out.  Returns (slaves', merge_blocks')."""
    slaves = [s.copy() for s in slaves]
    out = []
    for mb in merge_blocks:
        blocks = [dict(b) for b in mb["blocks"]]
        kind = int(rng.integers(0, 10))
        if kind == 0 and len(blocks) > 1:
            blocks = blocks[::-1]                                   # processed in reverse order (.cc:1677-1706)
        elif kind == 1:
            for b in blocks:                                        # frames that no longer match: chained starts drift
                d = int(rng.integers(-300, 300))
                b["m_begin"] += d; b["m_end"] += d
        elif kind == 2 and len(blocks) > 2:
            blocks = blocks[len(blocks) // 2:len(blocks) // 2 + 1]  # one frame in the middle: two long tails
        elif kind == 3:
            b = blocks[int(rng.integers(0, len(blocks)))]
            b["m_end"] = b["m_begin"] - 1                           # empty master frame: end_a = m_at + 0 - 1
        elif kind == 4:
            b = blocks[-1]
            b["s_end"] += 5000; b["m_end"] += 5000                  # frame beyond both contigs
        elif kind == 5:
            for b in blocks[::2]:
                b["s_strand"] ^= 1                                  # mixed strand evidence
        elif kind == 6:
            s = slaves[mb["s"]]                                     # a damaged stretch inside the slave
            lo = int(rng.integers(0, max(1, len(s) - 400)))
            s[lo:lo + 400] = rng.integers(0, 4, min(400, len(s) - lo)).astype(np.uint8)
        elif kind == 7:
            s = slaves[mb["s"]]                                     # damaged slave tails (tail alignments fail)
            s[:300] = rng.integers(0, 4, min(300, len(s))).astype(np.uint8)
            s[-300:] = rng.integers(0, 4, min(300, len(s))).astype(np.uint8)
        elif kind == 8 and len(blocks) > 2:
            b = blocks[len(blocks) // 2]                            # one good frame, the slave damaged right beside it:
            blocks = [b]                                            # the chained part passes, a tail alignment fails
            s = slaves[mb["s"]]
            for lo in (b["s_begin"] - 450, b["s_end"] + 50):
                lo = max(0, min(len(s) - 1, lo))
                hi = min(len(s), lo + 400)
                s[lo:hi] = rng.integers(0, 4, hi - lo).astype(np.uint8)
        tails = tuple(int(x) for x in rng.integers(0, 2, 4)) if rng.random() < 0.5 else (1, 1, 1, 1)
        out.append(dict(m=mb["m"], s=mb["s"], blocks=blocks, tails=tails))
    for _ in range(n_extra):                                        # unrelated pairs
        m, s = int(rng.integers(0, len(masters))), int(rng.integers(0, len(slaves)))
        lm, ls = len(masters[m]), len(slaves[s])
        n = int(rng.integers(1, 4))
        blocks = []
        for k in range(n):
            mb_, sb_ = int(rng.integers(0, lm // 2)), int(rng.integers(0, ls // 2))
            ln = int(rng.integers(50, 1500))
            blocks.append(dict(num_reads=int(rng.integers(1, 40)), m_strand=0, s_strand=int(rng.integers(0, 2)),
                               m_begin=mb_, m_end=min(lm - 1, mb_ + ln), s_begin=sb_, s_end=min(ls - 1, sb_ + ln)))
        out.append(dict(m=m, s=s, blocks=blocks, tails=tuple(int(x) for x in rng.integers(0, 2, 4))))
    return slaves, out
