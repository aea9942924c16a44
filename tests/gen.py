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


def bulk_pairs(rng, n, length, div=0.02, indel_share=0.5, len_lo=None, len_hi=None, chunk=32768):
    """Vectorised generator for large batches: n pairs (a_k, b_k), b_k = mutate(a_k).
    Lengths: fixed `length`, or uniform in [len_lo, len_hi].  Returns
    (a_codes, a_lens, b_codes, b_lens): concatenated uint8 codes and per-pair lengths."""
    a_parts, b_parts, a_lens_all, b_lens_all = [], [], [], []
    if len_lo is not None:
        per_chunk = max(1, int(chunk * 1000 // max(1, (len_lo + len_hi) // 2)))
    else:
        per_chunk = max(1, int(chunk * 1000 // max(1, length)))
    done = 0
    while done < n:
        m = min(per_chunk, n - done)
        if len_lo is None:
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
