"""Helpers shared by the parity tests."""
import json
import os

import numpy as np

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
LETTERS = "ATCGN"
_DEC = {c: i for i, c in enumerate(LETTERS)}


def dec(s):
    return np.array([_DEC[c] for c in s], dtype=np.uint8)


def load_golden():
    with open(os.path.join(HERE, "golden", "bsw_golden.json")) as f:
        g = json.load(f)
    cases = []
    for c in g["cases"]:
        job = dict(c["job"])
        job["a"], job["b"] = dec(job["a"]), dec(job["b"])
        exp = dict(c["expect"])
        if "ops" in exp:
            exp["ops"] = bytes(int(ch) for ch in exp["ops"])
        cases.append((job, exp))
    return cases


def load_merge_golden():
    """tests/golden/merge_golden.json (made by tests/golden/make_merge_golden.py from the compiled reference
    caller): [(masters, slaves, [merge block dicts with m, s, blocks, tails, expect])] per assembly."""
    with open(os.path.join(HERE, "golden", "merge_golden.json")) as f:
        g = json.load(f)
    code = {c: i for i, c in enumerate("ATCGN")}
    out = []
    for k, asm in enumerate(g["assemblies"]):
        M = [np.array([code[c] for c in m], dtype=np.uint8) for m in asm["masters"]]
        S = [np.array([code[c] for c in s], dtype=np.uint8) for s in asm["slaves"]]
        mbs = [dict(m=c["m"], s=c["s"], blocks=c["blocks"], tails=tuple(c["tails"]), expect=c["expect"])
               for c in g["cases"] if c["assembly"] == k]
        out.append((M, S, mbs))
    return out


def x_size_of(job):
    """banded_smith_waterman.cc:90-95 with the reference's unsigned wrap-around; None when
    the reference returns before sizing the matrix."""
    la, lb = len(job["a"]), len(job["b"])
    eb = job["end_b"]
    if eb < job["begin_b"]:
        return None
    if eb >= lb:
        eb = (lb - 1) % 2**64
    return min((eb - job["begin_b"] + 1) % 2**64, (la + job["band"] - job["begin_a"]) % 2**64, 500000)


def oracle_expect(job, impl=None):
    """Run a checker (default: the C restatement) and normalise like the golden file."""
    impl = impl or oracle.restatement()
    r, ops = impl.align(job["a"], job["begin_a"], job["end_a"], job["b"], job["begin_b"],
                        job["end_b"], job["band"], job["gap"], job["force_start"],
                        job["force_end"])
    d = oracle.result_dict(r, ops)
    if d["status"] == 0 and d["n_ops"] == 0:
        d = {"status": 1}
    return d
