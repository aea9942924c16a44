#!/usr/bin/env python3
"""SASS loop census of one kernel: tools/sass_loops.py <lib.so> <kernel substring> [min VIADDMNMX per loop]
Prints every backward branch whose body holds DP cell updates, with its length and opcode histogram,
and a histogram of the whole function (the committed evidence for DESIGN.md's instruction counts)."""
import collections
import re
import subprocess
import sys


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
            continue
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if cur:
        yield cur, body


def opcode(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    min_cells = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    for name, ins in functions(lib):
        if pat not in name:
            continue
        addr = {a: i for i, (a, _) in enumerate(ins)}
        print(f"{name[:110]}: {len(ins)} instructions")
        print("  all:", dict(collections.Counter(opcode(t) for _, t in ins).most_common(16)))
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(0x[0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                nv = sum(1 for _, x in body if "VIADDMNMX" in x or "VIMNMX" in x)
                if nv >= min_cells:
                    print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, {nv} min/max ops:",
                          dict(collections.Counter(opcode(x) for _, x in body).most_common(14)))


if __name__ == "__main__":
    main()
