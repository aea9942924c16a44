#!/usr/bin/env python3
"""Regions of a kernel by executed instructions, from the per-SASS table of tools/ncu_sass_dump.sh:
   python tools/ncu_sass_regions.py gpurun_out/<name>_sass.csv.gz [top]
Consecutive SASS instructions with the same execution count form a region (a basic-block run); for each of
the top regions: share of executed warp instructions, share of stall samples, ALU/FMA/LSU instruction counts and
the dominant stall reasons.  Offsets are relative to the kernel's first instruction (cuobjdump addresses)."""
import collections, csv, gzip, re, sys

ALU = {"LOP3", "PRMT", "VIADDMNMX", "VIMNMX", "VIMNMX3", "SEL", "ISETP", "SHF", "VIADD", "IADD3", "LEA", "IABS", "BREV", "FLO", "POPC", "IDP", "MOV", "PLOP3", "IADD", "VABSDIFF", "BMSK", "SGXT", "ISCADD", "P2R", "R2P"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "SHFL", "LD", "ST", "LDL", "STL", "ATOMG", "RED", "LDSM", "LDC", "MATCH", "VOTE"}

def opcode(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text.strip()).split()[0].split(".")[0]

def main():
    path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(gzip.open(path, "rt")))
    hdr = rows[1]
    iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ins = []
    base = None
    for r in rows[2:]:
        if len(r) <= iE: continue
        a = int(r[iA], 16)
        if base is None: base = a
        st = {h: int(r[i] or 0) for i, h in stall_cols}
        ins.append((a - base, r[iS].strip(), int(r[iN] or 0), int(r[iE] or 0), st))
    tot_e = sum(x[3] for x in ins) or 1; tot_s = sum(x[2] for x in ins) or 1
    regions = []; cur = None
    for off, src, smp, ex, st in ins:
        if cur is None or ex != cur["ex"]:
            cur = {"lo": off, "ex": ex, "n": 0, "smp": 0, "ops": collections.Counter(), "st": collections.Counter()}
            regions.append(cur)
        cur["hi"] = off; cur["n"] += 1; cur["smp"] += smp; cur["ops"][opcode(src)] += 1
        for k, v in st.items(): cur["st"][k] += v
    pipes = collections.Counter(); allops = collections.Counter()
    for off, src, smp, ex, st in ins:
        o = opcode(src); allops[o] += ex
        pipes["alu" if o in ALU else "fma" if o in FMA else "lsu" if o in LSU else "other"] += ex
    print(f"executed warp instructions {tot_e}, samples {tot_s}; by pipe:", {k: f"{v / tot_e * 100:.1f}%" for k, v in pipes.items()})
    print("  top opcodes:", {k: f"{v / tot_e * 100:.1f}%" for k, v in allops.most_common(18)})
    tst = collections.Counter()
    for r in regions: tst.update(r["st"])
    print("  stalls:", {k[6:]: f"{v / tot_s * 100:.1f}%" for k, v in tst.most_common(10)})
    for r in sorted(regions, key=lambda r: -(r["ex"] * r["n"]))[:top]:
        e = r["ex"] * r["n"]
        alu = sum(v for k, v in r["ops"].items() if k in ALU); fma = sum(v for k, v in r["ops"].items() if k in FMA)
        lsu = sum(v for k, v in r["ops"].items() if k in LSU)
        print(f"{r['lo']:#07x}..{r['hi']:#07x} n={r['n']:4d} x{r['ex']:>9d} ins {e / tot_e * 100:5.1f}% smp {r['smp'] / tot_s * 100:5.1f}% "
              f"alu {alu} fma {fma} lsu {lsu} | {dict(r['ops'].most_common(6))} | {{{', '.join(f'{k[6:]} {v * 100 // max(1, r['smp'])}%' for k, v in r['st'].most_common(4))}}}")

if __name__ == "__main__":
    main()
