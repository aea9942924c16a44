"""Per-source-line summary of an ncu report captured with --import-source on (-lineinfo build):
python tools/ncu_lines.py rep.ncu-rep [top]   -> share of warp-stall samples and executed instructions per line"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur = None; hdr = None; agg = {}
    for r in csv.reader(raw.splitlines()):
        if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
        if r and r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); continue
        if hdr is None or len(r) < 8 or r[0] == "": continue
        try:
            ln = int(r[0]); s = int(r[iS]); e = int(r[iE])
        except ValueError:
            continue
        a = agg.setdefault((cur, ln), [0, 0, r[1].strip()[:100]]); a[0] += s; a[1] += e
    ts = sum(v[0] for v in agg.values()) or 1; te = sum(v[1] for v in agg.values()) or 1
    print(f"samples {ts} instructions {te}")
    byfile = {}
    for (f, l), (s, e, _) in agg.items():
        a = byfile.setdefault(f, [0, 0]); a[0] += s; a[1] += e
    for f, (s, e) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:22s} {s / ts * 100:5.1f}% samples {e / te * 100:5.1f}% instr")
    for (f, l), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f}:{l:<4d} {s / ts * 100:5.1f}% smp {e / te * 100:5.1f}% ins | {src}")

if __name__ == "__main__":
    main()
