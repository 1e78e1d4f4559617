"""Per-CUDA-source-line executed instructions and stall samples of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):   python tools/ncu_lines.py rep.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
seen = 0
fname = ""
agg = []
hdr = None
for r in rows:
    if r and r[0] == "File Name": fname = r[1].split("/")[-1]
    if "Instructions Executed" in r:
        hdr = r; ci = r.index("Instructions Executed"); st = r.index("Warp Stall Sampling (All Samples)"); continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        try: agg.append((int(r[ci]), int(r[st]), fname, int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(a[0] for a in agg) or 1; tots = sum(a[1] for a in agg) or 1
print(f"total executed warp instructions {tot}, stall samples {tots}")
print("by executed instructions:")
for e, s, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"  {100*e/tot:5.1f}% inst {100*s/tots:5.1f}% stall  {f}:{ln:<5d} {src[:100]}")
