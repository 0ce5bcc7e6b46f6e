"""Per-SASS-chunk instruction / stall profile of the first kernel in an .ncu-rep (read here, no GPU):
python scripts/ncu_source.py gpurun_out/x.ncu-rep [chunk=60] [--lines: aggregate by source line instead]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"] + (["--print-source", "cuda,sass"] if "--lines" in sys.argv else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or "Instructions Executed" in r:
        break
    data.append(r)
iA, iS, iT = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Avg. Threads Executed")
tot = sum(int(r[iA]) for r in data)
st = sum(int(r[iS]) for r in data)
print("total warp instr", tot, "stall samples", st, "n sass", len(data))
for k in range(0, len(data), chunk):
    ch = data[k:k + chunk]
    a = sum(int(r[iA]) for r in ch)
    s = sum(int(r[iS]) for r in ch)
    thr = sum(float(r[iT]) * int(r[iA]) for r in ch) / max(a, 1)
    top = max(ch, key=lambda r: int(r[iS]))
    print(f"{k:5d} instr {100*a/tot:5.1f}%  stall {100*s/max(st,1):5.1f}%  thr {thr:4.1f}   top-stall: {top[1].strip()[:60]} ({top[iS]})")
