"""Top stall instructions from `ncu -i X.ncu-rep --page source --csv` output (one section per profiled launch)."""
import csv, sys
path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
s = secs[which]; hdr, data = s["hdr"], s["data"]
col, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(float(r[col] or 0) for r in data)
print(s["name"], "sections", len(secs), "total samples", tot, "instrs", len(data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for i in stall_cols:
    v = sum(float(r[i] or 0) for r in data)
    if v > 0.01 * tot: print(f"  {hdr[i]:28s} {100*v/tot:5.1f}%")
for idx, r in sorted(enumerate(data), key=lambda ir: -float(ir[1][col] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    reasons = sorted([(float(r[i] or 0), hdr[i]) for i in stall_cols], reverse=True)[:1]
    print(f"#{idx:5d} {float(r[col]):8.0f} {100*float(r[col])/tot:5.1f}% exec={r[ex]:>8s} {r[src].strip()[:80]:80s} {reasons[0][1]}")
