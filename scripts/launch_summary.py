"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the last forward of the run, per launch."""
import csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1), r["Grid Size"]) for r in rows]
idx = [i for i, (n, _, _) in enumerate(names) if "video_rows" in n]
start = idx[-1]
tot = 0
agg = {}
for n, v, g in names[start:]:
    u = v / 1000.0
    tot += u
    short = n.split("(")[0].replace("void ", "").replace("lsd::", "")
    agg.setdefault(short, [0, 0.0])
    agg[short][0] += 1; agg[short][1] += u
    if "-v" in sys.argv: print(f"{u:9.1f} us  grid={g:16s} {short[:50]}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us  n={v[0]:3d}  {k}")
print(f"total {tot:.1f} us over {len(names) - start} launches")
