"""Per-launch duration + tensor-pipe utilisation of the last forward in an ncu launch list with two metrics."""
import csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
byid = {}
for r in csv.DictReader(lines):
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ids = sorted(byid, key=int)
start = [i for i in ids if "video_rows" in byid[i]["name"]][-1]
tot = 0; agg = {}
for i in ids[ids.index(start):]:
    d = byid[i]
    us = d["gpu__time_duration.sum"] / 1000
    tot += us
    short = d["name"].split("(")[0].replace("void ", "").replace("lsd::", "")
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += us
    if "-v" in sys.argv and us > 15:
        print(f"{us:9.1f} us tensor={d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}% grid={d['grid']:14s} {short[:30]}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"{v[1]:9.1f} us  n={v[0]:3d}  {k}")
print(f"total {tot:.1f} us")
