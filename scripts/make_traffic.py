"""profiles/traffic.json from an ncu launch list that carries dram__bytes_{read,write}.sum: DRAM traffic per launch of the
dominant kernel class (the 9 tcgen05 launches of the 3-D conv visual encoder in one forward: stem_ring_kernel + 8 x umma_conv_kernel)."""
import csv, json, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
byid = {}
for r in csv.DictReader(lines):
    d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r["Metric Name"]] = v
ids = sorted(byid, key=int)
start = [i for i in ids if "video_rows" in byid[i]["name"]][-1]
umma = [byid[i] for i in ids[ids.index(start):] if "umma_conv" in byid[i]["name"] or "stem_ring" in byid[i]["name"]][:9]
tot = [d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in umma]
out = {"kernel": "stem_ring_kernel + umma_conv_kernel (stem + layer1-4 of the visual encoder, B=64)", "launches": len(umma),
       "dram_bytes_per_launch": sum(tot) / len(tot), "per_launch": [
           {"us": d["gpu__time_duration.sum"] / 1e3, "dram_read_MB": d["dram__bytes_read.sum"] / 1e6, "dram_write_MB": d["dram__bytes_write.sum"] / 1e6,
            "tensor_pipe_pct": d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")} for d in umma],
       "source": path, "note": "ncu --metrics pass, --clock-control none; per-launch times are cold-cache and serialised"}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out)[:600])
