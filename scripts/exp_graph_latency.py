"""Where does the time of a graph-replayed single-window call go, before / after Predictor.score_batches ran in the process?"""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
import lipsync_b200 as lb

m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"
v1, a1 = lb.synthetic_windows(7, 1); v1, a1 = v1[0].numpy(), a1[0].numpy()

def probe(p, tag, n=30):
    for _ in range(5): p._infer_confidence(v1, a1)
    ent = list(p._graphs.values())[0]
    t = {"stack": [], "copy": [], "replay": [], "sync": [], "total": []}
    for _ in range(n):
        t0 = time.perf_counter()
        v = np.stack([v1]); a = np.stack([a1])
        t1 = time.perf_counter()
        ent["vh"].copy_(torch.from_numpy(v)); ent["ah"].copy_(torch.from_numpy(a))
        t2 = time.perf_counter()
        ent["graph"].replay()
        t3 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t4 = time.perf_counter()
        t["stack"].append(t1 - t0); t["copy"].append(t2 - t1); t["replay"].append(t3 - t2); t["sync"].append(t4 - t3); t["total"].append(t4 - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ent["graph"].replay(); e1.record(); torch.cuda.synchronize()
    print(tag, {k: round(sorted(x)[len(x) // 2] * 1e3, 3) for k, x in t.items()}, "device ms of one replay", round(e0.elapsed_time(e1), 3),
          "torch threads", torch.get_num_threads(), flush=True)

mode = sys.argv[1] if len(sys.argv) > 1 else "u8"
p = lb.Predictor(m, batch_size=64, host_transport=mode, host_pack_threads=(int(os.environ["LSD_PACK_THREADS"]) if "LSD_PACK_THREADS" in os.environ else None))
probe(p, "fresh")
vh = (torch.randint(0, 256, (8, 3, 32, 96, 96), dtype=torch.uint8).float() / 255.0).pin_memory()
_, a = lb.synthetic_windows(3, 8)
p.score_batches((vh, a) for _ in range(4))
probe(p, f"after score_batches[{mode}]")
