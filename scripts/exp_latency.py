"""Single-window latency of the reference-shaped calls, graphs on/off (env LSD_TOK_FUSED etc. apply): python scripts/exp_latency.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lipsync_b200 as lb
import bench
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"
for graphs in (False, True):
    p = lb.Predictor(m, use_cuda_graphs=graphs)
    r = bench.latency_probe(p, lb, 50)
    print(json.dumps({"graphs": graphs, "tok_fused": os.environ.get("LSD_TOK_FUSED", "1"), "infer_ms": r["infer_confidence_ms"], "smoothed_ms": r["temporal_smoothed_confidence_ms"]}), flush=True)
# device-side time of one B=1 forward (events)
v, a = lb.synthetic_windows(7, 1); v, a = v.cuda(), a.cuda()
for _ in range(5): m(v, a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): m(v, a)
e1.record(); torch.cuda.synchronize()
print("B=1 forward, back to back on one stream: %.3f ms each" % (e0.elapsed_time(e1) / 20), flush=True)
