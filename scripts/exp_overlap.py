"""Experiment: does the 229 MB/step H2D copy slow the forward (or vice versa) when both run concurrently?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge; ge.build()
import lipsync_b200 as lb
dev = torch.device("cuda", 0)
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to(dev).eval(); m.compute_precision = "bf16"
v, a = lb.synthetic_windows(1, 4)
vh = v.repeat(16, 1, 1, 1, 1).contiguous().pin_memory(); ah = a.repeat(16, 1, 1, 1).contiguous().pin_memory()
vd, ad = vh.to(dev), ah.to(dev)
dst = torch.empty_like(vd)
cs = torch.cuda.Stream(dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(3): m(vd, ad)
torch.cuda.synchronize()
N = 20
# (a) copy alone
e0, e1 = ev(), ev()
with torch.cuda.stream(cs):
    e0.record(cs)
    for _ in range(N): dst.copy_(vh, non_blocking=True)
    e1.record(cs)
torch.cuda.synchronize(); print("copy alone   ms/iter", e0.elapsed_time(e1) / N)
# (b) compute alone
e0, e1 = ev(), ev(); e0.record()
for _ in range(N): m(vd, ad)
e1.record(); torch.cuda.synchronize(); print("compute alone ms/iter", e0.elapsed_time(e1) / N)
# (c) both, independent
c0, c1, k0, k1 = ev(), ev(), ev(), ev()
with torch.cuda.stream(cs):
    c0.record(cs)
    for _ in range(N): dst.copy_(vh, non_blocking=True)
    c1.record(cs)
k0.record()
for _ in range(N): m(vd, ad)
k1.record(); torch.cuda.synchronize()
print("concurrent: copy ms/iter", c0.elapsed_time(c1) / N, " compute ms/iter", k0.elapsed_time(k1) / N)
