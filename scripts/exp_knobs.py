"""Device time of one B=64 forward (inputs resident) under the tuning knobs currently in the environment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(1, 4)
v = v.repeat(16, 1, 1, 1, 1).cuda(); a = a.repeat(16, 1, 1, 1).cuda()
for _ in range(5): m(v, a)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m(v, a)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print(f"{os.environ.get('TAG', '')}: {best:.3f} ms per step = {64e3 / best:.0f} windows/s", flush=True)
