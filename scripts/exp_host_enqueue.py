"""How long does the host take to enqueue one forward (81 launches)?  If it is close to the device time of a step the
path is launch-bound on the CPU and the GPU idles between kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge; ge.build()
import lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to('cuda').eval(); m.compute_precision = 'bf16'
v, a = lb.synthetic_windows(1, 4)
v = v.repeat(16, 1, 1, 1, 1).cuda(); a = a.repeat(16, 1, 1, 1).cuda()
for _ in range(3): m(v, a)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(20): m(v, a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0)/20:.3f} ms per forward; wall incl. sync {1e3*(t2-t0)/20:.3f} ms per forward", flush=True)
