"""Experiment: timeline of Predictor.score_batches (copy stream vs compute stream) for 12 steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge; ge.build()
import lipsync_b200 as lb
dev = torch.device("cuda", 0)
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to(dev).eval(); m.compute_precision = "bf16"
v, a = lb.synthetic_windows(1, 4)
vh = v.repeat(16, 1, 1, 1, 1).contiguous().pin_memory(); ah = a.repeat(16, 1, 1, 1).contiguous().pin_memory()
pred = lb.Predictor(m, batch_size=64)
pred.score_batches((vh, ah) for _ in range(4))
torch.cuda.synchronize()
comp = torch.cuda.current_stream(dev); copy = torch.cuda.Stream(dev)
NS, K = 3, 12
slots = [(torch.empty_like(vh, device=dev), torch.empty_like(ah, device=dev)) for _ in range(NS)]
E = lambda: torch.cuda.Event(enable_timing=True)
cs, ce, fs, fe = [E() for _ in range(K)], [E() for _ in range(K)], [E() for _ in range(K)], [E() for _ in range(K)]
free = [torch.cuda.Event() for _ in range(NS)]
for e in free: e.record(comp)
t0 = E(); t0.record(comp); torch.cuda.synchronize()
cpu = []
w0 = time.perf_counter()
for k in range(K):
    s = k % NS
    c0 = time.perf_counter()
    with torch.cuda.stream(copy):
        copy.wait_event(free[s]); cs[k].record(copy)
        slots[s][0].copy_(vh, non_blocking=True); slots[s][1].copy_(ah, non_blocking=True)
        ce[k].record(copy)
    comp.wait_event(ce[k]); fs[k].record(comp)
    lg = m(slots[s][0], slots[s][1])
    fe[k].record(comp); free[s].record(comp)
    cpu.append((time.perf_counter() - c0) * 1e3)
torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - w0) * 1e3, "cpu enqueue ms/step", sum(cpu) / K)
for k in range(K):
    print(k, "copy %.2f-%.2f" % (t0.elapsed_time(cs[k]), t0.elapsed_time(ce[k])), "fwd %.2f-%.2f" % (t0.elapsed_time(fs[k]), t0.elapsed_time(fe[k])), "cpu %.2f" % cpu[k])
