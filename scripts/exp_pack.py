"""score_batches on pinned fp32 k/255 windows: ms per step and the pack times, by pack thread count (1 GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lipsync_b200 as lb
dev = torch.device("cuda", 0)
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0)); m.to(dev).eval(); m.compute_precision = "bf16"
B = 64
_, ah = lb.synthetic_windows(100, 4)
g = torch.Generator().manual_seed(100)
vh = (torch.randint(0, 256, (4, 3, 32, 96, 96), dtype=torch.uint8, generator=g).to(torch.float32) / 255.0).repeat(B // 4, 1, 1, 1, 1).contiguous().pin_memory()
ah = ah.repeat(B // 4, 1, 1, 1).contiguous().pin_memory()
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
for threads in (8, 12, 16, 20, 24, 32):
    p = lb.Predictor(m, batch_size=B, host_transport="u8", host_pack_threads=threads)
    p.score_batches((vh, ah) for _ in range(4))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p.score_batches((vh, ah) for _ in range(30))
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 30
    L = lb._cabi.lib()
    print(f"threads {threads:2d}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} windows/s, last pack {L.lsd_host_pack_last_ms():.2f} ms", flush=True)
