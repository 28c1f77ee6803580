"""e2e (score_batches) under torchrun: per-step host timing of the u8 transport, CPU budget of the container."""
import os, sys, time, json
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import lipsync_b200 as lb
from lipsync_b200 import _cabi

rank = int(os.environ.get("RANK", "0")); ws = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
def cg():
    try: return open("/sys/fs/cgroup/cpu.max").read().strip()
    except Exception as e: return str(e)
info = {"rank": rank, "cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "cgroup_cpu_max": cg(), "omp": os.environ.get("OMP_NUM_THREADS")}
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); m.to(f"cuda:{local}").eval(); m.compute_precision = "bf16"
B = 64
vh = (torch.randint(0, 256, (B, 3, 32, 96, 96), dtype=torch.uint8).float() / 255.0).pin_memory()
_, ah = lb.synthetic_windows(3, B); ah = ah.pin_memory()
L = _cabi.lib()
d = torch.empty(vh.shape, dtype=torch.uint8).pin_memory()
res = {}
for th in (2, 4, 8, 12, 16):
    L.lsd_host_pack_u8_exact(vh.data_ptr(), d.data_ptr(), vh.numel(), th)
    if ws > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(5): L.lsd_host_pack_u8_exact(vh.data_ptr(), d.data_ptr(), vh.numel(), th)
    res[th] = round((time.perf_counter() - t) / 5 * 1e3, 2)
info["pack_ms_by_threads_all_ranks_concurrently"] = res
for mode, th in (("u8", None), ("u8", 4), ("u8", 8), ("fp32", None)):
    p = lb.Predictor(m, batch_size=B, host_transport=mode, host_pack_threads=th)
    p.score_batches((vh, ah) for _ in range(5))
    if ws > 1: dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    p.score_batches((vh, ah) for _ in range(20))
    info[f"e2e_ms_per_step[{mode},{p.host_pack_threads if mode == 'u8' else '-'}]"] = round((time.perf_counter() - t) / 20 * 1e3, 3)
print(json.dumps(info), flush=True)
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
