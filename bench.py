#!/usr/bin/env python
"""Benchmark of the window-scoring hot path (BASELINE.json: windows/sec at 1/2/4/8 B200 + roofline fraction,
next to the reference's CPU path timed on the box's own host cores).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one JSON line on rank 0)
  python bench.py --impl reference --gpus N ...             # the reference algorithm on the host cores (oracle port)

A "step" is one pass of the hot path over one batch of B=64 synthetic canonical windows per GPU
(video (64,3,32,96,96) + log-mel (64,1,80,128); BASELINE.json configs[1]).  `value` is timed with the inputs already
resident in HBM; `e2e` is the same metric through the public API with pinned HOST buffers (H2D of the windows and
D2H of the logits inside the timed region).  Weak scaling: every rank scores its own batch; under N>1 the per-step
logits are all-gathered (NCCL) as in the long-video path (the only collective on the path).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW = 31.29e9   # 2*MACs of LipSyncModel.forward at the canonical window
BATCH = 64
WORKLOAD = "batched window scoring B=64 per GPU, canonical window video(3,32,96,96)+logmel(1,80,128), random-init R2Plus1D-Sync"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _dist():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return ws, rank, local


def cpu_reference_run(steps: int, warmup: int, sample_windows: int):
    """The reference algorithm (oracle port of LipSyncModel.forward, fp32) on all host cores: windows/s."""
    import lipsync_b200 as lb
    from oracle import lipsync_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    sd = lb.make_synthetic_state_dict(0)
    video, audio = lb.synthetic_windows(1, sample_windows)
    for _ in range(warmup):
        orc.forward(sd, video[:1], audio[:1])
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward(sd, video, audio)
    dt = time.perf_counter() - t0
    return steps * sample_windows / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, out=sys.stdout):
    ws, rank, _ = _dist()
    if rank != 0:
        return
    sample = 4
    steps = max(1, min(args.steps, 3))
    wps, ms, cores = cpu_reference_run(steps, min(args.warmup, 1), sample)
    line = {
        "impl": "reference", "metric": "windows_per_sec", "value": wps, "unit": "windows/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference is pure Python/PyTorch; timed through the CPU oracle port of LipSyncModel.forward"},
        "cpu_baseline": {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} canonical windows, fp32, torch CPU, all host threads"},
        "e2e": {"value": wps, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Keep a private handle on the real stdout for the single JSON line and point fd 1 at stderr: libraries (NCCL's version
    banner, nvcc through build()) write to fd 1 behind Python's back."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out)

    import __graft_entry__ as ge
    ws, rank, local = _dist()
    if rank == 0:
        ge.build()
    import torch.distributed as dist
    if ws > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    ge.build()
    import lipsync_b200 as lb

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, K, W = args.batch, args.steps, max(3, args.warmup)
    model = lb.LipSyncModel()
    model.load_state_dict(lb.make_synthetic_state_dict(0), strict=True)
    model.to(dev).eval()
    model.compute_precision = args.precision
    pred = lb.Predictor(model, batch_size=B)

    # synthetic inputs (host, pinned) and their device-resident copies; distinct data per rank
    vh, ah = lb.synthetic_windows(100 + rank, 4)
    vh = vh.repeat(B // 4 + 1, 1, 1, 1, 1)[:B].contiguous().pin_memory()
    ah = ah.repeat(B // 4 + 1, 1, 1, 1)[:B].contiguous().pin_memory()
    vd, ad = vh.to(dev), ah.to(dev)
    # N > 1: every rank keeps the logits of its K batches and ONE all-gather at the end of the timed region hands all of them
    # to every rank — the path's only exchange step (long videos gather per-window logits once, before the aggregation)
    local_logits = torch.empty(max(K, W), B, dtype=torch.float32, device=dev)
    gathered = torch.empty(ws * max(K, W) * B, dtype=torch.float32, device=dev) if ws > 1 else None
    gathered_e2e = torch.empty(ws * B, dtype=torch.float32, device=dev) if ws > 1 else None
    step_no = [0]

    def step_resident():
        logits = model(vd, ad)
        if ws > 1:
            local_logits[step_no[0] % local_logits.shape[0]].copy_(logits)
            step_no[0] += 1
        return logits

    def gather_resident():
        if ws > 1:
            dist.all_gather_into_tensor(gathered, local_logits.view(-1))

    def run_e2e(steps):
        """Public API on HOST buffers: Predictor.score_batches uploads every step's windows from pinned host memory
        (copy stream, overlapped with the previous step's scoring) and reads every step's logits back."""
        outs = pred.score_batches((vh, ah) for _ in range(steps))
        if ws > 1:
            last = outs[-1].to(dev)
            dist.all_gather_into_tensor(gathered_e2e, last)
            return gathered_e2e.cpu()
        return outs[-1]

    def timed(fn, steps, profile=False, whole=False, finalize=None):
        if whole:
            fn(W)
        else:
            for _ in range(W):
                fn()
        if finalize is not None:
            finalize()
        torch.cuda.synchronize(dev)
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(local)
        sampler.start()
        if profile:
            model.profile_enable(2 if args.precision == 'bf16' else 1)
        n0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn(steps)
        else:
            for _ in range(steps):
                fn()
        if finalize is not None:
            finalize()
        e1.record()
        torch.cuda.synchronize(dev)
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        launches = model.launch_count() - n0
        prof = model.profile_get() if profile else None
        if profile:
            model.profile_enable(False)
        sampler.stop_flag = True
        sampler.join()
        if ws > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, sampler.result()

    # uint8-track variant of the end-to-end path: per step the host hands over the uint8 mouth-crop span of
    # the step's 64 windows (stride 8: 536 frames, 14.8 MB instead of 226 MB of fp32 windows) from pinned memory, the windows
    # are built on the device (lsd_score_windows), and the logits come back to the host.
    n_tf = 8 * (B - 1) + 32
    gt = torch.Generator().manual_seed(200 + rank)
    track_h = torch.randint(0, 256, (n_tf, 96, 96, 3), dtype=torch.uint8, generator=gt).pin_memory()
    mel_h = (-80.0 * torch.rand(1, 80, int(n_tf / 15 * 100), generator=gt)).pin_memory()
    starts = [8 * i for i in range(B)]
    track_slots = [torch.empty_like(track_h, device=dev) for _ in range(2)]
    mel_slots = [torch.empty_like(mel_h, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)

    def run_e2e_track(steps):
        comp = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in free:
            e.record(comp)
        host_out = []
        for k in range(steps):
            sl = k & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[sl])
                track_slots[sl].copy_(track_h, non_blocking=True)
                mel_slots[sl].copy_(mel_h, non_blocking=True)
                ready[sl].record(copy_stream)
            comp.wait_event(ready[sl])
            lg = pred.score_track_logits(track_slots[sl], starts, mel_slots[sl], n_tf)
            free[sl].record(comp)
            ho = torch.empty(B, dtype=torch.float32, pin_memory=True)
            ho.copy_(lg, non_blocking=True)
            host_out.append(ho)
        comp.synchronize()
        return host_out[-1]

    ms, launches, prof, clocks = timed(step_resident, K, profile=True, finalize=gather_resident)
    ms_e2e, _, _, _ = timed(run_e2e, K, whole=True)
    ms_trk, _, _, _ = timed(run_e2e_track, K, whole=True)
    value = ws * B * K / (ms / 1e3)
    e2e = ws * B * K / (ms_e2e / 1e3)
    e2e_trk = ws * B * K / (ms_trk / 1e3)

    if rank == 0:
        peaks, peak_src = _peaks()
        kern_ms, kern_n, kern_flops = prof
        # dominant kernel class: the implicit-GEMM convolution/linear kernel (tensor-bound on the bf16 path)
        achieved = (kern_flops / max(kern_n, 1)) / ((kern_ms / max(kern_n, 1)) * 1e-3) / 1e12 if kern_n else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        line = {
            "metric": "windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": ws, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "precision": args.precision,
                       "l2": "inputs (226 MB fp32 video per step) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"dp{ws} (independent windows; all-gather of fp32 logits only)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "windows/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": vh.numel() * 4 + ah.numel() * 4, "d2h_bytes_per_step": (ws if ws > 1 else 1) * B * 4,
                    "api": "Predictor.score_batches on pinned host fp32 windows (per step: H2D of the windows on a copy stream, forward, D2H of the logits)"},
            "e2e_track_u8": {"value": e2e_trk, "unit": "windows/s", "ms_per_step": ms_trk / K,
                             "h2d_bytes_per_step": track_h.numel() + mel_h.numel() * 4, "d2h_bytes_per_step": B * 4,
                             "api": "Predictor.score_track_logits on a pinned host uint8 mouth-crop track (windows built on the device, lsd_score_windows)"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                         "kernel": ("umma_conv_kernel, the 9 launches of the 3-D conv visual encoder (stem + layer1-4: 26.77 of the 31.29 "
                                    "GFLOP per window); CUDA events on the launch stream"
                                    if args.precision == "bf16" else "conv_f32_kernel (all launches, CUDA events on the launch stream)"),
                         "kernel_ms_per_step": kern_ms / K, "kernel_launches_per_step": kern_n / K,
                         "kernel_share_of_step": kern_ms / ms,
                         "whole_model_tflops": FLOP_PER_WINDOW * B * K / (ms / 1e3) / 1e12},
        }
        if ws == 1 and not args.no_cpu_baseline:
            wps, cms, cores = cpu_reference_run(2, 1, 4)
            line["cpu_baseline"] = {"value": wps, "unit": "windows/s", "cores": cores, "kind": "port",
                                    "sample": "2 steps x 4 canonical windows through the CPU oracle port (fp32 torch, all host threads)"}
        print(json.dumps(line), file=out, flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
