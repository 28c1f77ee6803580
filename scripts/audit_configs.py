#!/usr/bin/env python
"""Audits of BASELINE.json configs[2], configs[3] and configs[4] on one B200 (configs[4] also under torchrun).

  python scripts/audit_configs.py --config 3          # log-mel front end + audio encoder sweep, B = 1 .. 512
  python scripts/audit_configs.py --config 4          # cross-modal attention + temporal transformer, B = 256
  python scripts/audit_configs.py --config 5 [--windows 10000]     # long-video path: uint8 track -> device-built windows
  python -m torch.distributed.run --nproc-per-node N ... scripts/audit_configs.py --config 5   # sharded over N GPUs + logit gather
  python scripts/audit_configs.py --config pcie       # pinned H2D bandwidth of this box (bounds bench.py's e2e)

Every measurement: >= 3 warm-ups, CUDA events on the launch stream, synchronize on both sides, inputs resident in HBM.
One JSON object per line on stdout (rank 0).  Algorithmic bytes / FLOPs per unit are the figures of SURVEY.md §8d.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HBM_PEAK_GBS, BF16_PEAK_TF = 6543.7, 1399.4
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        _p = json.load(fh)
    HBM_PEAK_GBS = float(_p.get("hbm_gbs", HBM_PEAK_GBS))
    BF16_PEAK_TF = float(_p.get("bf16_tflops_sustained", _p.get("bf16_tflops", BF16_PEAK_TF)))
except Exception:
    pass


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make_model(dev):
    import lipsync_b200 as lb
    m = lb.LipSyncModel()
    m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True)
    m.to(dev).eval()
    m.compute_precision = "bf16"
    return m


def config3(args):
    """log-mel (122 880 algorithmic bytes per 128-frame clip: 81 920 B fp32 PCM in + 40 960 B fp32 mel out) and the audio
    2-D ResNet encoder (0.450 GFLOP per clip) for B = 1 .. 512."""
    import lipsync_b200 as lb
    dev = torch.device("cuda", 0)
    m = make_model(dev)
    g = torch.Generator().manual_seed(3)
    pcm_all = (0.1 * torch.randn(512, 20480, generator=g)).to(dev)
    for B in args.batches:
        clips = pcm_all[:B]
        iters = 20 if B <= 64 else 5
        ms_mel = timed(lambda: lb.logmel_db(clips), iters)
        mels = lb.logmel_db(clips)                                                     # (B,80,129)
        audio = mels[:, :, :128].unsqueeze(1).contiguous()                             # (B,1,80,128): target_frames=128
        ms_enc = timed(lambda: m.encode_audio(audio), iters)
        out = m.encode_audio(audio)
        assert out.shape == (B, 256, 16) and bool(torch.isfinite(out).all())
        mel_bytes = 122880.0 * B
        print(json.dumps({
            "config": 3, "B": B,
            "logmel_ms": ms_mel, "logmel_clips_per_s": B / ms_mel * 1e3, "logmel_GBps": mel_bytes / ms_mel / 1e6,
            "logmel_frac_hbm": mel_bytes / ms_mel / 1e6 / HBM_PEAK_GBS,
            "audio_encoder_ms": ms_enc, "audio_encoder_clips_per_s": B / ms_enc * 1e3,
            "audio_encoder_TFLOPs": 0.450e9 * B / ms_enc / 1e9, "audio_encoder_frac_tensor": 0.450e9 * B / ms_enc / 1e9 / BF16_PEAK_TF,
            "audio_encoder_GBps": (1.6e6 * B + 4.9e6) / ms_enc / 1e6,
        }), flush=True)


def config4(args):
    """cross-modal attention + temporal transformer at B=256: 0.336 GFLOP per window (86 GFLOP per pass)."""
    dev = torch.device("cuda", 0)
    m = make_model(dev)
    g = torch.Generator().manual_seed(4)
    for B in [64, 256]:
        v = torch.randn(B, 32, 256, generator=g).to(dev)
        a = torch.randn(B, 16, 256, generator=g).to(dev)
        n0 = m.launch_count() if m._lsd_handle is not None else 0
        ms = timed(lambda: m.fuse_tokens(v, a), 20)
        fused, cls = m.fuse_tokens(v, a)
        assert cls.shape == (B, 256) and bool(torch.isfinite(cls).all())
        print(json.dumps({"config": 4, "B": B, "ms": ms, "windows_per_s": B / ms * 1e3, "TFLOPs": 0.336e9 * B / ms / 1e9,
                          "frac_tensor": 0.336e9 * B / ms / 1e9 / BF16_PEAK_TF,
                          "note": "token GEMMs run split-bf16 (3 MMAs per product): executed tensor FLOPs are 3x the algorithmic figure"}),
              flush=True)


def config5(args):
    """10k sliding windows of one uint8 track, contiguous block partition over the ranks, one all-gather of fp32 logits,
    host confidence aggregation on rank 0 (predictor.py:554-580 + :246-260)."""
    import torch.distributed as dist
    import lipsync_b200 as lb
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = make_model(dev)
    pred = lb.Predictor(m, batch_size=64)
    n_windows, stride, T = args.windows, 8, 32
    n_frames = stride * (n_windows - 1) + T + 16
    lo, hi = lb.partition_windows(n_windows, ws, rank)
    # every rank synthesises the same track deterministically, block by block, and keeps only its own span (+ halo) on its GPU
    f_lo, f_hi = stride * lo, stride * max(hi - 1, lo) + T
    g = torch.Generator(device=dev).manual_seed(5)
    track = torch.empty(f_hi - f_lo, 96, 96, 3, dtype=torch.uint8, device=dev)
    blk = 4096
    for f0 in range(0, n_frames, blk):          # same random stream on every rank -> identical track across shardings
        chunk = torch.randint(0, 256, (min(blk, n_frames - f0), 96, 96, 3), dtype=torch.uint8, device=dev, generator=g)
        a, b = max(f0, f_lo), min(f0 + chunk.shape[0], f_hi)
        if b > a:
            track[a - f_lo:b - f_lo] = chunk[a - f0:b - f0]
    ta_full = int(n_frames / 15 * 100)
    gm = torch.Generator().manual_seed(6)
    mel = (-80.0 * torch.rand(1, 80, ta_full, generator=gm)).to(dev)
    starts_abs = [stride * i for i in range(lo, hi)]

    def score_range(_lo, _hi):
        # local starts relative to this rank's span; the audio alignment needs absolute frame indices, so the rank-local
        # call passes the absolute starts and an offset track view (the kernel reads track[start - f_lo])
        return pred.score_track_logits(track, [s - f_lo for s in starts_abs], mel, n_frames, audio_starts_from=starts_abs)

    def run():
        return pred.score_windows_sharded(n_windows, score_range, ws, rank)

    run()
    torch.cuda.synchronize(dev)
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    logits = run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        lg = logits.cpu()
        confs = [pred._calibrate(float(x)) for x in lg.tolist()]
        t0 = time.perf_counter()
        agg = pred._robust_confidence(confs)
        host_ms = (time.perf_counter() - t0) * 1e3
        import hashlib
        digest = hashlib.sha256(lg.numpy().tobytes()).hexdigest()[:16]
        print(json.dumps({"config": 5, "n_gpus": ws, "windows": n_windows, "ms": ms, "windows_per_s": n_windows / ms * 1e3,
                          "TFLOPs": 31.29e9 * n_windows / ms / 1e9, "aggregate_confidence": agg, "host_aggregation_ms": host_ms,
                          "logits_sha256_16": digest, "fake_votes": int((lg < 0).sum()),
                          "note": "track resident in HBM (uint8, rank-local span); windows built on device; one NCCL all-gather of fp32 logits"}),
              flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def pcie(args):
    dev = torch.device("cuda", 0)
    h = torch.empty(229113856 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device=dev)
    ms = timed(lambda: d.copy_(h, non_blocking=True), 10)
    print(json.dumps({"config": "pcie", "bytes": h.numel() * 4, "ms": ms, "GBps": h.numel() * 4 / ms / 1e6,
                      "e2e_bound_windows_per_s": 64 / ms * 1e3}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--windows", type=int, default=10000)
    ap.add_argument("--batches", type=lambda v: [int(x) for x in v.split(",")], default=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512])
    a = ap.parse_args()
    import __graft_entry__ as ge
    if int(os.environ.get("RANK", "0")) == 0:
        ge.build()
    {"3": config3, "4": config4, "5": config5, "pcie": pcie}[a.config](a)
