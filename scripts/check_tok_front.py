"""A/B check of the fused token-path front (tok_front.cu) against the launch-by-launch chain and the oracle.
Run on a B200:  python scripts/check_tok_front.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lipsync_b200 as lb
from oracle import lipsync_oracle as orc

sd = lb.make_synthetic_state_dict(0)
m = lb.LipSyncModel(); m.load_state_dict(sd, strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"

def run(B, T, TA, front, fused="1"):
    os.environ["LSD_TOK_FRONT"] = "1" if front else "0"
    os.environ["LSD_TOK_FUSED"] = fused
    g = torch.Generator().manual_seed(4)
    v = torch.randn(B, T, 256, generator=g); a = torch.randn(B, TA, 256, generator=g)
    f, c = m.fuse_tokens(v.cuda(), a.cuda())
    torch.cuda.synchronize()
    return v, a, f.cpu(), c.cpu()

def timeit(B, T, TA, front, n=20):
    os.environ["LSD_TOK_FRONT"] = "1" if front else "0"
    os.environ["LSD_TOK_FUSED"] = "1"
    g = torch.Generator().manual_seed(4)
    v = torch.randn(B, T, 256, generator=g).cuda(); a = torch.randn(B, TA, 256, generator=g).cuda()
    for _ in range(3): m.fuse_tokens(v, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m.fuse_tokens(v, a)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

bad = 0
for (B, T, TA) in [(1, 32, 16), (2, 32, 16), (3, 32, 16), (5, 32, 16), (64, 32, 16), (4, 16, 8), (7, 16, 8), (3, 29, 16), (2, 40, 20)]:
    v, a, f0, c0 = run(B, T, TA, False, "0")          # pure chain
    _, _, f1, c1 = run(B, T, TA, True, "0")           # fused front + chain layers: isolates the front
    _, _, f2, c2 = run(B, T, TA, True, "1")           # both fused kernels
    with torch.no_grad():
        fr = orc.cross_modal(sd, v, a); cr = orc.temporal(sd, fr)
    fs, cs = float(fr.abs().max()), float(cr.abs().max())
    print(f"B={B} T={T}: fused |max| {fs:.2f}: chain {float((f0-fr).abs().max()):.2e} front {float((f1-fr).abs().max()):.2e} | "
          f"cls |max| {cs:.2f}: chain {float((c0-cr).abs().max()):.2e} front+chain {float((c1-cr).abs().max()):.2e} front+fused {float((c2-cr).abs().max()):.2e} "
          f"finite={bool(torch.isfinite(c2).all() and torch.isfinite(f1).all())}", flush=True)
    if not (float((f1 - fr).abs().max()) <= 2e-2 * max(1.0, fs) and float((c2 - cr).abs().max()) <= 3e-2 * max(1.0, cs)): bad += 1
# batch-composition independence: window i alone == window i inside a batch
v, a, f_all, c_all = run(5, 32, 16, True, "1")
for i in range(5):
    fi, ci = m.fuse_tokens(v[i:i + 1].cuda(), a[i:i + 1].cuda())
    same = bool((ci.cpu() == c_all[i:i + 1]).all() and (fi.cpu() == f_all[i:i + 1]).all())
    print(f"window {i}: alone == in batch of 5: {same}", flush=True)
    bad += 0 if same else 1
for B in (1, 4, 64, 256):
    print(f"B={B}: chain front {timeit(B, 32, 16, False)*1e3:.0f} us   fused front {timeit(B, 32, 16, True)*1e3:.0f} us  (transformer fused in both)", flush=True)
print("RESULT", "FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
