"""A/B check of the fused temporal-transformer kernel (tok_fused.cu) against the layer-by-layer GEMM chain and the oracle.
Run on a B200:  python scripts/check_tok_fused.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import lipsync_b200 as lb
from oracle import lipsync_oracle as orc

sd = lb.make_synthetic_state_dict(0)
m = lb.LipSyncModel(); m.load_state_dict(sd, strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"

def run(B, T, TA, fused):
    os.environ["LSD_TOK_FUSED"] = "1" if fused else "0"
    g = torch.Generator().manual_seed(4)
    v = torch.randn(B, T, 256, generator=g); a = torch.randn(B, TA, 256, generator=g)
    f, c = m.fuse_tokens(v.cuda(), a.cuda())
    torch.cuda.synchronize()
    return v, a, f.cpu(), c.cpu()

def timeit(B, T, TA, fused, n=20):
    os.environ["LSD_TOK_FUSED"] = "1" if fused else "0"
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
for (B, T, TA) in [(1, 32, 16), (2, 32, 16), (3, 32, 16), (5, 32, 16), (64, 32, 16), (4, 16, 8), (7, 16, 8)]:
    v, a, f0, c0 = run(B, T, TA, False)
    _, _, f1, c1 = run(B, T, TA, True)
    with torch.no_grad():
        fused_ref = orc.cross_modal(sd, v, a); cls_ref = orc.temporal(sd, fused_ref)
    d01 = float((c0 - c1).abs().max()); d0r = float((c0 - cls_ref).abs().max()); d1r = float((c1 - cls_ref).abs().max())
    scale = float(cls_ref.abs().max())
    print(f"B={B} T={T}: |cls| max {scale:.3f}  chain-vs-oracle {d0r:.3e}  fused-vs-oracle {d1r:.3e}  fused-vs-chain {d01:.3e}  finite={bool(torch.isfinite(c1).all())}", flush=True)
    if not (d1r <= 3e-2 * max(1.0, scale)): bad += 1
# batch-composition independence of the fused kernel: window i alone == window i inside a batch
v, a, _, c_all = run(5, 32, 16, True)
os.environ["LSD_TOK_FUSED"] = "1"
for i in range(5):
    _, ci = m.fuse_tokens(v[i:i + 1].cuda(), a[i:i + 1].cuda())
    same = bool((ci.cpu() == c_all[i:i + 1]).all())
    print(f"window {i}: alone == in batch of 5: {same}", flush=True)
    bad += 0 if same else 1
for B in (1, 4, 64, 256):
    print(f"B={B}: chain {timeit(B, 32, 16, False)*1e3:.0f} us   fused {timeit(B, 32, 16, True)*1e3:.0f} us", flush=True)
print("RESULT", "FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
