"""Phase timestamps of the fused temporal-transformer kernel (CTA 0), B=64: LSD_TOKF_TRACE=1 python scripts/trace_tok_fused.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import lipsync_b200 as lb
m = lb.LipSyncModel(); m.load_state_dict(lb.make_synthetic_state_dict(0), strict=True); m.to("cuda:0").eval(); m.compute_precision = "bf16"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(4)
v = torch.randn(B, 32, 256, generator=g).cuda(); a = torch.randn(B, 16, 256, generator=g).cuda()
os.environ.pop("LSD_TOKF_TRACE", None)
for _ in range(3): m.fuse_tokens(v, a)
torch.cuda.synchronize()
os.environ["LSD_TOKF_TRACE"] = "1"
m.fuse_tokens(v, a)
torch.cuda.synchronize()
