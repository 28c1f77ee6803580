"""Debug helper (GPU): stage-by-stage error of the bf16 tensor-core path against the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
import lipsync_b200 as lb
from oracle import lipsync_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = lb.make_synthetic_state_dict(0)
video, audio = lb.synthetic_windows(1, B)
inter = {}
ref = orc.forward(sd, video, audio, inter=inter)
m = lb.LipSyncModel(); m.load_state_dict(sd); m.to("cuda").eval()
m.compute_precision = "bf16"
t0 = time.time()
out, aux = m(video.cuda(), audio.cuda(), return_aux=True)   # v_emb / a_emb / cls exist only as aux outputs on the fused token path
out = out.cpu()
torch.cuda.synchronize()
print("forward done in", time.time() - t0, flush=True)
def rel(a, b): return float((a - b).abs().max()) / max(1e-12, float(b.abs().max()))
vf = m.stage("v_feat").cpu().view(B, -1, 256)
print("v_feat rel", rel(vf, inter["v_feat"].transpose(1, 2)))
comb = m.stage("comb").cpu().view(B, 448)
for i, k in enumerate(["art_raw", "art_delta", "art_hf"]):
    print(k, "rel", rel(comb[:, 256 + 64 * i: 320 + 64 * i], inter[k]))
print("cls rel", rel(aux["cls_output"].cpu(), inter["cls"]))
TA = inter["a_emb"].shape[1]
print("a_feat rel", rel(m.stage("a_feat").cpu().view(B, TA, 256), inter["a_feat"].transpose(1, 2)))
print("v_emb rel", rel(aux["visual_tokens"].cpu(), inter["v_emb"]))
print("a_emb rel", rel(aux["audio_tokens"].cpu(), inter["a_emb"]))
print("fused rel", rel(m.stage("fused").cpu().view(B, -1, 256), inter["fused"]))
print("t_preconv rel", rel(m.stage("tok").cpu().view(B, 33, 256)[:, 1:], inter["t_layer3"][:, 1:]), "(tok final vs t_layer3)")
print("logits", out.tolist(), "ref", ref.tolist(), "max abs", float((out - ref).abs().max()))
