"""ORACLE (test infrastructure, not product code): CPU restatement of `LipSyncModel.forward`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product path (`lipsync_b200`) never does and fails loudly without its CUDA library.

Each function restates one reference module in plain `torch.nn.functional` fp32 on the CPU, driven directly
by a reference-layout `state_dict` (no `nn.Module` from the reference is imported, so the file travels to
the GPU box where `/root/reference` does not exist).  Attention, the encoder layer, interpolation, BN folding
and pooling are written out by hand (Appendix A of SURVEY.md) rather than delegated to the fused torch
modules the reference uses.

Parity pin: `tests/golden/make_golden.py` runs the real reference (imported from `/root/reference`) on the
seeded weights/inputs of `state_spec.py` and commits logits + per-stage fingerprints under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against them (fp32, <= 2e-5 abs on logits).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]
BN_EPS = 1e-5  # nn.BatchNorm*d default, never overridden in the reference
LN_EPS = 1e-5  # nn.LayerNorm / TransformerEncoderLayer default


def _bn(sd: SD, p: str, x: Tensor) -> Tensor:
    """Eval-mode BatchNorm: y = (x-mean)/sqrt(var+eps)*gamma+beta over channel dim 1."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(shape) + shift.view(shape)


# ---------------------------------------------------------------- visual encoder (visual_encoder.py:166-201)
def _res_block3d(sd: SD, p: str, x: Tensor, stride: Tuple[int, int, int]) -> Tensor:
    # visual_encoder.py:81-87
    if (p + ".downsample.0.weight") in sd:
        idt = _bn(sd, p + ".downsample.1", F.conv3d(x, sd[p + ".downsample.0.weight"], stride=stride))
    else:
        idt = x
    out = F.relu(_bn(sd, p + ".conv1.1", F.conv3d(x, sd[p + ".conv1.0.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".conv2.1", F.conv3d(out, sd[p + ".conv2.0.weight"], stride=1, padding=1))
    return F.relu(out + idt)


def visual_encoder(sd: SD, x: Tensor, inter: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    if x.dim() != 5:
        raise ValueError(f"VisualEncoder expected input of shape (B, 3, T, H, W), got {tuple(x.shape)}")
    p = "visual_encoder"
    out = F.conv3d(x, sd[p + ".stem.0.weight"], stride=(1, 2, 2), padding=(1, 3, 3))  # :113-121
    out = F.relu(_bn(sd, p + ".stem.1", out))
    if inter is not None:
        inter["v_stem_conv"] = out
    out = F.max_pool3d(out, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))  # :124-128
    if inter is not None:
        inter["v_stem"] = out
    strides = [(1, 1, 1), (1, 2, 2), (1, 2, 2), (1, 2, 2)]  # :133-152
    for i, st in enumerate(strides, start=1):
        out = _res_block3d(sd, f"{p}.layer{i}", out, st)
        if inter is not None:
            inter[f"v_layer{i}"] = out
    fmap = out  # dropout is identity in eval (:154)
    pooled = out.mean(dim=(3, 4))  # adaptive_avg_pool3d(out,(T,1,1)) (:196-198)
    return pooled, fmap


# ---------------------------------------------------------------- audio encoder (audio_encoder.py:173-205)
def _res_block2d(sd: SD, p: str, x: Tensor, stride: Tuple[int, int]) -> Tensor:
    if (p + ".downsample.0.weight") in sd:
        idt = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    else:
        idt = x
    out = F.relu(_bn(sd, p + ".conv1.1", F.conv2d(x, sd[p + ".conv1.0.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".conv2.1", F.conv2d(out, sd[p + ".conv2.0.weight"], stride=1, padding=1))
    return F.relu(out + idt)


def audio_encoder(sd: SD, x: Tensor, inter: Optional[Dict[str, Tensor]] = None) -> Tensor:
    if x.dim() != 4:
        raise ValueError(f"AudioEncoder expected input of shape (B, 1, F, T), got {tuple(x.shape)}")
    p = "audio_encoder"
    out = F.conv2d(x, sd[p + ".stem.0.weight"], stride=(2, 2), padding=3)  # :128-136
    out = F.relu(_bn(sd, p + ".stem.1", out))
    out = F.max_pool2d(out, kernel_size=3, stride=(2, 2), padding=1)  # :139
    if inter is not None:
        inter["a_stem"] = out
    strides = [(1, 1), (2, 2), (2, 1), (2, 1)]  # preserve_audio_temporal=True (:144-156)
    for i, st in enumerate(strides, start=1):
        out = _res_block2d(sd, f"{p}.layer{i}", out, st)
        if inter is not None:
            inter[f"a_layer{i}"] = out
    return out.mean(dim=2)  # mean over F' (:202-204) -> (B, 256, T')


# ---------------------------------------------------------------- attention helpers (Appendix A)
def _mha(sd: SD, p: str, q_in: Tensor, kv_in: Tensor, heads: int = 8) -> Tensor:
    """nn.MultiheadAttention(batch_first=True), eval: packed in_proj rows [Q|K|V], softmax(QK^T/sqrt(hd))V, out_proj."""
    d = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = q_in @ w[0:d].t() + b[0:d]
    k = kv_in @ w[d : 2 * d].t() + b[d : 2 * d]
    v = kv_in @ w[2 * d : 3 * d].t() + b[2 * d : 3 * d]
    bsz, tq, _ = q.shape
    tk = k.shape[1]
    hd = d // heads
    q = q.view(bsz, tq, heads, hd).transpose(1, 2)
    k = k.view(bsz, tk, heads, hd).transpose(1, 2)
    v = v.view(bsz, tk, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(bsz, tq, d)
    return o @ sd[p + ".out_proj.weight"].t() + sd[p + ".out_proj.bias"]


def _layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))  # approximate='none'


def lerp_tokens(a: Tensor, t_out: int) -> Tensor:
    """F.interpolate(mode='linear', align_corners=False) along the token axis of (B, T_in, D)."""
    t_in = a.shape[1]
    if t_in == t_out:
        return a
    scale = t_in / t_out
    idx = torch.arange(t_out, dtype=torch.float32, device=a.device)
    src = ((idx + 0.5) * scale - 0.5).clamp_(min=0.0)
    i0 = src.floor().to(torch.int64).clamp_(max=t_in - 1)
    i1 = (i0 + 1).clamp_(max=t_in - 1)
    w1 = (src - i0.to(torch.float32)).view(1, -1, 1)
    return a[:, i0] * (1.0 - w1) + a[:, i1] * w1


# ---------------------------------------------------------------- projection + cross-modal (fusion_module.py)
def projection(sd: SD, v_feat: Tensor, a_feat: Tensor) -> Tuple[Tensor, Tensor]:
    if v_feat.dim() != 3 or a_feat.dim() != 3:
        raise ValueError("FeatureProjection expects visual_feat and audio_feat of shape (B, D, T)")
    v = v_feat.transpose(1, 2) @ sd["projection.visual_proj.weight"].t() + sd["projection.visual_proj.bias"]
    a = a_feat.transpose(1, 2) @ sd["projection.audio_proj.weight"].t() + sd["projection.audio_proj.bias"]
    return v, a


def cross_modal(sd: SD, v_emb: Tensor, a_emb: Tensor) -> Tensor:
    # fusion_module.py:54-87
    if v_emb.dim() != 3 or a_emb.dim() != 3:
        raise ValueError("CrossModalAttention expects visual_emb and audio_emb of shape (B, T, D_e)")
    if v_emb.shape[0] != a_emb.shape[0] or v_emb.shape[2] != a_emb.shape[2]:
        raise ValueError("visual_emb and audio_emb must have the same batch size and feature dim")
    p = "cross_modal"
    a_emb = lerp_tokens(a_emb, v_emb.shape[1])
    v_out = v_emb + _mha(sd, p + ".v2a_attn", v_emb, a_emb)
    a_out = a_emb + _mha(sd, p + ".a2v_attn", a_emb, v_emb)
    gi = torch.cat([v_out, a_out], dim=-1)
    h = _gelu(gi @ sd[p + ".gate.0.weight"].t() + sd[p + ".gate.0.bias"])
    g = torch.sigmoid(h @ sd[p + ".gate.2.weight"].t() + sd[p + ".gate.2.bias"])
    fused = g * v_out + (1.0 - g) * a_out
    return F.relu(fused @ sd[p + ".fuse.0.weight"].t() + sd[p + ".fuse.0.bias"])


# ---------------------------------------------------------------- temporal transformer (temporal.py:79-111)
def temporal(sd: SD, x: Tensor, inter: Optional[Dict[str, Tensor]] = None) -> Tensor:
    if x.dim() != 3:
        raise ValueError(f"TemporalTransformer expected input of shape (B, T, D), got {tuple(x.shape)}")
    p = "temporal"
    xt = x.transpose(1, 2)
    branches = []
    for k in (3, 5, 7):
        c = F.conv1d(xt, sd[f"{p}.branch_k{k}.0.weight"], padding=k // 2)
        branches.append(_gelu(_bn(sd, f"{p}.branch_k{k}.1", c)))
    xc = torch.cat(branches, dim=1).transpose(1, 2)
    xc = xc @ sd[p + ".pre_scale_proj.weight"].t() + sd[p + ".pre_scale_proj.bias"]
    x = x + xc
    if inter is not None:
        inter["t_preconv"] = x
    cls = sd[p + ".cls_token"].expand(x.shape[0], -1, -1)
    tok = torch.cat([cls, x], dim=1)
    for l in range(4):
        lp = f"{p}.transformer.layers.{l}"
        h = _layer_norm(tok, sd[lp + ".norm1.weight"], sd[lp + ".norm1.bias"])
        tok = tok + _mha(sd, lp + ".self_attn", h, h)
        h = _layer_norm(tok, sd[lp + ".norm2.weight"], sd[lp + ".norm2.bias"])
        h = _gelu(h @ sd[lp + ".linear1.weight"].t() + sd[lp + ".linear1.bias"])
        tok = tok + (h @ sd[lp + ".linear2.weight"].t() + sd[lp + ".linear2.bias"])
        if inter is not None:
            inter[f"t_layer{l}"] = tok
    return tok[:, 0]  # no final norm (temporal.py:110-111)


# ---------------------------------------------------------------- artifact detector (artifact_detector.py)
def _temporal_detector(sd: SD, x: Tensor) -> Tensor:
    p = "artifact_detector.temporal_detector.temporal_conv"
    out = F.relu(_bn(sd, p + ".1", F.conv3d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1)))
    out = F.relu(_bn(sd, p + ".4", F.conv3d(out, sd[p + ".3.weight"], sd[p + ".3.bias"], padding=1)))
    return out.mean(dim=(2, 3, 4))


def _high_freq(sd: SD, video: Tensor, inter: Optional[Dict[str, Tensor]] = None) -> Tensor:
    p = "artifact_detector.high_freq_detector"
    b, c, t, h, w = video.shape
    x = video.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    x = F.conv2d(x, sd[p + ".laplacian.weight"], padding=1)  # learnable parameter (:33-35)
    x = x.reshape(b, t, 3, h, w).permute(0, 2, 1, 3, 4)
    x = F.relu(_bn(sd, p + ".conv3d.1", F.conv3d(x, sd[p + ".conv3d.0.weight"], sd[p + ".conv3d.0.bias"], stride=(1, 2, 2), padding=1)))
    if inter is not None:
        inter["hf_front"] = x
    x = F.relu(_bn(sd, p + ".conv3d.4", F.conv3d(x, sd[p + ".conv3d.3.weight"], sd[p + ".conv3d.3.bias"], stride=(1, 2, 2), padding=1)))
    return x.mean(dim=(2, 3, 4))


def artifact_detector(sd: SD, fmap: Tensor, cls: Tensor, video: Tensor, inter: Optional[Dict[str, Tensor]] = None) -> Tensor:
    raw = _temporal_detector(sd, fmap)
    if fmap.shape[2] > 1:
        delta_map = fmap[:, :, 1:] - fmap[:, :, :-1]
    else:
        delta_map = torch.zeros_like(fmap)  # :168-171
    delta = _temporal_detector(sd, delta_map)
    hf = _high_freq(sd, video, inter)
    if inter is not None:
        inter["art_raw"], inter["art_delta"], inter["art_hf"] = raw, delta, hf
    comb = torch.cat([cls, raw, delta, hf], dim=-1)
    p = "artifact_detector.artifact_fusion"
    h = F.relu(comb @ sd[p + ".0.weight"].t() + sd[p + ".0.bias"])
    return F.relu(h @ sd[p + ".2.weight"].t() + sd[p + ".2.bias"])


def classifier(sd: SD, x: Tensor) -> Tensor:
    if x.dim() != 2:
        raise ValueError(f"ClassificationHead expected input of shape (B, D), got {tuple(x.shape)}")
    p = "classifier.net"
    h = _gelu(x @ sd[p + ".0.weight"].t() + sd[p + ".0.bias"])
    h = _layer_norm(h, sd[p + ".3.weight"], sd[p + ".3.bias"])
    return (h @ sd[p + ".4.weight"].t() + sd[p + ".4.bias"]).squeeze(-1)


# ---------------------------------------------------------------- whole model (lip_sync_model.py:86-136)
@torch.no_grad()
def forward(sd: SD, visual: Tensor, audio: Tensor, return_aux: bool = False, inter: Optional[Dict[str, Tensor]] = None):
    visual = visual.float()
    audio = audio.float()
    v_feat, v_map = visual_encoder(sd, visual, inter)
    a_feat = audio_encoder(sd, audio, inter)
    v_emb, a_emb = projection(sd, v_feat, a_feat)
    fused = cross_modal(sd, v_emb, a_emb)
    cls = temporal(sd, fused, inter)
    art = artifact_detector(sd, v_map, cls, visual, inter)
    logits = classifier(sd, torch.cat([cls, art], dim=-1))
    if inter is not None:
        inter.update({"v_feat": v_feat, "a_feat": a_feat, "v_emb": v_emb, "a_emb": a_emb, "fused": fused,
                      "cls": cls, "artifact": art, "logits": logits})
    if not return_aux:
        return logits
    return logits, {"visual_tokens": v_emb, "audio_tokens": a_emb, "fused_tokens": fused, "cls_output": cls}


FLOP_PER_WINDOW = 31.29e9  # SURVEY.md §2.3 (2*MACs, canonical window)
