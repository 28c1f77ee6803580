// Shared forward-pass plumbing of the C-ABI translation units: shape algebra, workspace plan, fp32 launch helpers and the
// fp32 forward split into the parts the bf16 path reuses (audio encoder, token path, fusion MLP + head).
#pragma once
#include "../../include/lsd_b200.h"
#include "lsd_internal.h"
#include "lsd_kernels.h"

#include <string>
#include <vector>

namespace lsdfw {
using namespace lsd;


inline int osz(int i, int k, int s, int p) { return (i + 2 * p - k) / s + 1; }

struct Shapes {
  int B, T, H, W, F, Ta;
  int Hs, Ws, H1, W1, H2, W2, H3, W3, H4, W4;          // visual: stem conv, pool(=layer1), layer2..4
  int Fs, As, F1, A1, F2, A2, F3, A3, F4, A4;          // audio: stem conv, pool(=layer1), layer2..4 (A = time)
  int Hh, Wh, Hg, Wg;                                  // hf: conv0 out, conv3 out
  int Td;                                              // delta frames (T-1, or 1 with zeros if T == 1)
};

inline int make_shapes(lsd_handle* h, int B, int T, int H, int W, int F, int Ta, Shapes& s) {
  if (B < 0 || T < 1 || H < 1 || W < 1 || F < 1 || Ta < 1)
    return lsd_fail(h, LSD_ERR_SHAPE, "expected visual (B,3,T,H,W) and audio (B,1,F,T_a) with positive extents, got B=%d T=%d H=%d W=%d F=%d Ta=%d", B, T, H, W, F, Ta);
  s.B = B; s.T = T; s.H = H; s.W = W; s.F = F; s.Ta = Ta;
  s.Hs = osz(H, 7, 2, 3); s.Ws = osz(W, 7, 2, 3);
  s.H1 = osz(s.Hs, 3, 2, 1); s.W1 = osz(s.Ws, 3, 2, 1);
  s.H2 = osz(s.H1, 3, 2, 1); s.W2 = osz(s.W1, 3, 2, 1);
  s.H3 = osz(s.H2, 3, 2, 1); s.W3 = osz(s.W2, 3, 2, 1);
  s.H4 = osz(s.H3, 3, 2, 1); s.W4 = osz(s.W3, 3, 2, 1);
  s.Fs = osz(F, 7, 2, 3); s.As = osz(Ta, 7, 2, 3);
  s.F1 = osz(s.Fs, 3, 2, 1); s.A1 = osz(s.As, 3, 2, 1);
  s.F2 = osz(s.F1, 3, 2, 1); s.A2 = osz(s.A1, 3, 2, 1);
  s.F3 = osz(s.F2, 3, 2, 1); s.A3 = s.A2;               // stride (2,1): preserve_audio_temporal (audio_encoder.py:144)
  s.F4 = osz(s.F3, 3, 2, 1); s.A4 = s.A3;
  s.Hh = osz(H, 3, 2, 1); s.Wh = osz(W, 3, 2, 1);
  s.Hg = osz(s.Hh, 3, 2, 1); s.Wg = osz(s.Wh, 3, 2, 1);
  s.Td = T > 1 ? T - 1 : 1;
  if (s.H4 < 1 || s.W4 < 1 || s.F4 < 1 || s.A4 < 1 || s.Hg < 1 || s.Wg < 1)
    return lsd_fail(h, LSD_ERR_SHAPE, "input extents too small for the encoder strides");
  return 0;
}

struct Plan {
  std::vector<Stage> stages;
  size_t cursor = 0;  // bytes
  size_t add(const char* name, int64_t numel, int dtype = LSD_F32) {
    const size_t esz = (dtype == LSD_F32) ? 4 : 2;
    cursor = (cursor + 255) & ~size_t(255);
    Stage st; st.name = name; st.offset = cursor; st.numel = numel; st.dtype = dtype;
    stages.push_back(st);
    cursor += (size_t)numel * esz;
    return st.offset;
  }
  size_t find(const char* name) const {
    for (const Stage& s : stages) if (s.name == name) return s.offset;
    return (size_t)-1;
  }
};

inline void make_plan_f32(const Shapes& s, Plan& p) {
  const int64_t B = s.B, T = s.T;
  p.add("vid", B * T * s.H * s.W * 3);
  p.add("aud", B * s.F * s.Ta);
  p.add("v_stem_conv", B * T * s.Hs * s.Ws * 64);
  p.add("v_stem", B * T * s.H1 * s.W1 * 64);
  p.add("v_l1a", B * T * s.H1 * s.W1 * 64);
  p.add("v_layer1", B * T * s.H1 * s.W1 * 64);
  p.add("v_l2a", B * T * s.H2 * s.W2 * 128);
  p.add("v_l2d", B * T * s.H2 * s.W2 * 128);
  p.add("v_layer2", B * T * s.H2 * s.W2 * 128);
  p.add("v_l3a", B * T * s.H3 * s.W3 * 256);
  p.add("v_l3d", B * T * s.H3 * s.W3 * 256);
  p.add("v_layer3", B * T * s.H3 * s.W3 * 256);
  p.add("v_l4a", B * T * s.H4 * s.W4 * 256);
  p.add("v_l4d", B * T * s.H4 * s.W4 * 256);
  p.add("v_layer4", B * T * s.H4 * s.W4 * 256);
  p.add("v_feat", B * T * 256);
  p.add("a_stem_conv", B * s.Fs * s.As * 64);
  p.add("a_stem", B * s.F1 * s.A1 * 64);
  p.add("a_l1a", B * s.F1 * s.A1 * 64);
  p.add("a_layer1", B * s.F1 * s.A1 * 64);
  p.add("a_l2a", B * s.F2 * s.A2 * 128);
  p.add("a_l2d", B * s.F2 * s.A2 * 128);
  p.add("a_layer2", B * s.F2 * s.A2 * 128);
  p.add("a_l3a", B * s.F3 * s.A3 * 256);
  p.add("a_l3d", B * s.F3 * s.A3 * 256);
  p.add("a_layer3", B * s.F3 * s.A3 * 256);
  p.add("a_l4a", B * s.F4 * s.A4 * 256);
  p.add("a_l4d", B * s.F4 * s.A4 * 256);
  p.add("a_layer4", B * s.F4 * s.A4 * 256);
  p.add("a_feat", B * s.A4 * 256);
  p.add("v_emb", B * T * 256);
  p.add("a_emb", B * s.A4 * 256);
  p.add("a_int", B * T * 256);
  p.add("proj_v", B * T * 768);
  p.add("proj_a", B * T * 768);
  p.add("att1", B * T * 256);
  p.add("att2", B * T * 256);
  p.add("gate_in", B * T * 512);
  p.add("gate_h", B * T * 256);
  p.add("blend", B * T * 256);
  p.add("fused", B * T * 256);
  p.add("ms_cat", B * T * 768);
  p.add("tok", B * (T + 1) * 256);
  p.add("tok_ln", B * (T + 1) * 256);
  p.add("tok_qkv", B * (T + 1) * 768);
  p.add("tok_att", B * (T + 1) * 256);
  p.add("tok_ff", B * (T + 1) * 1024);
  p.add("t_layer0", B * (T + 1) * 256);
  p.add("t_layer3", B * (T + 1) * 256);
  p.add("art_a", B * T * s.H4 * s.W4 * 128);
  p.add("art_b", B * T * s.H4 * s.W4 * 64);
  p.add("art_delta", B * s.Td * s.H4 * s.W4 * 256);
  p.add("hf_lap", B * T * s.H * s.W * 3);
  p.add("hf_front", B * T * s.Hh * s.Wh * 32);
  p.add("hf_back", B * T * s.Hg * s.Wg * 64);
  p.add("comb", B * 448);
  p.add("art_h", B * 256);
  p.add("feat", B * 384);
  p.add("head_h", B * 128);
}

struct Ctx {
  lsd_handle* h;
  char* ws;
  const Plan* plan;
  cudaStream_t st;
  float* buf(const char* name) const { return reinterpret_cast<float*>(ws + plan->find(name)); }
  const float* W(const ConvP& c) const { return h->warena + c.w_off; }
};

// one conv/linear launch on the fp32 path
inline void conv(const Ctx& c, const char* key, const float* x, int in_ld, int N, int Ti, int Hi, int Wi, int st, int sh, int sw,
          int pt, int ph, int pw, float* y, int out_ld, int act, const float* res = nullptr, int res_ld = 0,
          int grp = 0, int grp_stride = 0, int row_off = 0) {
  const ConvP& w = c.h->convs.at(key);
  ConvF32 p;
  p.x = x; p.w = c.h->warena + w.w_off;
  p.scale = w.has_scale ? c.h->warena + w.scale_off : nullptr;
  p.shift = c.h->warena + w.shift_off;
  p.res = res; p.y = y;
  p.N = N; p.Ti = Ti; p.Hi = Hi; p.Wi = Wi; p.Cin = w.Cin;
  p.kt = w.kt; p.kh = w.kh; p.kw = w.kw; p.st = st; p.sh = sh; p.sw = sw; p.pt = pt; p.ph = ph; p.pw = pw;
  p.To = osz(Ti, w.kt, st, pt); p.Ho = osz(Hi, w.kh, sh, ph); p.Wo = osz(Wi, w.kw, sw, pw); p.Cout = w.Cout;
  p.in_ld = in_ld; p.w_ld = w.Cout; p.out_ld = out_ld; p.res_ld = res_ld; p.act = act;
  p.grp = grp; p.grp_stride = grp_stride; p.row_off = row_off;
  c.h->prof.begin(c.st, 2.0 * (double)N * p.To * p.Ho * p.Wo * w.Cout * (double)(w.kt * w.kh * w.kw * w.Cin));
  launch_conv_f32(p, c.st);
  c.h->prof.end(c.st);
}
inline void linear(const Ctx& c, const char* key, const float* x, int in_ld, int rows, float* y, int out_ld, int act,
                   const float* res = nullptr, int res_ld = 0, int grp = 0, int grp_stride = 0, int row_off = 0) {
  conv(c, key, x, in_ld, 1, 1, 1, rows, 1, 1, 1, 0, 0, 0, y, out_ld, act, res, res_ld, grp, grp_stride, row_off);
}

// Residual stage (visual_encoder.py:81-87 / audio_encoder.py:82-89): conv1+BN+ReLU, conv2+BN, (+BN(ds(x)) | +x), ReLU.
inline void res_stage(const Ctx& c, const std::string& p, const float* x, int N, int Ti, int Hi, int Wi, int st, int sh, int sw,
               bool is3d, float* a, float* d, float* y, int Cout) {
  const int pt = is3d ? 1 : 0;
  const ConvP& w1 = c.h->convs.at(p + ".conv1");
  const int To = osz(Ti, w1.kt, st, pt), Ho = osz(Hi, 3, sh, 1), Wo = osz(Wi, 3, sw, 1);
  conv(c, (p + ".conv1").c_str(), x, w1.Cin, N, Ti, Hi, Wi, st, sh, sw, pt, 1, 1, a, Cout, ACT_RELU);
  const float* idt = x;
  if (c.h->convs.count(p + ".downsample")) {
    conv(c, (p + ".downsample").c_str(), x, w1.Cin, N, Ti, Hi, Wi, st, sh, sw, 0, 0, 0, d, Cout, ACT_NONE);
    idt = d;
  }
  conv(c, (p + ".conv2").c_str(), a, Cout, N, To, Ho, Wo, 1, 1, 1, pt, 1, 1, y, Cout, ACT_RELU, idt, Cout);
}

inline void fw_inputs(const Ctx& c, const Shapes& s, const void* video, int vdt, int vlayout, const void* audio, int adt) {
  const int B = s.B, T = s.T;
  if (vlayout == LSD_NCDHW) launch_video_to_ndhwc(video, vdt, c.buf("vid"), B, 3, T, s.H, s.W, c.st);
  else launch_cast_to_f32(video, vdt, c.buf("vid"), (int64_t)B * T * s.H * s.W * 3, vdt == LSD_U8 ? 255.0f : 1.0f, c.st);
  launch_cast_to_f32(audio, adt, c.buf("aud"), (int64_t)B * s.F * s.Ta, 1.0f, c.st);
}

inline void fw_visual_f32(const Ctx& c, const Shapes& s) {
  lsd_handle* h = c.h; cudaStream_t st = c.st; const int B = s.B, T = s.T; (void)h; (void)st; (void)B; (void)T;
  float* vid = c.buf("vid");
  // ---- visual encoder (visual_encoder.py:166-201)
  conv(c, "visual_encoder.stem", vid, 3, B, T, s.H, s.W, 1, 2, 2, 1, 3, 3, c.buf("v_stem_conv"), 64, ACT_RELU);
  launch_maxpool3x3s2(c.buf("v_stem_conv"), c.buf("v_stem"), B * T, s.Hs, s.Ws, 64, st);
  res_stage(c, "visual_encoder.layer1", c.buf("v_stem"), B, T, s.H1, s.W1, 1, 1, 1, true, c.buf("v_l1a"), nullptr, c.buf("v_layer1"), 64);
  res_stage(c, "visual_encoder.layer2", c.buf("v_layer1"), B, T, s.H1, s.W1, 1, 2, 2, true, c.buf("v_l2a"), c.buf("v_l2d"), c.buf("v_layer2"), 128);
  res_stage(c, "visual_encoder.layer3", c.buf("v_layer2"), B, T, s.H2, s.W2, 1, 2, 2, true, c.buf("v_l3a"), c.buf("v_l3d"), c.buf("v_layer3"), 256);
  res_stage(c, "visual_encoder.layer4", c.buf("v_layer3"), B, T, s.H3, s.W3, 1, 2, 2, true, c.buf("v_l4a"), c.buf("v_l4d"), c.buf("v_layer4"), 256);
  launch_mean_mid(c.buf("v_layer4"), c.buf("v_feat"), B * T, s.H4 * s.W4, 256, 256, st);  // (B,T,256) token layout
}

inline void fw_audio_f32(const Ctx& c, const Shapes& s) {
  lsd_handle* h = c.h; cudaStream_t st = c.st; const int B = s.B, T = s.T; (void)h; (void)st; (void)B; (void)T;
  float* aud = c.buf("aud");
  // ---- audio encoder (audio_encoder.py:173-205): (B,1,F,Ta) == channels-last (B,F,Ta,1)
  conv(c, "audio_encoder.stem", aud, 1, B, 1, s.F, s.Ta, 1, 2, 2, 0, 3, 3, c.buf("a_stem_conv"), 64, ACT_RELU);
  launch_maxpool3x3s2(c.buf("a_stem_conv"), c.buf("a_stem"), B, s.Fs, s.As, 64, st);
  res_stage(c, "audio_encoder.layer1", c.buf("a_stem"), B, 1, s.F1, s.A1, 1, 1, 1, false, c.buf("a_l1a"), nullptr, c.buf("a_layer1"), 64);
  res_stage(c, "audio_encoder.layer2", c.buf("a_layer1"), B, 1, s.F1, s.A1, 1, 2, 2, false, c.buf("a_l2a"), c.buf("a_l2d"), c.buf("a_layer2"), 128);
  res_stage(c, "audio_encoder.layer3", c.buf("a_layer2"), B, 1, s.F2, s.A2, 1, 2, 1, false, c.buf("a_l3a"), c.buf("a_l3d"), c.buf("a_layer3"), 256);
  res_stage(c, "audio_encoder.layer4", c.buf("a_layer3"), B, 1, s.F3, s.A3, 1, 2, 1, false, c.buf("a_l4a"), c.buf("a_l4d"), c.buf("a_layer4"), 256);
  const int TA = s.A4;
  launch_mean_mid(c.buf("a_layer4"), c.buf("a_feat"), B, s.F4, TA * 256, TA * 256, st);  // mean over F' -> (B,TA,256)
}

// projection -> cross-modal -> temporal transformer; leaves cls in comb[:, :256] and feat[:, :256]
inline void fw_tokens_f32(const Ctx& c, const Shapes& s) {
  lsd_handle* h = c.h; cudaStream_t st = c.st; const int B = s.B, T = s.T; (void)h; (void)st; (void)B; (void)T;
  const int TA = s.A4;
  // ---- projection (fusion_module.py:108-124)
  linear(c, "projection.visual_proj", c.buf("v_feat"), 256, B * T, c.buf("v_emb"), 256, ACT_NONE);
  linear(c, "projection.audio_proj", c.buf("a_feat"), 256, B * TA, c.buf("a_emb"), 256, ACT_NONE);
  // ---- cross-modal attention + gated fusion (fusion_module.py:54-87)
  const float* a_int = c.buf("a_emb");
  if (TA != T) { launch_lerp_tokens(c.buf("a_emb"), c.buf("a_int"), B, TA, T, 256, st); a_int = c.buf("a_int"); }
  float *pv = c.buf("proj_v"), *pa = c.buf("proj_a"), *gi = c.buf("gate_in");
  linear(c, "cross.in_v", c.buf("v_emb"), 256, B * T, pv, 768, ACT_NONE);
  linear(c, "cross.in_a", a_int, 256, B * T, pa, 768, ACT_NONE);
  launch_mha_core(pv, 768, pa + 256, 768, pa + 512, 768, c.buf("att1"), 256, B, T, T, 8, st);  // v2a: Q=v, K/V=a
  launch_mha_core(pa, 768, pv + 256, 768, pv + 512, 768, c.buf("att2"), 256, B, T, T, 8, st);  // a2v: Q=a, K/V=v
  linear(c, "cross.v2a.out", c.buf("att1"), 256, B * T, gi, 512, ACT_NONE, c.buf("v_emb"), 256);       // v_out -> gate_in[:, :256]
  linear(c, "cross.a2v.out", c.buf("att2"), 256, B * T, gi + 256, 512, ACT_NONE, a_int, 256);          // a_out -> gate_in[:, 256:]
  linear(c, "cross.gate0", gi, 512, B * T, c.buf("gate_h"), 256, ACT_GELU);
  launch_gate_blend(c.buf("gate_h"), h->warena + h->vecs.at("cross.gate2.w"), h->warena + h->vecs.at("cross.gate2.b"), gi, 512,
                    gi + 256, 512, c.buf("blend"), B * T, 256, st);
  linear(c, "cross.fuse", c.buf("blend"), 256, B * T, c.buf("fused"), 256, ACT_RELU);
  // ---- temporal transformer (temporal.py:79-111)
  const float* fused = c.buf("fused");
  for (int k : {3, 5, 7}) {
    const std::string key = "temporal.branch_k" + std::to_string(k);
    conv(c, key.c_str(), fused, 256, B, 1, 1, T, 1, 1, 1, 0, 0, k / 2, c.buf("ms_cat") + (k / 2 - 1) * 256, 768, ACT_GELU);
  }
  const int NT = T + 1;
  float* tok = c.buf("tok");
  launch_set_cls(h->warena + h->vecs.at("temporal.cls"), tok, B, NT, 256, st);
  // pre_scale_proj + residual, written straight into token rows 1..T of each window
  linear(c, "temporal.pre_scale_proj", c.buf("ms_cat"), 768, B * T, tok, 256, ACT_NONE, fused, 256, T, NT, 1);
  for (int l = 0; l < 4; ++l) {
    const std::string k = "t" + std::to_string(l);
    launch_layernorm(tok, 256, h->warena + h->vecs.at(k + ".ln1.w"), h->warena + h->vecs.at(k + ".ln1.b"), c.buf("tok_ln"), 256, B * NT, 256, st);
    linear(c, (k + ".in").c_str(), c.buf("tok_ln"), 256, B * NT, c.buf("tok_qkv"), 768, ACT_NONE);
    const float* qkv = c.buf("tok_qkv");
    launch_mha_core(qkv, 768, qkv + 256, 768, qkv + 512, 768, c.buf("tok_att"), 256, B, NT, NT, 8, st);
    linear(c, (k + ".out").c_str(), c.buf("tok_att"), 256, B * NT, tok, 256, ACT_NONE, tok, 256);
    launch_layernorm(tok, 256, h->warena + h->vecs.at(k + ".ln2.w"), h->warena + h->vecs.at(k + ".ln2.b"), c.buf("tok_ln"), 256, B * NT, 256, st);
    linear(c, (k + ".ff1").c_str(), c.buf("tok_ln"), 256, B * NT, c.buf("tok_ff"), 1024, ACT_GELU);
    linear(c, (k + ".ff2").c_str(), c.buf("tok_ff"), 1024, B * NT, tok, 256, ACT_NONE, tok, 256);
    if (l == 0) launch_copy_rows(tok, 256, c.buf("t_layer0"), 256, B * NT, 256, st);
    if (l == 3) launch_copy_rows(tok, 256, c.buf("t_layer3"), 256, B * NT, 256, st);
  }
  // cls = tok[:,0]: no final norm (temporal.py:110-111)
  float *comb = c.buf("comb"), *feat = c.buf("feat");
  launch_copy_rows(tok, (int64_t)NT * 256, comb, 448, B, 256, st);
  launch_copy_rows(tok, (int64_t)NT * 256, feat, 384, B, 256, st);
}

inline void fw_artifact_f32(const Ctx& c, const Shapes& s) {
  lsd_handle* h = c.h; cudaStream_t st = c.st; const int B = s.B, T = s.T; (void)h; (void)st; (void)B; (void)T;
  float* vid = c.buf("vid");
  const float* vmap = c.buf("v_layer4");
  float* comb = c.buf("comb");
  // ---- artifact detector (artifact_detector.py:149-183)
  conv(c, "art.td0", vmap, 256, B, T, s.H4, s.W4, 1, 1, 1, 1, 1, 1, c.buf("art_a"), 128, ACT_RELU);
  conv(c, "art.td3", c.buf("art_a"), 128, B, T, s.H4, s.W4, 1, 1, 1, 1, 1, 1, c.buf("art_b"), 64, ACT_RELU);
  launch_mean_mid(c.buf("art_b"), comb + 256, B, T * s.H4 * s.W4, 64, 448, st);
  if (T > 1) launch_delta_t(vmap, c.buf("art_delta"), B, T, (int64_t)s.H4 * s.W4 * 256, st);
  else launch_fill_zero(c.buf("art_delta"), (int64_t)B * s.H4 * s.W4 * 256, st);  // zeros_like (artifact_detector.py:168-171)
  conv(c, "art.td0", c.buf("art_delta"), 256, B, s.Td, s.H4, s.W4, 1, 1, 1, 1, 1, 1, c.buf("art_a"), 128, ACT_RELU);
  conv(c, "art.td3", c.buf("art_a"), 128, B, s.Td, s.H4, s.W4, 1, 1, 1, 1, 1, 1, c.buf("art_b"), 64, ACT_RELU);
  launch_mean_mid(c.buf("art_b"), comb + 320, B, s.Td * s.H4 * s.W4, 64, 448, st);
  conv(c, "art.lap", vid, 3, B * T, 1, s.H, s.W, 1, 1, 1, 0, 1, 1, c.buf("hf_lap"), 3, ACT_NONE);   // per-frame 3->3 (parameter)
  conv(c, "art.hf0", c.buf("hf_lap"), 3, B, T, s.H, s.W, 1, 2, 2, 1, 1, 1, c.buf("hf_front"), 32, ACT_RELU);
  conv(c, "art.hf3", c.buf("hf_front"), 32, B, T, s.Hh, s.Wh, 1, 2, 2, 1, 1, 1, c.buf("hf_back"), 64, ACT_RELU);
  launch_mean_mid(c.buf("hf_back"), comb + 384, B, T * s.Hg * s.Wg, 64, 448, st);
}

// artifact fusion MLP + classification head (+ optional aux outputs)
inline void fw_head_f32(const Ctx& c, const Shapes& s, float* logits, const lsd_aux* aux) {
  lsd_handle* h = c.h; cudaStream_t st = c.st; const int B = s.B, T = s.T; (void)h; (void)st; (void)B; (void)T;
  const int TA = s.A4, NT = T + 1;
  float *comb = c.buf("comb"), *feat = c.buf("feat"), *tok = c.buf("tok");
  linear(c, "art.fuse0", comb, 448, B, c.buf("art_h"), 256, ACT_RELU);
  linear(c, "art.fuse2", c.buf("art_h"), 256, B, feat + 256, 384, ACT_RELU);
  // ---- head (classifier.py:22-34)
  linear(c, "head.fc0", feat, 384, B, c.buf("head_h"), 128, ACT_GELU);
  launch_ln_dot(c.buf("head_h"), h->warena + h->vecs.at("head.ln.w"), h->warena + h->vecs.at("head.ln.b"),
                h->warena + h->vecs.at("head.out.w"), h->warena + h->vecs.at("head.out.b"), logits, B, 128, st);
  if (aux) {
    const size_t tb = (size_t)B * T * 256 * sizeof(float);
    if (aux->visual_tokens) cudaMemcpyAsync(aux->visual_tokens, c.buf("v_emb"), tb, cudaMemcpyDeviceToDevice, st);
    if (aux->audio_tokens) cudaMemcpyAsync(aux->audio_tokens, c.buf("a_emb"), (size_t)B * TA * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (aux->fused_tokens) cudaMemcpyAsync(aux->fused_tokens, c.buf("fused"), tb, cudaMemcpyDeviceToDevice, st);
    if (aux->cls_output) launch_copy_rows(tok, (int64_t)NT * 256, aux->cls_output, 256, B, 256, st);
  }
}

inline int forward_f32(lsd_handle* h, const Shapes& s, const Plan& plan, char* ws, const lsd_aux* aux, float* logits, cudaStream_t st,
                       bool inputs_ready, const void* video, int vdt, int vlayout, const void* audio, int adt) {
  Ctx c{h, ws, &plan, st};
  if (!inputs_ready) fw_inputs(c, s, video, vdt, vlayout, audio, adt);
  fw_visual_f32(c, s);
  fw_audio_f32(c, s);
  fw_tokens_f32(c, s);
  fw_artifact_f32(c, s);
  fw_head_f32(c, s, logits, aux);
  return 0;
}

}  // namespace lsdfw
