// bf16 tensor-core path: weight packing for the tcgen05 flat shift-GEMM kernel (umma_conv.cu) and the forward
// orchestration that runs the 3-D residual stages, the artifact-detector convolutions and the high-frequency
// back end on it.  Host-side only.
#include "forward_common.h"
#include "umma_conv.cuh"

#include <algorithm>
#include <cstring>

using namespace lsd;
using namespace lsdfw;

namespace {

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// Tap/band structure of a k_t x k_h x k_w convolution with stride (1, sh, sw) and "same" padding k/2 over the padded
// planar layout: taps that read the same parity plane set and the same temporal offset form one band.
BGroup make_group(const ConvP& c, int sh, int sw) {
  BGroup g;
  g.Cin = c.Cin;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b)
      for (int d = 0; d < c.kw; ++d) {
        BTap t;
        t.orig = (a * c.kh + b) * c.kw + d;
        t.dt = a - c.kt / 2;
        int hp = 0, wp = 0;
        const int oh = b - c.kh / 2, ow = d - c.kw / 2;
        if (sh == 2) { hp = oh & 1; t.dh = (oh - hp) / 2; } else t.dh = oh;
        if (sw == 2) { wp = ow & 1; t.dw = (ow - wp) / 2; } else t.dw = ow;
        const int set = hp * 2 + wp;
        BBand* band = nullptr;
        for (BBand& bb : g.bands)
          if (bb.set == set && bb.taps[0].dt == t.dt && (int)bb.taps.size() < UC_MAX_TAPS) band = &bb;
        if (!band) { g.bands.push_back(BBand{set, {}}); band = &g.bands.back(); }
        band->taps.push_back(t);
      }
  g.taps_total = c.kt * c.kh * c.kw;
  g.k16 = c.Cin / 16;
  return g;
}

// Toeplitz group (Cin = 3 convolutions with stride 2 in w, reading bf16 "pixel rows": 4 channels x 2 B per pixel, h-parity
// split, 2 pixels per 16-byte unit): a "tap" is one (kt, kh) pair, K runs over `kpix` consecutive pixels of the input row
// starting at pixel 2*wo + kw0 - pad_w (kw0 = -1: one leading zero-weight pixel keeps the start 16-byte aligned).
BGroup make_group_toeplitz(const ConvP& c, int kpix, int dw_units) {
  BGroup g;
  g.Cin = c.Cin;
  g.toeplitz = 1;
  g.k16 = kpix * 4 / 16;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b) {
      BTap t;
      t.orig = a * c.kh + b;
      t.dt = a - c.kt / 2;
      const int oh = b - c.kh / 2;
      const int hp = oh & 1;
      t.dh = (oh - hp) / 2;
      t.dw = dw_units;
      BBand* band = nullptr;
      for (BBand& bb : g.bands)
        if (bb.set == hp && bb.taps[0].dt == t.dt) band = &bb;
      if (!band) { g.bands.push_back(BBand{hp, {}}); band = &g.bands.back(); }
      band->taps.push_back(t);
    }
  g.taps_total = c.kt * c.kh;
  return g;
}

struct Packer {
  lsd_handle* h;
  const std::vector<float>& f32;
  std::vector<uint16_t> w;
  std::vector<float> bias;
  // packed order: [k16 chunk][tap in band order][2 k-chunks][Cout][8]
  void pack_group(BGroup& g, const ConvP& c) {
    g.w_off = w.size();
    const int Cout = c.Cout;
    const float* W = &f32[c.w_off];
    const float* sc = c.has_scale ? &f32[c.scale_off] : nullptr;
    for (int ch = 0; ch < c.Cin / 16; ++ch)
      for (const BBand& b : g.bands)
        for (const BTap& t : b.taps)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < Cout; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = ch * 16 + kc * 8 + e;
                const float v = W[((size_t)t.orig * c.Cin + ci) * Cout + n] * (sc ? sc[n] : 1.0f);
                w.push_back(f2bf(v));
              }
  }
  // Toeplitz packing: K index k = c*16 + kc*8 + e  <->  pixel j = k/4 of the row window, channel k%4; kw = j - 1.
  void add_toeplitz(const std::string& name, const std::string& key, int kpix, int dw_units) {
    const ConvP& c = h->convs.at(key);
    BLayer L;
    L.Cout = c.Cout;
    L.groups.push_back(make_group_toeplitz(c, kpix, dw_units));
    BGroup& g = L.groups[0];
    g.w_off = w.size();
    const float* W = &f32[c.w_off];
    const float* sc = c.has_scale ? &f32[c.scale_off] : nullptr;
    for (int ch = 0; ch < g.k16; ++ch)
      for (const BBand& b : g.bands)
        for (const BTap& t : b.taps)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < c.Cout; ++n)
              for (int e = 0; e < 8; ++e) {
                const int k = ch * 16 + kc * 8 + e, j = k / 4, ci = k % 4, kw = j - 1;
                float v = 0.f;
                if (ci < c.Cin && kw >= 0 && kw < c.kw) v = W[((size_t)(t.orig * c.kw + kw) * c.Cin + ci) * c.Cout + n] * (sc ? sc[n] : 1.0f);
                w.push_back(f2bf(v));
              }
    L.bias_off = bias.size();
    for (int i = 0; i < c.Cout; ++i) bias.push_back(f32[c.shift_off + i]);
    while (w.size() % 64) w.push_back(0);
    h->blayers[name] = L;
  }
  void add(const std::string& name, const std::string& key, int sh, int sw, const std::string& ds_key = "") {
    const ConvP& c = h->convs.at(key);
    BLayer L;
    L.Cout = c.Cout;
    L.groups.push_back(make_group(c, sh, sw));
    pack_group(L.groups[0], c);
    L.bias_off = bias.size();
    for (int i = 0; i < c.Cout; ++i) bias.push_back(f32[c.shift_off + i]);
    if (!ds_key.empty()) {
      const ConvP& d = h->convs.at(ds_key);
      L.groups.push_back(make_group(d, 2, 2));  // 1x1x1 stride (1,2,2) reads parity set (0,0) with zero shift
      pack_group(L.groups[1], d);
      for (int i = 0; i < c.Cout; ++i) bias[L.bias_off + i] += f32[d.shift_off + i];
    }
    while (w.size() % 64) w.push_back(0);
    h->blayers[name] = L;
  }
};

// ---- planar activation buffers ----------------------------------------------------------------
struct PBuf {
  size_t off = 0;          // byte offset into the workspace
  int C = 0, sets = 1;
  UcGeom g;                // geometry of each plane set
  int64_t plane_stride = 0, set_stride = 0, origin = 0;  // elements
  size_t bytes() const { return (size_t)sets * set_stride * 2; }
};

constexpr int TILE_MAX = 1024;  // largest MT*128

PBuf make_pbuf(int C, int sets, UcGeom g) {
  PBuf b;
  b.C = C; b.sets = sets; b.g = g;
  const int64_t gf = ((int64_t)g.SL + 3 * g.RW + 16 + 7) / 8 * 8;
  const int64_t gb = gf + TILE_MAX;
  const int64_t plane_pos = (gf + g.P_total + gb + 7) / 8 * 8;
  b.origin = gf * 8;
  b.plane_stride = plane_pos * 8;
  b.set_stride = b.plane_stride * (C / 8);
  return b;
}

struct BPlan {
  Shapes s;
  Plan f32;                               // fp32 buffers shared with the fp32 path's tail (same names)
  std::map<std::string, PBuf> pb;
  size_t planar_begin = 0, planar_end = 0;
  void addp(const char* name, int C, int sets, UcGeom g) {
    PBuf b = make_pbuf(C, sets, g);
    f32.cursor = (f32.cursor + 255) & ~size_t(255);
    b.off = f32.cursor;
    f32.cursor += b.bytes();
    pb[name] = b;
  }
};

void build_plan(const Shapes& s, BPlan& P) {
  P.s = s;
  Plan& p = P.f32;
  const int64_t B = s.B, T = s.T;
  // fp32 buffers (names shared with make_plan_f32 so the fw_* helpers work on either plan)
  p.add("vid", B * T * s.H * s.W * 3);
  p.add("aud", B * s.F * s.Ta);
  p.add("v_feat", B * T * 256);
  for (const char* n : {"a_stem_conv"}) p.add(n, B * s.Fs * s.As * 64);
  for (const char* n : {"a_stem", "a_l1a", "a_layer1"}) p.add(n, B * s.F1 * s.A1 * 64);
  for (const char* n : {"a_l2a", "a_l2d", "a_layer2"}) p.add(n, B * s.F2 * s.A2 * 128);
  for (const char* n : {"a_l3a", "a_l3d", "a_layer3"}) p.add(n, B * s.F3 * s.A3 * 256);
  for (const char* n : {"a_l4a", "a_l4d", "a_layer4"}) p.add(n, B * s.F4 * s.A4 * 256);
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
  for (const char* n : {"tok", "tok_ln", "tok_att", "t_layer0", "t_layer3"}) p.add(n, B * (T + 1) * 256);
  p.add("tok_qkv", B * (T + 1) * 768);
  p.add("tok_ff", B * (T + 1) * 1024);
  p.add("comb", B * 448);
  p.add("art_h", B * 256);
  p.add("feat", B * 384);
  p.add("head_h", B * 128);
  // planar bf16 buffers
  P.planar_begin = (p.cursor + 255) & ~size_t(255);
  const int Bn = s.B, Tn = s.T;
  const UcGeom g1 = make_geom(Bn, Tn, s.H1, s.W1), g2 = make_geom(Bn, Tn, s.H2, s.W2), g3 = make_geom(Bn, Tn, s.H3, s.W3),
               g4 = make_geom(Bn, Tn, s.H4, s.W4), gd = make_geom(Bn, s.Td, s.H4, s.W4), gh = make_geom(Bn, Tn, s.Hg, s.Wg);
  // bf16 pixel rows of the video and of its per-frame 3->3 "laplacian" conv (h-parity split; 16-byte unit = 2 pixels x 4 ch)
  const UcGeom gs = make_geom_ex(Bn, Tn, s.Hs, s.Ws, 1, 2, 0, 0, 4);
  P.addp("xs", 8, 2, gs);
  P.addp("xl", 8, 2, gs);
  P.addp("s_out", 64, 1, gs);    // stem conv output (stem geometry), before the max-pool
  P.addp("x1", 64, 1, g1);
  P.addp("l1a", 64, 1, g1);
  P.addp("y1", 64, 4, g2);       // parity-split: four half-resolution plane sets
  P.addp("l2a", 128, 1, g2);
  P.addp("y2", 128, 4, g3);
  P.addp("l3a", 256, 1, g3);
  P.addp("y3", 256, 4, g4);
  P.addp("l4a", 256, 1, g4);
  P.addp("y4", 256, 1, g4);
  P.addp("art_a", 128, 1, g4);
  P.addp("art_b", 64, 1, g4);
  P.addp("delta", 256, 1, gd);
  P.addp("artd_a", 128, 1, gd);
  P.addp("artd_b", 64, 1, gd);
  P.addp("hf_f", 32, 4, gh);
  P.addp("hf_b", 64, 1, gh);
  P.planar_end = p.cursor;
}

struct BCtx {
  lsd_handle* h;
  char* ws;
  const BPlan* P;
  cudaStream_t st;
  __nv_bfloat16* base(const PBuf& b) const { return reinterpret_cast<__nv_bfloat16*>(ws + b.off); }
};

// One launch of the tcgen05 kernel: layer `name`, main input `in` (+ `in_ds` for the fused downsample group),
// output `out` (plain when out.sets == 1, parity-split when 4), optional residual (plain, output geometry).
int run_umma(const BCtx& c, const std::string& name, const PBuf& in, const PBuf* in_ds, const PBuf& out, UcGeom og, int act,
             const PBuf* res, float* y32 = nullptr, int y32_ld = 0) {
  const BLayer& L = c.h->blayers.at(name);
  UmmaConvP p;
  memset(&p, 0, sizeof(p));
  p.w = reinterpret_cast<const __nv_bfloat16*>(c.h->barena);
  p.bias = c.h->bbias + L.bias_off;
  p.Cout = L.Cout;
  p.act = act;
  p.g = og;
  if (y32) {
    p.out_mode = UC_OUT_F32_ROWS; p.y32 = y32; p.y32_ld = y32_ld;
  } else {
    p.out_mode = out.sets == 4 ? UC_OUT_PARITY : UC_OUT_PLAIN;
    p.y = c.base(out) + out.origin;
    p.y_plane_stride = out.plane_stride; p.y_set_stride = out.set_stride;
    p.g2 = out.g;
  }
  if (res) { p.res = c.base(*res) + res->origin; p.res_plane_stride = res->plane_stride; }
  // tile shape: all Cout columns x MT M-tiles in TMEM (512 columns)
  p.MT = L.Cout <= 64 ? 4 : 2;
  uint32_t cols = 32;
  while ((int)cols < p.MT * L.Cout) cols *= 2;
  p.tmem_cols = cols;
  const int S = p.MT * 128;
  int nb = 0, max_extra = 0, max_taps = 0;
  p.ngroups = (int)L.groups.size();
  for (int gi = 0; gi < p.ngroups; ++gi) {
    const BGroup& G = L.groups[gi];
    const PBuf& src = (gi == 0) ? in : *in_ds;
    UcGroup& ug = p.groups[gi];
    ug.band_begin = nb;
    ug.k16 = G.k16;
    ug.taps_total = G.taps_total;
    ug.w_off = (int64_t)G.w_off;
    int tap_begin = 0;
    for (const BBand& b : G.bands) {
      if (nb >= UC_MAX_BANDS) return lsd_fail(c.h, LSD_ERR_UNSUPPORTED, "%s: too many bands", name.c_str());
      UcBand& ub = p.bands[nb++];
      ub.base = c.base(src) + (int64_t)b.set * src.set_stride + src.origin;
      ub.plane_stride = src.plane_stride;
      ub.toeplitz = G.toeplitz;
      ub.chunk_stride = G.toeplitz ? 16 : 2 * src.plane_stride;
      int mn = INT32_MAX, mx = INT32_MIN;
      for (const BTap& t : b.taps) {
        const int sft = t.dt * og.SL + t.dh * og.RW + t.dw;
        mn = std::min(mn, sft); mx = std::max(mx, sft);
      }
      ub.start = mn;
      ub.len_extra = mx - mn;
      ub.ntaps = (int)b.taps.size();
      ub.tap_begin = tap_begin;
      for (int j = 0; j < ub.ntaps; ++j) ub.rel[j] = b.taps[j].dt * og.SL + b.taps[j].dh * og.RW + b.taps[j].dw - mn;
      tap_begin += ub.ntaps;
      max_extra = std::max(max_extra, ub.len_extra + ub.toeplitz);
      max_taps = std::max(max_taps, ub.ntaps);
      // the band must stay inside the guard zones of the source buffer
      if (-(int64_t)ub.start * 8 > src.origin) return lsd_fail(c.h, LSD_ERR_UNSUPPORTED, "%s: front guard too small", name.c_str());
    }
    ug.band_end = nb;
  }
  p.nbands = nb;
  p.a_stage_bytes = ((uint32_t)(2 * (S + max_extra) * 16) + 127u) & ~127u;
  p.w_stage_bytes = (uint32_t)(max_taps * L.Cout * 32);
  const uint32_t stage = p.a_stage_bytes + p.w_stage_bytes;
  const uint32_t budget = (cols <= 256 ? 110u : 218u) * 1024u;
  int stages = (int)(budget / stage);
  stages = std::max(2, std::min(stages, 6));
  if ((size_t)stages * stage + 1024 > 224u * 1024u) return lsd_fail(c.h, LSD_ERR_UNSUPPORTED, "%s: stage of %u bytes does not fit", name.c_str(), stage);
  p.stages = stages;
  double kflop = 0;
  for (const BGroup& G : L.groups) kflop += (double)G.taps_total * G.Cin;
  if (L.groups[0].toeplitz) { const ConvP& cp = c.h->convs.at(name); kflop = (double)cp.kt * cp.kh * cp.kw * cp.Cin; }  // algorithmic K, not the padded one
  c.h->prof.begin(c.st, 2.0 * (double)og.N * og.T * og.H * og.W * L.Cout * kflop, 2);
  launch_umma_conv(p, c.st);
  c.h->prof.end(c.st);
  return 0;
}

}  // namespace

int pack_bf16_weights(lsd_handle* h, const std::vector<float>& f32_arena) {
  Packer P{h, f32_arena, {}, {}};
  h->blayers.clear();
  for (int l = 1; l <= 4; ++l) {
    const std::string p = "visual_encoder.layer" + std::to_string(l);
    const int s = l == 1 ? 1 : 2;
    P.add(p + ".conv1", p + ".conv1", s, s);
    P.add(p + ".conv2", p + ".conv2", 1, 1, l == 1 ? "" : p + ".downsample");
  }
  P.add_toeplitz("visual_encoder.stem", "visual_encoder.stem", 8, 0);  // 7 taps in w -> 8-pixel window starting at 2*wo-4
  P.add_toeplitz("art.hf0", "art.hf0", 4, 1);                          // 3 taps in w -> 4-pixel window starting at 2*wo-2
  P.add("art.td0", "art.td0", 1, 1);
  P.add("art.td3", "art.td3", 1, 1);
  P.add("art.hf3", "art.hf3", 2, 2);
  if (h->barena) { cudaFree(h->barena); h->barena = nullptr; }
  if (h->bbias) { cudaFree(h->bbias); h->bbias = nullptr; }
  cudaError_t e = cudaMalloc(&h->barena, P.w.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(h->barena, P.w.data(), P.w.size() * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->bbias, P.bias.size() * 4);
  if (e == cudaSuccess) e = cudaMemcpy(h->bbias, P.bias.data(), P.bias.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "pack_bf16_weights: %s", cudaGetErrorString(e));
  return 0;
}

void make_plan_bf16(lsd_handle* h, int B, int T, int H, int W, int F, int Ta, std::vector<Stage>& stages, size_t& bytes) {
  Shapes s;
  if (make_shapes(h, B, T, H, W, F, Ta, s) != 0) { bytes = 0; return; }
  BPlan P;
  build_plan(s, P);
  stages = P.f32.stages;
  bytes = P.f32.cursor;
}

static int forward_bf16_impl(lsd_handle* h, const Shapes& s, float* logits, const lsd_aux* aux, char* ws, size_t ws_bytes,
                             cudaStream_t st, bool inputs_ready, const void* video, int vdt, int vlayout, const void* audio, int adt) {
  BPlan P;
  build_plan(s, P);
  if (P.f32.cursor > ws_bytes) return lsd_fail(h, LSD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", P.f32.cursor, ws_bytes);
  h->stages = P.f32.stages;
  // Zero padding of the planar buffers lives in the workspace across calls: (re)initialise when the workspace or the
  // shapes change.  Kernels only ever write valid positions (or zeros at pad positions), so the padding stays intact.
  const int sig[6] = {s.B, s.T, s.H, s.W, s.F, s.Ta};
  if (h->ws_sig_ptr != ws || h->ws_sig_bytes != ws_bytes || memcmp(h->ws_sig_shape, sig, sizeof(sig)) != 0) {
    cudaError_t e = cudaMemsetAsync(ws + P.planar_begin, 0, P.planar_end - P.planar_begin, st);
    if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "workspace init: %s", cudaGetErrorString(e));
    h->ws_sig_ptr = ws; h->ws_sig_bytes = ws_bytes; memcpy(h->ws_sig_shape, sig, sizeof(sig));
  }
  Ctx c{h, ws, &P.f32, st};
  BCtx b{h, ws, &P, st};
  const int B = s.B, T = s.T;
  if (!inputs_ready) launch_cast_to_f32(audio, adt, c.buf("aud"), (int64_t)B * s.F * s.Ta, 1.0f, st);
  auto& pb = P.pb;
  int rc = 0;
  // ---- video -> bf16 pixel rows (+ per-frame laplacian conv), stem conv on tcgen05 (Toeplitz K), max-pool in planar layout
  const PBuf &xs = pb["xs"], &xl = pb["xl"], &so = pb["s_out"], &x1 = pb["x1"];
  if (inputs_ready) launch_video_rows(c.buf("vid"), LSD_F32, LSD_NDHWC, h->warena + h->convs.at("art.lap").w_off, b.base(xs) + xs.origin,
                                      b.base(xl) + xl.origin, xs.set_stride, xs.g, s.H, s.W, st);
  else launch_video_rows(video, vdt, vlayout, h->warena + h->convs.at("art.lap").w_off, b.base(xs) + xs.origin, b.base(xl) + xl.origin,
                         xs.set_stride, xs.g, s.H, s.W, st);
  if ((rc = run_umma(b, "visual_encoder.stem", xs, nullptr, so, xs.g, ACT_RELU, nullptr))) return rc;
  launch_planar_maxpool(b.base(so) + so.origin, so.plane_stride, so.g, b.base(x1) + x1.origin, x1.plane_stride, x1.g, 64, st);
  // ---- residual stages on tcgen05 (visual_encoder.py:81-87, 133-152)
  if ((rc = run_umma(b, "visual_encoder.layer1.conv1", x1, nullptr, pb["l1a"], x1.g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer1.conv2", pb["l1a"], nullptr, pb["y1"], x1.g, ACT_RELU, &x1))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer2.conv1", pb["y1"], nullptr, pb["l2a"], pb["l2a"].g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer2.conv2", pb["l2a"], &pb["y1"], pb["y2"], pb["l2a"].g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer3.conv1", pb["y2"], nullptr, pb["l3a"], pb["l3a"].g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer3.conv2", pb["l3a"], &pb["y2"], pb["y3"], pb["l3a"].g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer4.conv1", pb["y3"], nullptr, pb["l4a"], pb["l4a"].g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "visual_encoder.layer4.conv2", pb["l4a"], &pb["y3"], pb["y4"], pb["l4a"].g, ACT_RELU, nullptr))) return rc;
  const PBuf& y4 = pb["y4"];
  launch_planar_mean(b.base(y4) + y4.origin, y4.plane_stride, y4.g, 256, c.buf("v_feat"), 256, 0, st);  // spatial mean -> tokens
  // ---- audio encoder + token path (fp32 kernels)
  fw_audio_f32(c, s);
  fw_tokens_f32(c, s);
  float* comb = c.buf("comb");
  // ---- artifact detector (artifact_detector.py:149-183)
  if ((rc = run_umma(b, "art.td0", y4, nullptr, pb["art_a"], y4.g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "art.td3", pb["art_a"], nullptr, pb["art_b"], y4.g, ACT_RELU, nullptr))) return rc;
  launch_planar_mean(b.base(pb["art_b"]) + pb["art_b"].origin, pb["art_b"].plane_stride, y4.g, 64, comb + 256, 448, 1, st);
  const PBuf& dl = pb["delta"];
  if (T > 1) launch_planar_delta(b.base(y4) + y4.origin, y4.plane_stride, y4.g, b.base(dl) + dl.origin, dl.plane_stride, dl.g, 256, st);
  // (T == 1: the delta map is all zeros — the buffer is never written and keeps its zero initialisation)
  if ((rc = run_umma(b, "art.td0", dl, nullptr, pb["artd_a"], dl.g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "art.td3", pb["artd_a"], nullptr, pb["artd_b"], dl.g, ACT_RELU, nullptr))) return rc;
  launch_planar_mean(b.base(pb["artd_b"]) + pb["artd_b"].origin, pb["artd_b"].plane_stride, dl.g, 64, comb + 320, 448, 1, st);
  // high-frequency branch: Conv3d 3->32 s(1,2,2) on the laplacian pixel rows (Toeplitz K), Conv3d 32->64 s(1,2,2) planar
  const PBuf& hf = pb["hf_f"];
  if ((rc = run_umma(b, "art.hf0", xl, nullptr, hf, xl.g, ACT_RELU, nullptr))) return rc;
  if ((rc = run_umma(b, "art.hf3", hf, nullptr, pb["hf_b"], hf.g, ACT_RELU, nullptr))) return rc;
  launch_planar_mean(b.base(pb["hf_b"]) + pb["hf_b"].origin, pb["hf_b"].plane_stride, hf.g, 64, comb + 384, 448, 1, st);
  // ---- fusion MLP + head (fp32)
  fw_head_f32(c, s, logits, aux);
  return 0;
}

int forward_bf16(lsd_handle* h, int B, int T, int H, int W, int F, int Ta, const void* video, int vdt, int vlayout,
                 const void* audio, int adt, float* logits, const lsd_aux* aux, char* ws, size_t ws_bytes, cudaStream_t st,
                 bool inputs_ready) {
  Shapes s;
  int rc = make_shapes(h, B, T, H, W, F, Ta, s);
  if (rc) return rc;
  return forward_bf16_impl(h, s, logits, aux, ws, ws_bytes, st, inputs_ready, video, vdt, vlayout, audio, adt);
}

int score_batch_bf16(lsd_handle* h, const uint8_t* track, int n_frames, const int32_t* d_vstarts, const int32_t* d_astarts,
                     const float* mel_full, int Ta_full, int nb, int T, int H, int W, int F, int Ta, float* logits, char* ws,
                     size_t ws_bytes, cudaStream_t st) {
  Shapes s;
  int rc = make_shapes(h, nb, T, H, W, F, Ta, s);
  if (rc) return rc;
  BPlan P;
  build_plan(s, P);
  if (P.f32.cursor > ws_bytes) return lsd_fail(h, LSD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", P.f32.cursor, ws_bytes);
  Ctx c{h, ws, &P.f32, st};
  launch_gather_windows_u8(track, n_frames, d_vstarts, c.buf("vid"), nb, T, H * W * 3, st);
  launch_gather_audio(mel_full, F, Ta_full, d_astarts, c.buf("aud"), nb, Ta, st);
  return forward_bf16_impl(h, s, logits, nullptr, ws, ws_bytes, st, true, nullptr, 0, 0, nullptr, 0);
}
