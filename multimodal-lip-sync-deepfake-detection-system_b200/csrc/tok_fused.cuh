// Fused temporal-transformer kernel (tok_fused.cu): the four pre-norm encoder layers of TemporalTransformer.forward
// (app/models/temporal.py:64-77,107-111) in ONE launch, a CTA per group of windows, activations resident on the SM.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace lsd {

constexpr int TF_LAYERS = 4, TF_D = 256, TF_HEADS = 8, TF_FF = 1024;
constexpr int TF_SLOT_BYTES = 16384;      // largest weight-ring stage

// One stage of the weight stream = `kps` K16 steps of a (N x K) block, packed [k16][2 planes][N][8] fp16.
// Order of the blocks of one layer (host packer and device issue loop walk the same list, see tf_layer_blocks):
//   QKV(0), OUT(0), QKV(1), OUT(1), QKV(2), OUT(2), QKV(3), OUT(3), FF1(0), FF1(1), FF2(0), FF1(2), FF2(1), ..., FF1(7), FF2(6), FF2(7)
// where the device issues OUT(hp-1) together with QKV(hp), and FF2(c) together with FF1(c+2) while the GELU of chunk c+1 runs.
struct TfBlock { int kind, idx, N, k16, kps; };   // kind: 0 QKV, 1 OUT, 2 FF1, 3 FF2
inline void tf_layer_blocks(std::vector<TfBlock>& out) {
  out.clear();
  for (int hp = 0; hp < 4; ++hp) {
    out.push_back({0, hp, 192, 16, 2});   // Q|K|V rows of heads 2hp, 2hp+1; K = 256 input channels
    out.push_back({1, hp, 256, 4, 2});    // out_proj columns, K = the 64 attention channels of this head pair
  }
  // (stream order: the device consumes QKV(hp) before OUT(hp); OUT(hp) is issued in the same MMA phase as QKV(hp+1))
  // FFN, software-pipelined: FF1(0), FF1(1), then FF2(c) followed by FF1(c+2)
  out.push_back({2, 0, 128, 16, 4});      // linear1 rows 128c..128c+127; K = 256
  out.push_back({2, 1, 128, 16, 4});
  for (int c = 0; c < 8; ++c) {
    out.push_back({3, c, 256, 8, 2});     // linear2 all rows, K = hidden 128c..128c+127
    if (c + 2 < 8) out.push_back({2, c + 2, 128, 16, 4});
  }
}

struct TokFusedP {
  float* tok;                 // [B][NT][256] fp32 rows, in / out
  const __half* w;            // packed weight stream, TF_LAYERS consecutive layer images
  const uint32_t* stage_bytes;// device: bytes of each ring stage of ONE layer (n_stage_layer entries)
  int n_stage_layer;
  uint32_t layer_bytes;       // bytes of one layer image
  const float* vec;           // device: per layer [ln1_g 256 | ln1_b 256 | ln2_g 256 | ln2_b 256 | qkv_bias 768 (packed order) | ff1_bias 1024],
                              //         then (2*TF_LAYERS + 1) cumulative bias vectors of 256 (see tok_fused.cu)
  int B, NT, SL, G, KW;       // windows, tokens per window (T+1), row slot per window (32 or 64), windows per CTA, keys per window (16-multiple)
  void* dbg;                  // optional (LSD_TOKF_TRACE): 384 x int64 of phase timestamps of CTA 0, see tok_fused.cu
};
constexpr int TF_VEC_LAYER = 4 * 256 + 768 + 1024;

cudaError_t tok_fused_device_init();
bool tok_fused_supported(int NT);
void tok_fused_geometry(int NT, int& SL, int& G, int& KW);
void launch_tok_fused(const TokFusedP& p, cudaStream_t s);

}  // namespace lsd
