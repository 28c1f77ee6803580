// "Temporal ring" 3x3x3 convolution for the 64 -> 64 channel residual stage (conv_ring.cu): Conv3d(64 -> 64, 3x3x3, pad 1) + BN
// (+ residual) + ReLU of app/models/visual_encoder.py:46-87 (layer1) with the three temporal taps as three 64-column blocks of one
// N = 192 tcgen05 MMA — the scheme of stem_ring.cu with planar activations and streamed weights.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "stem_ring.cuh"
#include "umma_conv.cuh"

namespace lsd {

constexpr int CR_K16 = 4;                            // 64 input channels
constexpr int CR_TAPS = 9;                           // spatial taps (dh, dw)
constexpr int CR_WCHUNK = CR_TAPS * 2 * 192 * 16;    // bytes of the weights of one K16 chunk: [tap][2 K halves][192 columns][8 bf16]

struct ConvRingP {
  const __nv_bfloat16* x;       // input, plane 0 / position 0 (plain planar, geometry g, 8 planes)
  int64_t x_plane_stride;
  const __nv_bfloat16* w;       // packed weights [K16 chunk][tap][K half][dt block j: 0..2][64][8]
  const float* bias;            // 64
  const __nv_bfloat16* res;     // optional residual (plain planar, geometry g)
  int64_t res_plane_stride;
  __nv_bfloat16* y;             // destination: plain (UC_Y_PLAIN, geometry g) or parity-split (UC_Y_PARITY, geometry g2 per set)
  int64_t y_plane_stride, y_set_stride;
  int y_mode;
  UcGeom g, g2;
  const SrStep* steps;          // [2 * gridDim.x slots][nsteps] (same table format as the ring stem)
  int nsteps;
  int nst;                      // ring stages of the activation regions
  int start, units;             // region: first position relative to the chunk start, length in positions
  long long* dbg;               // optional (LSD_CR_TRACE): per-CTA spans
  int skip;
};

cudaError_t conv_ring_device_init();
size_t conv_ring_smem_bytes(const ConvRingP& p);
void launch_conv_ring(const ConvRingP& p, int grid, cudaStream_t s);

}  // namespace lsd
