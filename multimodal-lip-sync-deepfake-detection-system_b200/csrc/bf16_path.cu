// bf16 tensor-core path (tcgen05 implicit-GEMM convolutions).  Placeholder until umma_conv.cu lands.
#include "../../include/lsd_b200.h"
#include "lsd_internal.h"
#include "lsd_kernels.h"

int pack_bf16_weights(lsd_handle*, const std::vector<float>&) { return 0; }
void make_plan_bf16(lsd_handle*, int, int, int, int, int, int, std::vector<Stage>&, size_t& bytes) { bytes = 0; }
int forward_bf16(lsd_handle* h, int, int, int, int, int, int, const void*, int, int, const void*, int, float*, const lsd_aux*,
                 char*, size_t, cudaStream_t, bool) {
  return lsd_fail(h, LSD_ERR_UNSUPPORTED, "bf16 path not built");
}
int score_batch_bf16(lsd_handle* h, const uint8_t*, int, const int32_t*, const int32_t*, const float*, int, int, int, int, int, int,
                     int, float*, char*, size_t, cudaStream_t) {
  return lsd_fail(h, LSD_ERR_UNSUPPORTED, "bf16 path not built");
}
