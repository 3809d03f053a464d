// Host-side helpers shared by the launchers: status codes, last-error string, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200voc.h"

namespace b200 {

void set_error(const char* fmt, ...);
const char* get_error();

#define B200_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::b200::set_error(__VA_ARGS__);             \
      return B200VOC_ERR_BAD_ARG;                 \
    }                                             \
  } while (0)

#define B200_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B200VOC_ERR_CUDA;                                                               \
    }                                                                                        \
  } while (0)

#define B200_TRY(call)            \
  do {                            \
    int s__ = (call);             \
    if (s__ != B200VOC_OK) return s__; \
  } while (0)

// 16-bit row-major tensor maps.  `swizzle_bytes` is 128 or 64 and must equal box0 * 2.
// dims/strides are innermost-first; strides (bytes) are given for dims 1.. (dim 0 is contiguous).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t box0,
                 uint32_t box1, int swizzle_bytes);
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box0, uint32_t box1, int swizzle_bytes);

int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint64_t stride3_bytes, uint32_t box0, uint32_t box1,
                 uint32_t box2, int swizzle_bytes);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One launch of the fused narrow-stage kernel (stage_fused.cu).  x_in: in_ct ? [N, Lin, 2C] raw (ConvT stride 2 runs
// first, L = 2 Lin) : [N, Lin, C] leaky_relu(x) (L = Lin).  out_mode: 0 = store leaky_relu(x), 1 = store raw x,
// 2 = band_merge + tanh -> wav.
struct StageFusedArgs {
  const void* x_in;
  int N, Lin, C, T, num_bands, fmt;
  int in_ct, nblk, out_mode;
  const void* ct_w; const float* ct_b;
  const void* blk_w[3]; const float* b_conv[3]; const float* b_proj[3]; int dil[3]; int film_col[3];
  const float* film; int film_stride;
  void* out16;
  const void* merge_w16; const float* merge_b; const int* valid_samples; int pcm16; void* wav;
};
int stage_fused_launch(const StageFusedArgs& a, cudaStream_t stream);

}  // namespace b200
