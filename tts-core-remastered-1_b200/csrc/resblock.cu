// ResidualBlock (generator.py:40-41,89-90 / repair R2): weight packing and the dispatch to the two fused
// kernels -- resblock3.cu (C = 128, 256: streamed weights, CTA pairs) and resblock2.cu (C = 32, 64:
// resident weights, residual on the tensor core).  The first-generation one-tile-per-CTA kernel that
// used to live here is in the git history (commit e561616 and earlier).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

// ---------------------------------------------------------------------------- packing
// w_conv [2C][C][3] (value rows 0..C-1, gate rows C..2C-1)  ->  w1[row][tap*C + ci] with rows
// GLU-interleaved per chunk of CH channels: row = j*2CH + {0..CH-1: value ch j*CH+i | CH..: gate}.
// w_proj [C][C][1] -> w2[co][ci].  Layout of w_packed: w1 (2C*3C elems) then w2.
// w1 is pre-scaled by 1/2 (exact; the kernels evaluate sigmoid through tanh(g/2) and fold the other
// 1/2 into the value half).  Narrow stages (C <= 64, resblock2.cu): w2 is [C][2C] = [W_proj | I]: the
// identity block adds the residual x on the tensor core.
__global__ void pack_resblock_kernel(const float* __restrict__ w_conv, const float* __restrict__ w_proj, int C,
                                     int CH, int fmt, uint16_t* __restrict__ out) {
  const bool narrow = C <= 64;
  const int w2_cols = narrow ? 2 * C : C;
  const long long n1 = 2ll * C * 3 * C, total = n1 + (long long)C * w2_cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v;
    if (i < n1) {
      const int col = (int)(i % (3 * C)), row = (int)(i / (3 * C));
      const int tap = col / C, ci = col % C;
      const int j = row / (2 * CH), within = row % (2 * CH);
      const int src_row = within < CH ? (j * CH + within) : (C + j * CH + within - CH);
      v = w_conv[((long long)src_row * C + ci) * 3 + tap];
      v *= 0.5f;
    } else {
      const int col = (int)((i - n1) % w2_cols), row = (int)((i - n1) / w2_cols);
      v = col < C ? w_proj[(long long)row * C + col] : (col - C == row ? 1.0f : 0.0f);
    }
    out[i] = fmt == 0 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int pack_resblock_launch(const float* w_conv, const float* w_proj, int C, int fmt, void* out, cudaStream_t stream) {
  B200_CHECK_ARG(C == 32 || C == 64 || C == 128 || C == 256, "resblock: C=%d unsupported (32/64/128/256)", C);
  const int CH = C >= 64 ? 64 : 32;
  pack_resblock_kernel<<<1024, 256, 0, stream>>>(w_conv, w_proj, C, CH, fmt, reinterpret_cast<uint16_t*>(out));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int resblock2_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream);

int resblock3_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream);

int resblock_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                    const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                    int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(N > 0 && L > 0 && T > 0 && L % T == 0, "resblock: L=%d must be a multiple of T=%d", L, T);
  B200_CHECK_ARG(N % num_bands == 0, "resblock: N=%d not a multiple of num_bands=%d", N, num_bands);
  B200_CHECK_ARG(a16 != out16, "resblock: in-place is not supported (neighbour tiles read the halo)");
  switch (C) {
    case 32:
    case 64:
      // narrow stages: RAW input convention, weights packed for resblock2.cu only
      return resblock2_launch(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, C, dilation, T, num_bands, fmt,
                              out_fmt, store_lrelu, out16, stream);
    case 128:
    case 256:
      return resblock3_launch(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, C, dilation, T, num_bands, fmt,
                              out_fmt, store_lrelu, out16, stream);
  }
  set_error("resblock: C=%d unsupported (32/64/128/256)", C);
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
