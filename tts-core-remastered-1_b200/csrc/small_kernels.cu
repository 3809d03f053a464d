// fp32 CUDA-core kernels around the tensor-core layers.  These are the precision-sensitive,
// bandwidth-bound parts: conditioning MLPs (generator.py:65-73), FiLM projections
// (repair R2), band_split (generator.py:76-81), band_merge + tanh (generator.py:96-98), plus
// layout helpers.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  return fmt == 0 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t v, int fmt) {
  return fmt == 0 ? __half2float(__ushort_as_half(v)) : __bfloat162float(__ushort_as_bfloat16(v));
}

// ------------------------------------------------------------------ K5a: style + emotion
// se[b, c] = style_proj(style[b])*w_style*(!style_drop) ; emo likewise (kept separate so the final
// sum has the reference's association (c_pros + c_sty) + c_emo, generator.py:72).
__global__ void style_emo_kernel(const float* __restrict__ style, const float* __restrict__ emotion,
                                 const float* __restrict__ ws, const float* __restrict__ bs,
                                 const float* __restrict__ we, const float* __restrict__ be, int style_dim,
                                 int cond_dim, float w_style, float w_emo, int style_drop, int emo_drop,
                                 float* __restrict__ sty_out, float* __restrict__ emo_out) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < cond_dim; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < style_dim; ++k) s = fmaf(style[b * style_dim + k], ws[c * style_dim + k], s);
    s = (s + bs[c]) * w_style;
    float e = 0.f;
    for (int k = 0; k < 6; ++k) e = fmaf(emotion[b * 6 + k], we[c * 6 + k], e);
    e = (e + be[c]) * w_emo;
    sty_out[b * cond_dim + c] = style_drop ? 0.f : s;
    emo_out[b * cond_dim + c] = emo_drop ? 0.f : e;
  }
}

// ------------------------------------------------------------------ K5b: prosody MLP -> cond[B,T,cond_dim]
// cond = Linear(64->128)(SiLU(Linear(18->64)(prosody))) + sty + emo.  One block loops over frames;
// both weight matrices live in shared memory (W2 transposed so the inner loop is conflict-free).
__global__ void __launch_bounds__(128) cond_kernel(const float* __restrict__ prosody, const float* __restrict__ w0,
                                                   const float* __restrict__ b0, const float* __restrict__ w2,
                                                   const float* __restrict__ b2, const float* __restrict__ sty,
                                                   const float* __restrict__ emo, int B, int T,
                                                   float* __restrict__ cond) {
  constexpr int HID = 64, CD = 128, PIN = 18;
  __shared__ float sW0[HID * PIN];
  __shared__ float sW2t[HID * CD];   // [j][c]
  __shared__ float sHid[HID];
  __shared__ float sIn[PIN];
  for (int i = threadIdx.x; i < HID * PIN; i += blockDim.x) sW0[i] = w0[i];
  for (int i = threadIdx.x; i < HID * CD; i += blockDim.x) {
    const int c = i / HID, j = i % HID;
    sW2t[j * CD + c] = w2[i];
  }
  __syncthreads();
  const int c = threadIdx.x;
  for (long long f = blockIdx.x; f < (long long)B * T; f += gridDim.x) {
    const int b = (int)(f / T);
    if (threadIdx.x < PIN) sIn[threadIdx.x] = prosody[f * PIN + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < HID) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < PIN; ++k) a = fmaf(sIn[k], sW0[threadIdx.x * PIN + k], a);
      a += b0[threadIdx.x];
      sHid[threadIdx.x] = a / (1.0f + expf(-a));   // SiLU
    }
    __syncthreads();
    float o = 0.f;
#pragma unroll 8
    for (int j = 0; j < HID; ++j) o = fmaf(sHid[j], sW2t[j * CD + c], o);
    o += b2[c];
    cond[f * CD + c] = (o + sty[b * CD + c]) + emo[b * CD + c];
    __syncthreads();
  }
}

// ------------------------------------------------------------------ FiLM projection (fp32 SGEMM)
// out[r, n] = sum_k A[r, k] * W[n, k] + bias[n],  K = 128 fixed, r < M, n < Ncols (multiple of 64).
// 64x64 tile per block, 4x4 per thread, both operands transposed into smem ([k][m], [k][n]).
__global__ void __launch_bounds__(256) film_sgemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int M, int Ncols,
                                                         float* __restrict__ out) {
  constexpr int K = 128, KH = 64, BM = 64, BN = 64, PAD = 4;
  __shared__ float sA[KH][BM + PAD];
  __shared__ float sB[KH][BN + PAD];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tm = (threadIdx.x / 16) * 4, tn = (threadIdx.x % 16) * 4;
  float acc[4][4] = {};
  for (int kh = 0; kh < K; kh += KH) {
    if (kh) __syncthreads();
    for (int i = threadIdx.x; i < BM * (KH / 4); i += 256) {
      const int r = i / (KH / 4), k4 = i % (KH / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * K + kh + k4 * 4);
      sA[k4 * 4 + 0][r] = v.x; sA[k4 * 4 + 1][r] = v.y; sA[k4 * 4 + 2][r] = v.z; sA[k4 * 4 + 3][r] = v.w;
      const float4 w = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * K + kh + k4 * 4);
      sB[k4 * 4 + 0][r] = w.x; sB[k4 * 4 + 1][r] = w.y; sB[k4 * 4 + 2][r] = w.z; sB[k4 * 4 + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KH; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[k][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[k][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + tn);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (m0 + tm + i < M) {
      float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
      *reinterpret_cast<float4*>(out + (long long)(m0 + tm + i) * Ncols + n0 + tn) = o;
    }
  }
}

// ------------------------------------------------------------------ K3: band_split
// out16[(b*nb + band), t, co] = bias[band][co] + sum_{ci,k} wt[band][ci*7+k][co] * mel[b][band*bs+ci][t+k-3]
// Block = (t-tile of 32 frames, band, b); thread = output channel (co, co+256, ...).
constexpr int kSplitTT = 32;
__global__ void __launch_bounds__(256) band_split_kernel(const float* __restrict__ mel, const float* __restrict__ wt,
                                                         const float* __restrict__ bias, int B, int channels,
                                                         int band_size, int T, int H, int fmt,
                                                         uint16_t* __restrict__ out) {
  extern __shared__ float s_mel[];   // [band_size][kSplitTT + 8]
  constexpr int W = kSplitTT + 8;
  const int t0 = blockIdx.x * kSplitTT, band = blockIdx.y, b = blockIdx.z;
  const int nb = channels / band_size;
  for (int i = threadIdx.x; i < band_size * W; i += blockDim.x) {
    const int ci = i / W, tt = i % W;
    const int t = t0 + tt - 3;
    s_mel[i] = (t >= 0 && t < T && tt < kSplitTT + 6) ? mel[((long long)b * channels + band * band_size + ci) * T + t] : 0.f;
  }
  __syncthreads();
  const float* wb = wt + (long long)band * band_size * 7 * H;
  for (int co = threadIdx.x; co < H; co += blockDim.x) {
    float acc[kSplitTT];
    const float bv = bias[band * H + co];
#pragma unroll
    for (int i = 0; i < kSplitTT; ++i) acc[i] = bv;
    for (int ci = 0; ci < band_size; ++ci) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float w = __ldg(wb + (long long)(ci * 7 + k) * H + co);
        const float* m = s_mel + ci * W + k;
#pragma unroll
        for (int i = 0; i < kSplitTT; ++i) acc[i] = fmaf(w, m[i], acc[i]);
      }
    }
    uint16_t* o = out + (((long long)(b * nb + band)) * T + t0) * H + co;
#pragma unroll
    for (int i = 0; i < kSplitTT; ++i)
      if (t0 + i < T) o[(long long)i * H] = to16(acc[i], fmt);
  }
}

// reference Conv1d weight [H][band_size][7] per band -> wt[band][ci*7+k][H]
__global__ void pack_split_kernel(const float* __restrict__ w, int band_size, int H, float* __restrict__ wt_band) {
  const int total = H * band_size * 7;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i / (band_size * 7), r = i % (band_size * 7);
    wt_band[(long long)r * H + co] = w[i];
  }
}

// ------------------------------------------------------------------ K4: band_merge + tanh
// wav[b, l] = tanh(bias + sum_{band,c,k} w[band*Cb + c][k] * x[(b*nb+band), l+k-3, c]); x16 raw,
// channels-last with Cb = 32 channels (64-byte rows).  Thread = output sample.
template <int CB>
__global__ void __launch_bounds__(256) band_merge_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int nb, int L,
                                                         int fmt, float* __restrict__ wav) {
  extern __shared__ float s_w[];   // [nb][7][CB]
  for (int i = threadIdx.x; i < nb * 7 * CB; i += blockDim.x) {
    const int band = i / (7 * CB), k = (i / CB) % 7, c = i % CB;
    s_w[i] = w[(band * CB + c) * 7 + k];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  float acc = bias[0];
  for (int band = 0; band < nb; ++band) {
    const uint16_t* xs = x + ((long long)(b * nb + band) * L) * CB;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int ll = l + k - 3;
      if (ll < 0 || ll >= L) continue;
      const uint4* row = reinterpret_cast<const uint4*>(xs + (long long)ll * CB);
      const float* wk = s_w + (band * 7 + k) * CB;
#pragma unroll
      for (int v = 0; v < CB / 8; ++v) {
        const uint4 u = __ldg(row + v);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2(uw[e], fmt);
          acc = fmaf(f.x, wk[v * 8 + e * 2], acc);
          acc = fmaf(f.y, wk[v * 8 + e * 2 + 1], acc);
        }
      }
    }
  }
  wav[(long long)b * L + l] = tanhf(acc);
}

// ------------------------------------------------------------------ helpers
// 16-bit channels-last [N, L, C] -> fp32 channels-first [N, C, L] (undoing the lrelu storage form).
__global__ void tap_extract_kernel(const uint16_t* __restrict__ x, int N, int L, int C, int fmt, int stored_lrelu,
                                   float* __restrict__ out) {
  const long long total = (long long)N * L * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long nl = i / C;
    const int l = (int)(nl % L);
    const long long n = nl / L;
    float v = from16(x[i], fmt);
    if (stored_lrelu) v = lrelu_inv(v);
    out[(n * C + c) * L + l] = v;
  }
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n, float add,
                                float mul) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i] * mul + add;
}
__global__ void cvt16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n, int fmt,
                             float mul) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = to16(src[i] * mul, fmt);
}

// ------------------------------------------------------------------ launchers
static inline int grid_for(long long n, int block = 256, int cap = 8192) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int style_emo_launch(const float* style, const float* emotion, const float* ws, const float* bs, const float* we,
                     const float* be, int B, int style_dim, int cond_dim, float w_style, float w_emo, int style_drop,
                     int emo_drop, float* sty_out, float* emo_out, cudaStream_t st) {
  style_emo_kernel<<<B, 128, 0, st>>>(style, emotion, ws, bs, we, be, style_dim, cond_dim, w_style, w_emo,
                                      style_drop, emo_drop, sty_out, emo_out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int cond_launch(const float* prosody, const float* w0, const float* b0, const float* w2, const float* b2,
                const float* sty, const float* emo, int B, int T, float* cond, cudaStream_t st) {
  long long frames = (long long)B * T;
  int grid = (int)(frames < 592 ? frames : 592);
  cond_kernel<<<grid, 128, 0, st>>>(prosody, w0, b0, w2, b2, sty, emo, B, T, cond);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int film_launch(const float* cond, const float* w_all, const float* b_all, int M, int ncols, float* out,
                cudaStream_t st) {
  B200_CHECK_ARG(ncols % 64 == 0, "film: ncols=%d must be a multiple of 64", ncols);
  dim3 grid(ceil_div(M, 64), ncols / 64);
  film_sgemm_kernel<<<grid, 256, 0, st>>>(cond, w_all, b_all, M, ncols, out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int band_split_launch(const float* mel, const float* wt, const float* bias, int B, int channels, int band_size, int T,
                      int H, int fmt, void* out16, cudaStream_t st) {
  dim3 grid(ceil_div(T, kSplitTT), channels / band_size, B);
  size_t smem = (size_t)band_size * (kSplitTT + 8) * sizeof(float);
  band_split_kernel<<<grid, 256, smem, st>>>(mel, wt, bias, B, channels, band_size, T, H, fmt,
                                             reinterpret_cast<uint16_t*>(out16));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int pack_split_launch(const float* w, int band_size, int H, float* wt_band, cudaStream_t st) {
  pack_split_kernel<<<grid_for((long long)H * band_size * 7), 256, 0, st>>>(w, band_size, H, wt_band);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int band_merge_launch(const void* x16, const float* w, const float* bias, int B, int nb, int L, int Cb, int fmt,
                      float* wav, cudaStream_t st) {
  B200_CHECK_ARG(Cb == 32, "band_merge: per-band channels %d unsupported (32)", Cb);
  dim3 grid(ceil_div(L, 256), B);
  band_merge_kernel<32><<<grid, 256, (size_t)nb * 7 * 32 * sizeof(float), st>>>(
      reinterpret_cast<const uint16_t*>(x16), w, bias, B, nb, L, fmt, wav);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int tap_extract_launch(const void* x16, int N, int L, int C, int fmt, int stored_lrelu, float* out, cudaStream_t st) {
  tap_extract_kernel<<<grid_for((long long)N * L * C), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(x16), N, L, C,
                                                                     fmt, stored_lrelu, out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int copy_f32_launch(const float* src, float* dst, long long n, float add, cudaStream_t st, float mul) {
  copy_f32_kernel<<<grid_for(n), 256, 0, st>>>(src, dst, n, add, mul);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int cvt16_launch(const float* src, void* dst, long long n, int fmt, cudaStream_t st, float mul) {
  cvt16_kernel<<<grid_for(n), 256, 0, st>>>(src, reinterpret_cast<uint16_t*>(dst), n, fmt, mul);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
