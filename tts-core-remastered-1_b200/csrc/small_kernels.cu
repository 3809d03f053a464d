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
                                                   float* __restrict__ cond, uint16_t* __restrict__ cond3) {
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
    const float v = (o + sty[b * CD + c]) + emo[b * CD + c];
    cond[f * CD + c] = v;
    if (cond3 != nullptr) {
      // split-fp16 operand of the tensor-core FiLM GEMM: [hi | lo | hi] (see film_tc_launch / splitgemm_kernel)
      const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
      uint16_t* r3 = cond3 + f * (3 * CD);
      r3[c] = __half_as_ushort(hi);
      r3[CD + c] = __half_as_ushort(lo);
      r3[2 * CD + c] = __half_as_ushort(hi);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ FiLM projection (fp32 SGEMM)
// out[r, n] = sum_k A[r, k] * W[n, k] + bias[n],  K = 128 fixed, r < M, n < Ncols (multiple of 128).
// 128x128 tile per block, 8x8 per thread (16 FMA per shared-memory vector load), K in chunks of 16,
// operands transposed into smem ([k][m], [k][n]).  fp32 on CUDA cores: the FiLM projection is one
// of the precision-sensitive layers.
__global__ void __launch_bounds__(256) film_sgemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int M, int Ncols,
                                                         float* __restrict__ out) {
  constexpr int K = 128, BK = 16, BM = 128, BN = 128, PAD = 4;
  __shared__ __align__(16) float sA[BK][BM + PAD];
  __shared__ __align__(16) float sB[BK][BN + PAD];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tm = (threadIdx.x / 16) * 8, tn = (threadIdx.x % 16) * 8;
  float acc[8][8] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    if (k0) __syncthreads();
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int i = threadIdx.x + v * 256;         // 512 float4 per operand chunk
      const int r = i >> 2, k4 = i & 3;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) a = __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * K + k0 + k4 * 4));
      sA[k4 * 4 + 0][r] = a.x; sA[k4 * 4 + 1][r] = a.y; sA[k4 * 4 + 2][r] = a.z; sA[k4 * 4 + 3][r] = a.w;
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + r) * K + k0 + k4 * 4));
      sB[k4 * 4 + 0][r] = w.x; sB[k4 * 4 + 1][r] = w.y; sB[k4 * 4 + 2][r] = w.z; sB[k4 * 4 + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[k][tm]), a1 = *reinterpret_cast<const float4*>(&sA[k][tm + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[k][tn]), b1 = *reinterpret_cast<const float4*>(&sB[k][tn + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  const float4 c0 = *reinterpret_cast<const float4*>(bias + n0 + tn), c1 = *reinterpret_cast<const float4*>(bias + n0 + tn + 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (m0 + tm + i < M) {
      float* o = out + (long long)(m0 + tm + i) * Ncols + n0 + tn;
      *reinterpret_cast<float4*>(o) = make_float4(acc[i][0] + c0.x, acc[i][1] + c0.y, acc[i][2] + c0.z, acc[i][3] + c0.w);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[i][4] + c1.x, acc[i][5] + c1.y, acc[i][6] + c1.z, acc[i][7] + c1.w);
    }
  }
}

// ------------------------------------------------------------------ FiLM projection on the tensor cores
// Same contraction with fp32-level accuracy from fp16 operands: x = hi + lo (hi = fp16(x), lo = fp16(x - hi), 22
// significant bits), and  A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T  (the lo*lo term is 2^-22 relative), i.e.
// ONE K = 384 GEMM of A' = [A_hi | A_lo | A_hi] (written by cond_kernel) with W' = [W_hi | W_hi | W_lo] (packed at
// load time), fp32 accumulation in TMEM, fp32 output (splitgemm_kernel below).  34 TFLOP/s of CUDA-core SGEMM was
// 3 % of the step.
__global__ void pack_film3_kernel(const float* __restrict__ w, long long rows, uint16_t* __restrict__ w3) {
  const long long total = rows * 128;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / 128;
    const int k = (int)(i % 128);
    const float v = w[i];
    const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
    uint16_t* o = w3 + r * 384;
    o[k] = __half_as_ushort(hi);
    o[128 + k] = __half_as_ushort(hi);
    o[256 + k] = __half_as_ushort(lo);
  }
}
int pack_film3_launch(const float* w, long long rows, void* w3, cudaStream_t st) {
  const long long blocks = (rows * 128 + 255) / 256;
  pack_film3_kernel<<<(int)(blocks < 4096 ? blocks : 4096), 256, 0, st>>>(w, rows, reinterpret_cast<uint16_t*>(w3));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

// Shared mainloop of the two split-fp16 GEMMs (FiLM, band_split): one CTA = one 128 x 128 output tile, the K
// dimension streams through a 3-stage TMA ring of (A k-block, W k-block) pairs, 96 KB of shared memory so that
// TWO CTAs are resident per SM and one's loads / epilogue overlap the other's MMAs.
constexpr int kSgStages = 3, kSgStageBytes = 2 * 16384, kSgSmem = kSgStages * kSgStageBytes + 128 + 1024;
struct SplitGemmParams {
  int M, KB;             // rows of A per band, k-blocks of 64
  int T, H, nb, fmt;     // OUT16 (band_split): frames per utterance, row pitch, bands, 16-bit format
  int ncols;             // !OUT16 (FiLM): row pitch of the fp32 output
  const float* bias;
  void* out;
};
template <bool OUT16>
__global__ void __launch_bounds__(192, 2)
splitgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO, const SplitGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSgStages * kSgStageBytes);
  uint64_t* empty = full + kSgStages;
  uint64_t* acc_full = empty + kSgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128, band = blockIdx.z;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kSgStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % kSgStages;
        mbar_wait(&empty[s], ((kb / kSgStages) & 1) ^ 1);
        mbar_expect_tx(&full[s], kSgStageBytes);
        uint8_t* st = smem + s * kSgStageBytes;
        tma_load_3d(st, &tmA, &full[s], kb * 64, m0, band);              // rows past M are zero-filled
        tma_load_3d(st + 16384, &tmW, &full[s], kb * 64, n0, band);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(OUT16 ? 0 : p.fmt, 128)   /* fp32-output form: fmt 0 = fp16 operands, 1 = bf16 */;
    for (int kb = 0; kb < p.KB; ++kb) {
      const int s = kb % kSgStages;
      mbar_wait(&full[s], (kb / kSgStages) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kSgStageBytes);
      const uint64_t a_desc = make_kmajor_desc<128>(a_addr), b_desc = make_kmajor_desc<128>(a_addr + 16384);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
        if (kb == p.KB - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quadrant warp % 4; thread = output row.  Each warp stages its 32 rows in the
    // (now dead: every MMA that read it has completed) ring as 128-byte-swizzled boxes and hands them to TMA, so
    // global writes are full lines and rows past M / past the end of an utterance are clipped by the tensor map.
    const int q = warp & 3, row0 = m0 + q * 32;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    uint8_t* stg = smem + q * 16384;
    const uint32_t sw = (uint32_t)(lane & 7);
    if (row0 < p.M) {
      if (OUT16) {
        const float* bb = p.bias + band * p.H + n0;
        // TMA path only for warps whose 32 rows lie inside ONE utterance (row pitch changes at an utterance
        // boundary; a box that starts at a negative frame index is not a legal store).  The few boundary warps
        // (and every warp when T < 32) write their rows directly.
        const int m = row0 + lane, b_first = row0 / p.T;
        const bool use_tma = b_first == min(row0 + 31, p.M - 1) / p.T;
        const int b = m < p.M ? m / p.T : 0, t = m < p.M ? m - b * p.T : 0;
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + (((long long)(b * p.nb + band)) * p.T + t) * p.H + n0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          uint8_t* box = stg + (c >> 1) * 4096 + lane * 128;       // box = 32 rows x 64 halves
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float y0 = __uint_as_float(v[8 * j + 2 * e]) + __ldg(bb + c * 32 + 8 * j + 2 * e);
              const float y1 = __uint_as_float(v[8 * j + 2 * e + 1]) + __ldg(bb + c * 32 + 8 * j + 2 * e + 1);
              w[e] = pack2(y0, y1, p.fmt);
            }
            if (use_tma)
              *reinterpret_cast<uint4*>(box + ((((uint32_t)(4 * (c & 1) + j)) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            else if (m < p.M)
              *reinterpret_cast<uint4*>(o + c * 32 + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if (use_tma && (c & 1)) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmO, stg + (c >> 1) * 4096, n0 + 64 * (c >> 1), row0 - b_first * p.T, band, b_first);
              tma_store_commit();
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          uint8_t* box = stg + c * 4096 + lane * 128;              // box = 32 rows x 32 floats
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            *reinterpret_cast<float4*>(box + (((uint32_t)j ^ sw) << 4)) =
                make_float4(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y,
                            __uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && n0 + c * 32 < p.ncols) {                           // boxes past the last column: nothing to store
            tma_store_3d(&tmO, stg + c * 4096, 2 * (n0 + c * 32), row0, 0);   // the map counts 16-bit units
            tma_store_commit();
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}
template <bool OUT16>
static int launch_splitgemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO,
                            const SplitGemmParams& p, dim3 grid, cudaStream_t st) {
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(splitgemm_kernel<OUT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmem));
    configured[dev & 15] = true;
  }
  splitgemm_kernel<OUT16><<<grid, 192, kSgSmem, st>>>(tmA, tmW, tmO, p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int film_tc_launch(const void* cond3, const void* w3, const float* b_all, int M, int ncols, float* out, cudaStream_t st) {
  B200_CHECK_ARG(ncols % 128 == 0, "film: ncols=%d must be a multiple of 128", ncols);
  CUtensorMap tmA, tmW;
  B200_TRY(make_tmap_3d(&tmA, cond3, 384, M, 1, 384 * 2, (uint64_t)M * 384 * 2, 64, 128, 128));
  B200_TRY(make_tmap_3d(&tmW, w3, 384, ncols, 1, 384 * 2, (uint64_t)ncols * 384 * 2, 64, 128, 128));
  // fp32 output seen through a 16-bit map: 2 units per float, box = 32 rows x 128 bytes
  CUtensorMap tmO;
  B200_TRY(make_tmap_3d(&tmO, out, 2ull * ncols, M, 1, (uint64_t)ncols * 4, (uint64_t)M * ncols * 4, 64, 32, 128));
  SplitGemmParams p{};
  p.M = M; p.KB = 6; p.ncols = ncols; p.bias = b_all; p.out = out;
  return launch_splitgemm<false>(tmA, tmW, tmO, p, dim3(ceil_div(M, 128), ncols / 128, 1), st);
}

// General form of the same kernel: out[M][N] (fp32, row pitch N, N % 4 == 0) = A[M][Kdim] . W[N][Kdim]^T + bias, both
// operands 16-bit K-major (fmt 0 = fp16, 1 = bf16), Kdim a multiple of 64; rows past M / N are zero-filled on load and
// clipped on store by the tensor maps.  `bias` must hold N rounded up to 128 floats.  The critics' wgrad (disc_bwd.cu)
// runs through this with split-bf16 operands.
int splitgemm_f32_launch(const void* A, const void* W, const float* bias, int M, int N, long long Kdim, int fmt,
                         float* out, cudaStream_t st) {
  B200_CHECK_ARG(A && W && bias && out, "splitgemm: null argument");
  B200_CHECK_ARG(M > 0 && N > 0 && N % 4 == 0 && Kdim > 0 && Kdim % 64 == 0 && Kdim / 64 < (1ll << 30),
                 "splitgemm: bad shape (M=%d N=%d K=%lld)", M, N, Kdim);
  CUtensorMap tmA, tmW, tmO;
  B200_TRY(make_tmap_3d(&tmA, A, (uint64_t)Kdim, M, 1, (uint64_t)Kdim * 2, (uint64_t)M * Kdim * 2, 64, 128, 128));
  B200_TRY(make_tmap_3d(&tmW, W, (uint64_t)Kdim, N, 1, (uint64_t)Kdim * 2, (uint64_t)N * Kdim * 2, 64, 128, 128));
  B200_TRY(make_tmap_3d(&tmO, out, 2ull * N, M, 1, (uint64_t)N * 4, (uint64_t)M * N * 4, 64, 32, 128));
  SplitGemmParams p{};
  p.M = M; p.KB = (int)(Kdim / 64); p.ncols = N; p.bias = bias; p.out = out; p.fmt = fmt;
  return launch_splitgemm<false>(tmA, tmW, tmO, p, dim3(ceil_div(M, 128), ceil_div(N, 128), 1), st);
}

// ------------------------------------------------------------------ K3 on the tensor cores: band_split as 4 GEMMs
// The four Conv1d(20 -> H, k7) are GEMMs with K = 140 (im2col of the band's 20 mel bins x 7 taps).  Same split-fp16
// scheme as the FiLM GEMM (splitgemm_kernel; fp32-level accuracy): A' = [hi | lo | hi | 0] (3 x 140 padded to 448 = 7 k-blocks),
// W' = [W_hi | W_hi | W_lo | 0], fp32 accumulation in TMEM, +bias, 16-bit channels-last store.
constexpr int kSplitK = 448, kSplitKB = 7;
__global__ void __launch_bounds__(256) band_im2col3_kernel(const float* __restrict__ mel, int B, int channels, int band_size,
                                                           int T, int time_major, uint16_t* __restrict__ a3) {
  // a3[band][b*T + t][448]; one thread per (band, frame, j): j = ci*7 + k < band_size*7
  const int KJ = band_size * 7, nb = channels / band_size;
  const long long total = (long long)nb * B * T * KJ;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % KJ);
    const long long r = i / KJ;                      // band * (B*T) + b*T + t
    const int t = (int)(r % T);
    const int b = (int)((r / T) % B), band = (int)(r / ((long long)B * T));
    const int ci = j / 7, tt = t + j % 7 - 3;
    float v = 0.f;
    if (tt >= 0 && tt < T)
      v = time_major ? mel[((long long)b * T + tt) * channels + band * band_size + ci]
                     : mel[((long long)b * channels + band * band_size + ci) * T + tt];
    const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
    uint16_t* o = a3 + r * kSplitK;
    o[j] = __half_as_ushort(hi);
    o[KJ + j] = __half_as_ushort(lo);
    o[2 * KJ + j] = __half_as_ushort(hi);
    if (j < kSplitK - 3 * KJ) o[3 * KJ + j] = 0;     // zero padding (28 columns for KJ = 140)
  }
}
// w[H][band_size*7] fp32 (the reference layout [H][ci][k]) -> w3[H][448] = [hi | hi | lo | 0]
__global__ void pack_split3_kernel(const float* __restrict__ w, int KJ, int H, uint16_t* __restrict__ w3) {
  const long long total = (long long)H * KJ;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % KJ);
    const long long r = i / KJ;
    const float v = w[i];
    const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
    uint16_t* o = w3 + r * kSplitK;
    o[j] = __half_as_ushort(hi);
    o[KJ + j] = __half_as_ushort(hi);
    o[2 * KJ + j] = __half_as_ushort(lo);
    if (j < kSplitK - 3 * KJ) o[3 * KJ + j] = 0;
  }
}
int pack_split3_launch(const float* w, int band_size, int H, void* w3, cudaStream_t st) {
  B200_CHECK_ARG(3 * band_size * 7 <= kSplitK, "band_split: band of %d bins needs K > %d", band_size, kSplitK);
  pack_split3_kernel<<<256, 256, 0, st>>>(w, band_size * 7, H, reinterpret_cast<uint16_t*>(w3));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

long long band_split_tc_scratch_elems(int B, int T, int nb) { return (long long)nb * B * T * kSplitK; }
int band_split_tc_launch(const float* mel, const void* w3, const float* bias, int B, int channels, int band_size, int T,
                         int H, int fmt, int time_major, void* a3, void* out16, cudaStream_t st) {
  const int nb = channels / band_size, M = B * T;
  B200_CHECK_ARG(H % 128 == 0 && 3 * band_size * 7 <= kSplitK, "band_split: H=%d / band_size=%d unsupported", H, band_size);
  const long long total = (long long)nb * M * band_size * 7;
  band_im2col3_kernel<<<(int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192), 256, 0, st>>>(
      mel, B, channels, band_size, T, time_major, reinterpret_cast<uint16_t*>(a3));
  B200_CUDA(cudaGetLastError());
  CUtensorMap tmA, tmW;
  B200_TRY(make_tmap_3d(&tmA, a3, kSplitK, M, nb, (uint64_t)kSplitK * 2, (uint64_t)M * kSplitK * 2, 64, 128, 128));
  B200_TRY(make_tmap_3d(&tmW, w3, kSplitK, H, nb, (uint64_t)kSplitK * 2, (uint64_t)H * kSplitK * 2, 64, 128, 128));
  CUtensorMap tmO;   // out16 [B][nb][T][H]
  B200_TRY(make_tmap_4d(&tmO, out16, H, T, nb, B, (uint64_t)H * 2, (uint64_t)T * H * 2, (uint64_t)nb * T * H * 2, 64, 32,
                        1, 128));
  SplitGemmParams p{};
  p.M = M; p.KB = kSplitKB; p.T = T; p.H = H; p.nb = nb; p.fmt = fmt; p.bias = bias; p.out = out16;
  return launch_splitgemm<true>(tmA, tmW, tmO, p, dim3(ceil_div(M, 128), H / 128, nb), st);
}

// ------------------------------------------------------------------ K3: band_split
// out16[(b*nb + band), t, co] = bias[band][co] + sum_{ci,k} wt[band][ci*7+k][co] * mel[b][band*bs+ci][t+k-3]
// Block = (t-tile of 32 frames, band, b); thread = output channel (co, co+256, ...).
constexpr int kSplitTT = 32;
__global__ void __launch_bounds__(256) band_split_kernel(const float* __restrict__ mel, const float* __restrict__ wt,
                                                         const float* __restrict__ bias, int B, int channels,
                                                         int band_size, int T, int H, int fmt, int time_major,
                                                         uint16_t* __restrict__ out) {
  extern __shared__ float s_mel[];   // [band_size][kSplitTT + 8]
  constexpr int W = kSplitTT + 8;
  const int t0 = blockIdx.x * kSplitTT, band = blockIdx.y, b = blockIdx.z;
  const int nb = channels / band_size;
  // time_major: mel is [B, T, channels] (the refiner / acoustic output layout, sde_refiner5/model.py:304-306;
  // vocoder7/trainer.py:77 transposes it on the host) -- the transpose is folded into this load.
  // Either way consecutive threads read consecutive addresses (t fastest, or channel fastest).
  for (int i = threadIdx.x; i < band_size * W; i += blockDim.x) {
    const int ci = time_major ? i % band_size : i / W, tt = time_major ? i / band_size : i % W;
    const int t = t0 + tt - 3;
    const long long src = time_major ? ((long long)b * T + t) * channels + band * band_size + ci
                                     : ((long long)b * channels + band * band_size + ci) * T + t;
    s_mel[ci * W + tt] = (t >= 0 && t < T && tt < kSplitTT + 6) ? mel[src] : 0.f;
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
// channels-last, Cb = 32 channels per band (64-byte rows), nb = 4 bands (= 128 input channels).
// One warp produces 32 consecutive samples: lane = (band, 4 channels) keeps its 4x7 weights in
// registers, slides over the 38 input rows (each warp load = four full 64-byte rows, one per band)
// accumulating 32 partial outputs, then a 5-stage transpose-reduce leaves output j in lane j
// (31 shuffles per 32 outputs), tanh, one coalesced 128-byte store.  HBM-bound by design.
// Output formats (the wire formats after the path): fp32 in (-1, 1), or 16-bit PCM
// round(clamp(wav, -1, 1) * 32767).  `valid` (optional, [B]) = number of valid samples per utterance
// (hop * frame_length from the collator, batching2/colate.py:140-146,184-191): the padded tail is zeroed.
template <int FMT, int PCM16>
__global__ void __launch_bounds__(256) band_merge_kernel(const uint16_t* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int L,
                                                         const int* __restrict__ valid, void* __restrict__ wav_out) {
  constexpr int CB = 32, NB = 4;
  const int lane = threadIdx.x & 31;
  const int tiles = (L + 31) / 32;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= (long long)B * tiles) return;
  const int b = (int)(gw / tiles), l0 = (int)(gw % tiles) * 32;
  const int band = lane >> 3, c0 = (lane & 7) * 4;
  float wr[4][7];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 7; ++k) wr[i][k] = __ldg(w + (band * CB + c0 + i) * 7 + k);
  const uint16_t* xs = x + ((long long)(b * NB + band) * L) * CB + c0;
  float p[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) p[j] = 0.f;
#pragma unroll
  for (int r = 0; r < 38; ++r) {
    const int l = l0 + r - 3;
    float xf[4] = {0.f, 0.f, 0.f, 0.f};
    if (l >= 0 && l < L) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(xs + (long long)l * CB));
      const float2 f0 = unpack2t<FMT>(u.x), f1 = unpack2t<FMT>(u.y);
      xf[0] = f0.x; xf[1] = f0.y; xf[2] = f1.x; xf[3] = f1.y;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int j = r - k;               // output l0 + j uses input row l0 + j + k - 3 = row index r
      if (j >= 0 && j < 32) {
        float a = p[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(wr[i][k], xf[i], a);
        p[j] = a;
      }
    }
  }
  // transpose-reduce over the 32 lanes: after stage s each lane keeps s values
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? p[i] : p[i + s];
      const float keep = up ? p[i + s] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  const int l = l0 + lane;
  if (l < L) {
    float y = tanhf(p[0] + __ldg(bias));
    if (valid != nullptr && l >= __ldg(valid + b)) y = 0.f;
    if (PCM16) {
      const float c = fminf(fmaxf(y, -1.f), 1.f) * 32767.f;
      reinterpret_cast<int16_t*>(wav_out)[(long long)b * L + l] = (int16_t)__float2int_rn(c);
    } else {
      reinterpret_cast<float*>(wav_out)[(long long)b * L + l] = y;
    }
  }
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
                const float* sty, const float* emo, int B, int T, float* cond, void* cond3, cudaStream_t st) {
  long long frames = (long long)B * T;
  int grid = (int)(frames < 592 ? frames : 592);
  cond_kernel<<<grid, 128, 0, st>>>(prosody, w0, b0, w2, b2, sty, emo, B, T, cond, reinterpret_cast<uint16_t*>(cond3));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int film_launch(const float* cond, const float* w_all, const float* b_all, int M, int ncols, float* out,
                cudaStream_t st) {
  B200_CHECK_ARG(ncols % 128 == 0, "film: ncols=%d must be a multiple of 128", ncols);
  dim3 grid(ceil_div(M, 128), ncols / 128);
  film_sgemm_kernel<<<grid, 256, 0, st>>>(cond, w_all, b_all, M, ncols, out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int band_split_launch(const float* mel, const float* wt, const float* bias, int B, int channels, int band_size, int T,
                      int H, int fmt, int time_major, void* out16, cudaStream_t st) {
  dim3 grid(ceil_div(T, kSplitTT), channels / band_size, B);
  size_t smem = (size_t)band_size * (kSplitTT + 8) * sizeof(float);
  band_split_kernel<<<grid, 256, smem, st>>>(mel, wt, bias, B, channels, band_size, T, H, fmt, time_major,
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
                      int pcm16, const int* valid, void* wav, cudaStream_t st) {
  B200_CHECK_ARG(Cb == 32 && nb == 4, "band_merge: %d bands x %d channels unsupported (4 x 32)", nb, Cb);
  const long long warps = (long long)B * ceil_div(L, 32);
  const int grid = (int)((warps + 7) / 8);
  const uint16_t* x = reinterpret_cast<const uint16_t*>(x16);
  if (fmt == 0) {
    if (pcm16) band_merge_kernel<0, 1><<<grid, 256, 0, st>>>(x, w, bias, B, L, valid, wav);
    else band_merge_kernel<0, 0><<<grid, 256, 0, st>>>(x, w, bias, B, L, valid, wav);
  } else {
    if (pcm16) band_merge_kernel<1, 1><<<grid, 256, 0, st>>>(x, w, bias, B, L, valid, wav);
    else band_merge_kernel<1, 0><<<grid, 256, 0, st>>>(x, w, bias, B, L, valid, wav);
  }
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
// overflow check (debug): how many stored 16-bit activations are Inf / NaN (exponent field all ones)
__global__ void overflow_count_kernel(const uint16_t* __restrict__ x, long long n, int fmt, unsigned long long* counter) {
  const uint16_t mask = fmt == 0 ? 0x7C00 : 0x7F80;
  unsigned int c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (x[i] & mask) == mask;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, (unsigned long long)c);
}
int overflow_count_launch(const void* x16, long long n, int fmt, unsigned long long* counter, cudaStream_t st) {
  overflow_count_kernel<<<1184, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(x16), n, fmt, counter);
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


// ------------------------------------------------------------------ GlobalStyleTokens (vocoder7/gst.py:16-35)
// logits[b, n, t] = W1 relu(conv3(mel))[.., t] + b1; weights = softmax over t; style[b, :] =
// sum_n sum_t weights[b, n, t] * tokens[n, :].  (Because the softmax runs over the same axis the
// einsum sums over, sum_t weights = 1 and the reference's style is sum_n tokens[n] up to fp32
// rounding -- reproduced here as it is written, not simplified.)
// Pass 1: one CTA per (64-frame chunk, utterance): hidden = relu(conv) for its frames, the 10 token
// logits per frame, and per token the chunk's online-softmax partial (max, sum exp).  Pass 2 combines.
constexpr int kGstTT = 64;
__global__ void __launch_bounds__(128) gst_partial_kernel(const float* __restrict__ mel, int time_major, int B, int T,
                                                          int channels, int sd, int nt, const float* __restrict__ w0,
                                                          const float* __restrict__ b0, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, float* __restrict__ part) {
  extern __shared__ float sm[];
  float* s_mel = sm;                               // [channels][kGstTT + 2]
  float* s_hid = sm + channels * (kGstTT + 2);     // [kGstTT][sd + 1]
  float* s_log = s_hid + kGstTT * (sd + 1);        // [nt][kGstTT]
  constexpr int W = kGstTT + 2;
  const int t0 = blockIdx.x * kGstTT, b = blockIdx.y;
  for (int i = threadIdx.x; i < channels * W; i += blockDim.x) {   // consecutive threads -> consecutive addresses
    const int ci = time_major ? i % channels : i / W, tt = time_major ? i / channels : i % W;
    const int t = t0 + tt - 1;
    const long long src = time_major ? ((long long)b * T + t) * channels + ci : ((long long)b * channels + ci) * T + t;
    s_mel[ci * W + tt] = (t >= 0 && t < T) ? mel[src] : 0.f;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < sd; d += blockDim.x) {
    float acc[kGstTT];
    const float bv = b0[d];
#pragma unroll
    for (int i = 0; i < kGstTT; ++i) acc[i] = bv;
    for (int ci = 0; ci < channels; ++ci)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float w = __ldg(w0 + ((long long)d * channels + ci) * 3 + k);
        const float* m = s_mel + ci * W + k;
#pragma unroll
        for (int i = 0; i < kGstTT; ++i) acc[i] = fmaf(w, m[i], acc[i]);
      }
#pragma unroll
    for (int i = 0; i < kGstTT; ++i) s_hid[i * (sd + 1) + d] = fmaxf(acc[i], 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nt * kGstTT; i += blockDim.x) {
    const int n = i / kGstTT, tt = i % kGstTT;
    float a = b1[n];
    for (int d = 0; d < sd; ++d) a = fmaf(__ldg(w1 + n * sd + d), s_hid[tt * (sd + 1) + d], a);
    s_log[i] = (t0 + tt < T) ? a : -INFINITY;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < nt; n += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int tt = lane; tt < kGstTT; tt += 32) m = fmaxf(m, s_log[n * kGstTT + tt]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int tt = lane; tt < kGstTT; tt += 32) sum += expf(s_log[n * kGstTT + tt] - m);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      float* dst = part + (((long long)b * gridDim.x + blockIdx.x) * nt + n) * 2;
      dst[0] = m;
      dst[1] = sum;
    }
  }
}
__global__ void __launch_bounds__(128) gst_combine_kernel(const float* __restrict__ part, int chunks, int sd, int nt,
                                                          const float* __restrict__ tokens, float* __restrict__ style) {
  __shared__ float s_r[64];
  const int b = blockIdx.x;
  for (int n = threadIdx.x; n < nt; n += blockDim.x) {
    float M = -INFINITY;
    for (int c = 0; c < chunks; ++c) M = fmaxf(M, part[(((long long)b * chunks + c) * nt + n) * 2]);
    float Z = 0.f;
    for (int c = 0; c < chunks; ++c) {
      const float* q = part + (((long long)b * chunks + c) * nt + n) * 2;
      Z += q[1] * expf(q[0] - M);
    }
    // sum_t softmax_t: each chunk contributes (its sum of exponentials) / Z
    float r = 0.f;
    for (int c = 0; c < chunks; ++c) {
      const float* q = part + (((long long)b * chunks + c) * nt + n) * 2;
      r += q[1] * expf(q[0] - M) / Z;
    }
    s_r[n] = r;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < sd; d += blockDim.x) {
    float a = 0.f;
    for (int n = 0; n < nt; ++n) a = fmaf(s_r[n], __ldg(tokens + n * sd + d), a);
    style[(long long)b * sd + d] = a;
  }
}
long long gst_scratch_floats(int B, int T, int nt) { return (long long)B * ceil_div(T, kGstTT) * nt * 2; }
int gst_launch(const float* mel, int time_major, int B, int T, int channels, int sd, int nt, const float* w0,
               const float* b0, const float* w1, const float* b1, const float* tokens, float* scratch, float* style,
               cudaStream_t st) {
  B200_CHECK_ARG(nt >= 1 && nt <= 64 && sd >= 1 && channels >= 1, "gst: bad sizes (tokens %d, style_dim %d)", nt, sd);
  const int chunks = ceil_div(T, kGstTT);
  const size_t smem = ((size_t)channels * (kGstTT + 2) + (size_t)kGstTT * (sd + 1) + (size_t)nt * kGstTT) * sizeof(float);
  B200_CHECK_ARG(smem <= 200 * 1024, "gst: style_dim %d / channels %d too large", sd, channels);
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(gst_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[dev & 15] = true;
  }
  gst_partial_kernel<<<dim3(chunks, B), 128, smem, st>>>(mel, time_major, B, T, channels, sd, nt, w0, b0, w1, b1, scratch);
  B200_CUDA(cudaGetLastError());
  gst_combine_kernel<<<B, 128, 0, st>>>(scratch, chunks, sd, nt, tokens, style);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
