// Critic convolutions on the tensor cores (forward half): the stride-1 layers of
// MultiScaleDiscriminator (vocoder7/discriminators.py:71-92: Conv1d(64 -> 256, k) and Conv1d(256 -> 1024, k),
// k = 15 / 41) hold 97 % of the critics' FLOPs and are GEMM-shaped.  Same construction as the Generator's
// convolutions (conv_gemm.cu / resblock3.cu):
//   * time on the GEMM M axis (128 output positions per tile), output channels on N (128 per tile), K = taps x Cin;
//   * the input map is repacked channels-last, 16-bit; ONE TMA load brings the 128 + k - 1 rows a tile needs for a
//     64-channel block, and every tap is the same tile behind a row-shifted UMMA descriptor (zero rows outside
//     [0, L) come from the TMA unit = the convolution's zero padding);
//   * the maps these layers produce are returned features compared at fp32 tolerances and spectral-norm weights of a
//     fresh critic are large (W / sigma, sigma ~ 1e-4 .. 1e-2), so operands are SPLIT bf16: x = hi + lo (16
//     significant bits, bf16 range), y = x_hi w_hi + x_lo w_hi + x_hi w_lo, three MMAs per k-step with fp32
//     accumulation in TMEM -- measured 4e-6 relative on the layer, inside the 2e-4 feature tolerance;
//   * weights stream from L2 through a 4-stage TMA ring (hi and lo tile per stage), accumulators are double buffered
//     in TMEM, the epilogue adds the bias and writes BOTH returned maps (conv, LeakyReLU(conv)) as fp32 [B, C, L],
//     coalesced along time (lane = output position).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int kDgBN = 128;                       // output channels per tile
constexpr int kDgStages = 4;                     // weight ring
constexpr int kDgARowsMax = 168;                 // 128 + 41 - 1 (k <= 41), multiple of 8
constexpr int kDgAPlaneMax = kDgARowsMax * 128;  // one 64-channel block of one plane (hi or lo)
constexpr int kDgABuf = 2 * kDgAPlaneMax;        // hi + lo
constexpr int kDgBTile = kDgBN * 128;
constexpr int kDgBStage = 2 * kDgBTile;          // hi + lo
constexpr int kDgSmem = 2 * kDgABuf + kDgStages * kDgBStage + 256 + 1024;
static_assert(kDgSmem <= 227 * 1024, "shared memory budget");

struct DiscGemmParams {
  int B, L, Lout, K, pad, Cin, Cout;
  int a_rows;                 // rows per A tile: 128 + K - 1 rounded up to 8
  int m_tiles, n_tiles;
  const float* bias;
  float slope;
  float* y_pre;               // [B, Cout, Lout] conv + bias (may be null)
  float* y_act;               // [B, Cout, Lout] LeakyReLU (may be null)
};

__global__ void __launch_bounds__(192, 1)
disc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const DiscGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kDgABuf;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kDgStages * kDgBStage);
  uint64_t* a_full = bars;                    // [2]
  uint64_t* a_empty = a_full + 2;             // [2]
  uint64_t* b_full = a_empty + 2;             // [stages]
  uint64_t* b_empty = b_full + kDgStages;     // [stages]
  uint64_t* acc_full = b_empty + kDgStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin >> 6;
  const int total_tiles = p.m_tiles * p.n_tiles * p.B;
  const int a_plane = p.a_rows * 128;
  const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int G = my_tiles * KB;                // (tile, 64-channel block) units this CTA walks
  auto decode = [&](int t, int& b, int& m0, int& n0) {       // column tile fastest: neighbours share the input rows
    const int nt = t % p.n_tiles, rest = t / p.n_tiles;
    n0 = nt * kDgBN;
    b = rest / p.m_tiles;
    m0 = (rest - b * p.m_tiles) * 128;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kDgStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kDgBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: input tiles + weight ring
    if (lane == 0) {
      auto issue_a = [&](int g) {
        const int t = (int)blockIdx.x + (g / KB) * (int)gridDim.x, kb = g % KB, ab = g & 1;
        int b, m0, n0;
        decode(t, b, m0, n0);
        mbar_wait(&a_empty[ab], ((g >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[ab], 2 * a_plane);
        tma_load_3d(sA + ab * kDgABuf, &tmA, &a_full[ab], kb * 64, m0 - p.pad, b * 2);
        tma_load_3d(sA + ab * kDgABuf + a_plane, &tmA, &a_full[ab], kb * 64, m0 - p.pad, b * 2 + 1);
      };
      const int prefetch_tap = p.K > 4 ? 4 : p.K - 1;     // the next unit's input tile is requested a few taps in
      if (G > 0) issue_a(0);
      int gs = 0;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const int t = (int)blockIdx.x + (g / KB) * (int)gridDim.x, kb = g % KB;
        int b, m0, n0;
        decode(t, b, m0, n0);
#pragma unroll 1
        for (int tap = 0; tap < p.K; ++tap, ++gs) {
          if (tap == prefetch_tap && g + 1 < G) issue_a(g + 1);
          const int s = gs % kDgStages;
          mbar_wait(&b_empty[s], ((gs / kDgStages) & 1) ^ 1);
          mbar_expect_tx(&b_full[s], kDgBStage);
          uint8_t* st = sB + s * kDgBStage;
          tma_load_2d(st, &tmB, &b_full[s], tap * p.Cin + kb * 64, n0);                    // W_hi
          tma_load_2d(st + kDgBTile, &tmB, &b_full[s], tap * p.Cin + kb * 64, p.Cout + n0);  // W_lo
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control flow, one elected lane)
    const uint32_t idesc = make_idesc_f16(1 /*bf16*/, kDgBN);
    int gs = 0;
    bool ready = false;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const int kb = g % KB, ti = g / KB, buf = ti & 1, ab = g & 1;
      if (kb == 0) {
        mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);
      }
      mbar_wait(&a_full[ab], (g >> 1) & 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + ab * kDgABuf);
#pragma unroll 1
      for (int tap = 0; tap < p.K; ++tap, ++gs) {
        const int s = gs % kDgStages;
        if (!ready) mbar_wait(&b_full[s], (gs / kDgStages) & 1);
        tc_fence_after();
        ready = mbar_test(&b_full[(gs + 1) % kDgStages], ((gs + 1) / kDgStages) & 1);
        const uint64_t a_hi = make_kmajor_desc<128>(a_base + tap * 128);              // row-shifted: tap t reads rows t ..
        const uint64_t a_lo = make_kmajor_desc<128>(a_base + a_plane + tap * 128);
        const uint64_t b_hi = make_kmajor_desc<128>(smem_u32(sB + s * kDgBStage));
        const uint64_t b_lo = make_kmajor_desc<128>(smem_u32(sB + s * kDgBStage + kDgBTile));
        const uint32_t d = tmem_base + buf * kDgBN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | tap | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          umma_commit(&b_empty[s]);
          if (tap == p.K - 1) {
            umma_commit(&a_empty[ab]);
            if (kb == KB - 1) umma_commit(&acc_full[buf]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5: TMEM lane quadrant = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int t = (int)blockIdx.x + ti * (int)gridDim.x, buf = ti & 1;
      int b, m0, n0;
      decode(t, b, m0, n0);
      const int l = m0 + row;
      const bool valid = l < p.Lout;
      mbar_wait(&acc_full[buf], (ti >> 1) & 1);
      tc_fence_after();
      const long long o0 = ((long long)b * p.Cout + n0) * p.Lout + (valid ? l : 0);
#pragma unroll 1
      for (int c = 0; c < kDgBN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * kDgBN + c * 32, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float y = __uint_as_float(v[4 * j + e]) + bv[e];
            const long long o = o0 + (long long)(c * 32 + 4 * j + e) * p.Lout;
            if (valid) {
              if (p.y_pre) p.y_pre[o] = y;
              if (p.y_act) p.y_act[o] = y > 0.f ? y : p.slope * y;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kDgBN);
}

// ------------------------------------------------------------------ packing
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
}

// x fp32 [B][C][L] -> out bf16 [B][2 (hi, lo)][L][C]: 64 positions x 64 channels per block through a padded tile
__global__ void __launch_bounds__(256) disc_pack_act_kernel(const float* __restrict__ x, int C, int L,
                                                            uint16_t* __restrict__ out) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, l0 = blockIdx.x * 64;
  const float* xb = x + ((long long)b * C + c0) * L;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int cc = i >> 6, ll = i & 63;
    tile[cc][ll] = (l0 + ll < L) ? __ldg(xb + (long long)cc * L + l0 + ll) : 0.f;
  }
  __syncthreads();
  uint32_t* ohi = reinterpret_cast<uint32_t*>(out + ((long long)(b * 2) * L) * C);
  uint32_t* olo = reinterpret_cast<uint32_t*>(out + ((long long)(b * 2 + 1) * L) * C);
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int ll = i >> 5, cp = i & 31;          // channel pair
    if (l0 + ll >= L) continue;
    uint16_t h0, l0w, h1, l1w;
    split_bf16(tile[2 * cp][ll], h0, l0w);
    split_bf16(tile[2 * cp + 1][ll], h1, l1w);
    const long long o = ((long long)(l0 + ll) * C + c0) / 2 + cp;
    ohi[o] = (uint32_t)h0 | ((uint32_t)h1 << 16);
    olo[o] = (uint32_t)l0w | ((uint32_t)l1w << 16);
  }
}

// w fp32 [Cout][Cin][K] -> out bf16 [2 (hi, lo)][Cout][K * Cin], k index = tap * Cin + ci
__global__ void __launch_bounds__(256) disc_pack_w_kernel(const float* __restrict__ w, int Cout, int Cin, int K,
                                                          uint16_t* __restrict__ out) {
  const long long n = (long long)Cout * Cin * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int tap = (int)(r % K), co = (int)(r / K);
    uint16_t hi, lo;
    split_bf16(__ldg(w + ((long long)co * Cin + ci) * K + tap), hi, lo);
    out[i] = hi;
    out[n + i] = lo;
  }
}

static int dg_num_sms() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!n[dev & 15]) cudaDeviceGetAttribute(&n[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  return n[dev & 15];
}

int disc_gemm_launch(const float* x, const void* w_split, const float* bias, int B, int Cin, int Cout, int L, int K,
                     int pad, float slope, float* y_pre, float* y_act, void* workspace, cudaStream_t st) {
  const int Lout = L + 2 * pad - K + 1;
  uint16_t* x16 = reinterpret_cast<uint16_t*>(workspace);
  {
    dim3 grid((unsigned)ceil_div(L, 64), (unsigned)(Cin / 64), (unsigned)B);
    disc_pack_act_kernel<<<grid, 256, 0, st>>>(x, Cin, L, x16);
    B200_CUDA(cudaGetLastError());
  }
  DiscGemmParams p{};
  p.B = B; p.L = L; p.Lout = Lout; p.K = K; p.pad = pad; p.Cin = Cin; p.Cout = Cout;
  p.a_rows = (128 + K - 1 + 7) & ~7;
  p.m_tiles = ceil_div(Lout, 128);
  p.n_tiles = Cout / kDgBN;
  p.bias = bias; p.slope = slope; p.y_pre = y_pre; p.y_act = y_act;
  CUtensorMap tmA, tmB;
  B200_TRY(make_tmap_3d(&tmA, x16, Cin, L, 2ull * B, (uint64_t)Cin * 2, (uint64_t)L * Cin * 2, 64, p.a_rows, 128));
  B200_TRY(make_tmap_2d(&tmB, w_split, (uint64_t)K * Cin, 2ull * Cout, (uint64_t)K * Cin * 2, 64, kDgBN, 128));
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(disc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDgSmem));
    configured[dev & 15] = true;
  }
  const long long total = (long long)p.m_tiles * p.n_tiles * B;
  const int sms = dg_num_sms();
  const int grid = (int)(total < sms ? total : sms);
  disc_gemm_kernel<<<grid, 192, kDgSmem, st>>>(tmA, tmB, p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int disc_pack_w_launch(const float* w, int Cout, int Cin, int K, void* out, cudaStream_t st) {
  const long long n = (long long)Cout * Cin * K;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  disc_pack_w_kernel<<<blocks, 256, 0, st>>>(w, Cout, Cin, K, reinterpret_cast<uint16_t*>(out));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (include/b200voc.h)
extern "C" {

int b200voc_disc_conv_tc_supported(int Cin, int Cout, int K, int stride, int P) {
  return stride == 1 && P == 1 && Cin >= 64 && Cin % 64 == 0 && Cout % b200::kDgBN == 0 && K >= 1 && K <= 41 && (K & 1);
}

int64_t b200voc_disc_conv_tc_workspace_bytes(int B, int Cin, int L) { return (int64_t)B * 2 * L * Cin * 2; }

int64_t b200voc_disc_split_weight_elems(int Cout, int Cin, int K) { return 2ll * Cout * Cin * K; }

int b200voc_disc_pack_weight_split(const float* w, int Cout, int Cin, int K, void* out, void* stream) {
  B200_CHECK_ARG(w && out, "disc_pack_weight_split: null argument");
  B200_CHECK_ARG(Cout > 0 && Cin > 0 && K > 0, "disc_pack_weight_split: bad shape (%d, %d, %d)", Cout, Cin, K);
  return b200::disc_pack_w_launch(w, Cout, Cin, K, out, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_disc_conv_tc(const float* x, const void* w_split, const float* bias, int B, int Cin, int Cout, int L, int K,
                         int pad, float slope, float* y_pre, float* y_act, void* workspace, int64_t workspace_bytes,
                         void* stream) {
  B200_CHECK_ARG(x && w_split && bias && workspace && (y_pre || y_act), "disc_conv_tc: null argument");
  B200_CHECK_ARG(b200voc_disc_conv_tc_supported(Cin, Cout, K, 1, 1),
                 "disc_conv_tc: Cin=%d (multiple of 64), Cout=%d (multiple of 128), odd K=%d <= 41 required", Cin, Cout, K);
  B200_CHECK_ARG(B > 0 && L > 0 && pad >= 0 && pad < K && L + 2 * pad - K + 1 > 0, "disc_conv_tc: bad shape (B=%d L=%d pad=%d)", B, L, pad);
  B200_CHECK_ARG(workspace_bytes >= b200voc_disc_conv_tc_workspace_bytes(B, Cin, L), "disc_conv_tc: workspace too small");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "disc_conv_tc: workspace must be 16-byte aligned");
  return b200::disc_gemm_launch(x, w_split, bias, B, Cin, Cout, L, K, pad, slope, y_pre, y_act, workspace,
                                reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
