// K6: SelfAttention (generator.py:43-44,91-92; body = repair R3): residual single-head softmax
// attention over time, d = C = 64, at 128x the mel frame rate.
//
//   qkv = Conv1d(C,3C,1)(x)                 -> linear_launch (tcgen05 GEMM), q pre-scaled by
//                                              log2(e)/sqrt(C) at weight-pack time
//   o   = softmax(q k^T) v                  -> flash-style kernel below: S = Q K^T and O += P V on
//                                              tcgen05 with TMEM accumulators, online softmax in
//                                              registers, P staged through swizzled shared memory,
//                                              L x L is never materialised
//   out = x + Conv1d(C,C,1)(o)              -> fused into the same kernel's tail (one more MMA)
//
// One CTA owns 128 queries of one sequence and walks the key tiles of its attention window
// (global: all of L).  K and V^T tiles stream through a 2-deep TMA ring; S and the PV product are
// double buffered in TMEM so the MMAs of key tile j+1 overlap the softmax of tile j.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

int linear_launch(const void* x16, const void* w_packed, const float* bias, int N, int L, int Cin, int Cout, int fmt,
                  int store_lrelu, void* out16, cudaStream_t stream);

struct AttnParams {
  int L, n_ktiles_window;   // keys per window / 128 (== L/128 for global attention)
  const uint16_t* x16;      // [N, L, 64] residual input (raw)
  const float* bo;          // [64]
  uint16_t* out;            // [N, L, 64]
};

constexpr int kD = 64;
constexpr int kQ_BYTES = 128 * 128;        // Q tile 128 x 64 x 2B
constexpr int kK_BYTES = 128 * 128;        // K tile
constexpr int kV_BYTES = 2 * 64 * 128;     // V^T: 2 k-blocks of [64 d rows x 64 keys]
constexpr int kP_BYTES = 2 * 128 * 128;    // P: 2 k-blocks of [128 q rows x 64 keys]
constexpr int kWO_BYTES = 64 * 128;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQ_BYTES;
constexpr int kOffV = kOffK + 2 * kK_BYTES;
constexpr int kOffP = kOffV + 2 * kV_BYTES;
constexpr int kOffWo = kOffP + 2 * kP_BYTES;
constexpr int kOffBar = kOffWo + kWO_BYTES;
constexpr int kAttnSmem = kOffBar + 256 + 1024;
constexpr int kColS = 0;       // S[2]: 2 x 128 columns
constexpr int kColPV = 256;    // PV[2]: 2 x 64 columns
constexpr int kColF = 384;     // out-projection accumulator: 64 columns

template <int FMT>
__global__ void __launch_bounds__(320, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmVt,
                 const __grid_constant__ CUtensorMap tmWo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  uint8_t* sWo = smem + kOffWo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;            // [1]  Q + Wo landed
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2]
  uint64_t* s_empty = bars + 11;      // [2]
  uint64_t* p_full = bars + 13;       // [2]
  uint64_t* p_empty = bars + 15;      // [2]
  uint64_t* o_full = bars + 17;       // [2]
  uint64_t* o_empty = bars + 19;      // [2]
  uint64_t* n_full = bars + 21;       // [1]  normalised O staged in smem
  uint64_t* f_full = bars + 22;       // [1]  out-projection accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, seq = blockIdx.y;
  const int kt0 = (blockIdx.x / p.n_ktiles_window) * p.n_ktiles_window;   // first key tile of this query tile's window
  const int nkt = min(p.n_ktiles_window, p.L / 128 - kt0);                // (the last window of a sequence may be shorter)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmWo);
    mbar_init(q_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&k_full[b], 1); mbar_init(&k_empty[b], 1);
      mbar_init(&v_full[b], 1); mbar_init(&v_empty[b], 1);
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128);
      mbar_init(&p_full[b], 128); mbar_init(&p_empty[b], 1);
      mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 128);
    }
    mbar_init(n_full, 128);
    mbar_init(f_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, kQ_BYTES + kWO_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, 0, q0, seq);
      tma_load_2d(sWo, &tmWo, q_full, 0, 0);
      for (int j = 0; j < nkt; ++j) {
        const int b = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int key0 = (kt0 + j) * 128;
        mbar_wait(&k_empty[b], ph ^ 1);
        mbar_expect_tx(&k_full[b], kK_BYTES);
        tma_load_3d(sK + b * kK_BYTES, &tmQKV, &k_full[b], kD, key0, seq);
        mbar_wait(&v_empty[b], ph ^ 1);
        mbar_expect_tx(&v_full[b], kV_BYTES);
        tma_load_3d(sV + b * kV_BYTES, &tmVt, &v_full[b], key0, 0, seq);
        tma_load_3d(sV + b * kV_BYTES + 64 * 128, &tmVt, &v_full[b], key0 + 64, 0, seq);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(FMT, 128);
      const uint32_t idesc_o = make_idesc_f16(FMT, 64);
      mbar_wait(q_full, 0);
      const uint64_t q_desc = make_kmajor_desc<128>(smem_u32(sQ));
      auto issue_pv = [&](int j) {
        const int b = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&p_full[b], ph);
        mbar_wait(&v_full[b], ph);
        mbar_wait(&o_empty[b], ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sP + b * kP_BYTES + kb * 128 * 128));
          const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sV + b * kV_BYTES + kb * 64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + kColPV + b * 64, a_desc + 2 * k, b_desc + 2 * k, idesc_o, (kb | k) != 0);
        }
        umma_commit(&o_full[b]);
        umma_commit(&p_empty[b]);
        umma_commit(&v_empty[b]);
      };
      for (int j = 0; j < nkt; ++j) {
        const int b = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[b], ph);
        mbar_wait(&s_empty[b], ph ^ 1);
        tc_fence_after();
        const uint64_t k_desc = make_kmajor_desc<128>(smem_u32(sK + b * kK_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColS + b * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[b]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(nkt - 1);
      // tail: out = x + Wo * (O / l) + bo
      mbar_wait(n_full, 0);
      tc_fence_after();
      const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sP));
      const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sWo));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColF, a_desc + 2 * k, b_desc + 2 * k, idesc_o, k != 0);
      umma_commit(f_full);
    }
  } else {
    // ------------------------------------------------------------ softmax / accumulate warps 2..9: TWO groups of four
    // warps (one thread per query row each).  Group g owns the key tiles j = g, g+2, ... -- i.e. always S / P / PV
    // buffer g -- and keeps its own running (max, sum, output): with one group the kernel was bound by the latency of
    // one warp per scheduler walking ~700 dependent instructions per key tile (2 k clk against 640 clk of MMAs); two
    // groups overlap the softmax of tile j+1 with that of tile j.  The two partial results are merged once at the
    // end (the usual rescale by 2^(m_g - m)).
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float o_acc[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto accumulate_pv = [&](int j, float alpha) {
      const int b = j & 1;
      mbar_wait(&o_full[b], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kColPV + b * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(&o_empty[b]);
    };

    int jprev = -1;
    for (int j = g; j < nkt; j += 2) {
      const int b = g;                                   // == j & 1
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[b], ph);
      tc_fence_after();
      // pass 1: row maximum (scores are already in the log2 domain)
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kColS + b * 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = ex2_approx(m_run - m_new);     // exp2(-inf) = 0 on the first tile
      // pass 2: p = exp2(s - m), row sum, 16-bit P into the swizzled K-major operand tile
      mbar_wait(&p_empty[b], ph ^ 1);
      float l_tile = 0.f;
      uint8_t* prow = sP + b * kP_BYTES + row * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kColS + b * 128 + c * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = ex2_approx(__uint_as_float(v[i]) - m_new);
          l_tile += pv[i];
        }
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const int col = c * 32 + i8 * 8;              // key column inside the 128-key tile
          const int kb = col >> 6, chunk = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(prow + kb * 128 * 128 + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack2t<FMT>(pv[i8 * 8 + 0], pv[i8 * 8 + 1]), pack2t<FMT>(pv[i8 * 8 + 2], pv[i8 * 8 + 3]),
                         pack2t<FMT>(pv[i8 * 8 + 4], pv[i8 * 8 + 5]), pack2t<FMT>(pv[i8 * 8 + 6], pv[i8 * 8 + 7]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[b]);
      mbar_arrive(&s_empty[b]);
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // fold this group's previous tile's P V product into the running output (its alpha was computed last turn)
      if (jprev >= 0) accumulate_pv(jprev, alpha_prev);
      alpha_prev = alpha;
      jprev = j;
    }
    if (jprev >= 0) accumulate_pv(jprev, alpha_prev);

    // ---- merge the two groups' partial results through the (now idle) K / V ring: every MMA has completed once both
    // groups are past their last o_full
    float* xch = reinterpret_cast<float*>(sK) + row * 67;         // [128][67]: o[64], m, l (34 KB <= the 64 KB ring)
    named_bar_sync(2, 256);
    if (g == 1) {
#pragma unroll
      for (int i = 0; i < kD; ++i) xch[i] = o_acc[i];
      xch[64] = m_run;
      xch[65] = l_run;
    }
    named_bar_sync(2, 256);
    if (g == 0) {                                                 // group 0 finishes the tile
    {
      const float m1 = xch[64], l1 = xch[65];
      const float m = fmaxf(m_run, m1);
      const float a0 = ex2_approx(m_run - m), a1 = ex2_approx(m1 - m);   // m1 = -inf (no odd tile): a1 = 0
      l_run = l_run * a0 + l1 * a1;
#pragma unroll
      for (int i = 0; i < kD; ++i) o_acc[i] = o_acc[i] * a0 + xch[i] * a1;
    }

    // ---- tail: normalise, stage as the A operand of the out projection (reuses P buffer 0)
    // (o_full of the last tile was committed after every earlier MMA, so both P buffers are free here)
    const float inv_l = 1.0f / l_run;
    uint8_t* nrow = sP + row * 128;
#pragma unroll
    for (int i8 = 0; i8 < 8; ++i8) {
      *reinterpret_cast<uint4*>(nrow + ((i8 ^ (row & 7)) << 4)) = make_uint4(
          pack2t<FMT>(o_acc[i8 * 8 + 0] * inv_l, o_acc[i8 * 8 + 1] * inv_l),
          pack2t<FMT>(o_acc[i8 * 8 + 2] * inv_l, o_acc[i8 * 8 + 3] * inv_l),
          pack2t<FMT>(o_acc[i8 * 8 + 4] * inv_l, o_acc[i8 * 8 + 5] * inv_l),
          pack2t<FMT>(o_acc[i8 * 8 + 6] * inv_l, o_acc[i8 * 8 + 7] * inv_l));
    }
    fence_proxy_async_smem();
    mbar_arrive(n_full);
    mbar_wait(f_full, 0);
    tc_fence_after();
    const long long roff = ((long long)seq * p.L + q0 + row) * kD;
    const uint4* xin = reinterpret_cast<const uint4*>(p.x16 + roff);
    uint4* dst = reinterpret_cast<uint4*>(p.out + roff);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(lane_addr + kColF + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {
        const uint4 xa = __ldg(xin + c * 4 + i8);
        const uint32_t xw[4] = {xa.x, xa.y, xa.z, xa.w};
        uint32_t ow[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 xs = unpack2t<FMT>(xw[e2]);
          const int ch = c * 32 + i8 * 8 + e2 * 2;
          const float y0 = xs.x + __uint_as_float(v[i8 * 8 + e2 * 2]) + __ldg(p.bo + ch);
          const float y1 = xs.y + __uint_as_float(v[i8 * 8 + e2 * 2 + 1]) + __ldg(p.bo + ch + 1);
          ow[e2] = pack2t<FMT>(y0, y1);
        }
        dst[c * 4 + i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// v part of qkv16[N, L, 192] (columns 128..191) -> vt16[N, 64, L]
__global__ void __launch_bounds__(256) transpose_v_kernel(const uint16_t* __restrict__ qkv, int L,
                                                          uint16_t* __restrict__ vt) {
  __shared__ uint16_t tile[64][66];
  const int n = blockIdx.y, l0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;           // r: time, c: channel
    tile[r][c] = (l0 + r < L) ? qkv[((long long)n * L + l0 + r) * 192 + 128 + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, r = i & 63;           // c: channel (row of vt), r: time (contiguous)
    if (l0 + r < L) vt[((long long)n * 64 + c) * L + l0 + r] = tile[r][c];
  }
}

// scratch: qkv16 [N*L*192] + vt16 [N*64*L]
long long attention_scratch_elems(int N, int L, int C) { return (long long)N * L * (3 * C) + (long long)N * C * L; }

int attention_launch(const void* x16, const void* wqkv, const float* bqkv, const void* wo, const float* bo, int N,
                     int L, int C, int window, int fmt, void* scratch, void* out16, cudaStream_t st) {
  B200_CHECK_ARG(C == 64, "attention: channels %d unsupported (64)", C);
  B200_CHECK_ARG(L % 128 == 0, "attention: L=%d must be a multiple of 128", L);
  int win = (window <= 0 || window >= L) ? L : window;
  B200_CHECK_ARG(win % 128 == 0, "attention: window %d must be a multiple of 128", win);
  uint16_t* qkv = reinterpret_cast<uint16_t*>(scratch);
  uint16_t* vt = qkv + (long long)N * L * 192;
  // 1) fused q|k|v projection (q rows of wqkv / bqkv carry log2(e)/sqrt(C))
  B200_TRY(linear_launch(x16, wqkv, bqkv, N, L, C, 3 * C, fmt, 0, qkv, st));
  // 2) V^T (keys contiguous) for the K-major B operand of P V
  transpose_v_kernel<<<dim3(ceil_div(L, 64), N), 256, 0, st>>>(qkv, L, vt);
  B200_CUDA(cudaGetLastError());
  // 3) flash attention + out projection + residual
  CUtensorMap tmQKV, tmVt, tmWo;
  B200_TRY(make_tmap_3d(&tmQKV, qkv, 192, L, N, 192 * 2, (uint64_t)L * 192 * 2, 64, 128, 128));
  B200_TRY(make_tmap_3d(&tmVt, vt, L, 64, N, (uint64_t)L * 2, (uint64_t)64 * L * 2, 64, 64, 128));
  B200_TRY(make_tmap_2d(&tmWo, wo, 64, 64, 128, 64, 64, 128));
  AttnParams p{};
  p.L = L;
  p.n_ktiles_window = win / 128;
  p.x16 = reinterpret_cast<const uint16_t*>(x16);
  p.bo = bo;
  p.out = reinterpret_cast<uint16_t*>(out16);
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(attention_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    B200_CUDA(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured[dev & 15] = true;
  }
  dim3 grid(L / 128, N);
  if (fmt == 0) attention_kernel<0><<<grid, 320, kAttnSmem, st>>>(tmQKV, tmVt, tmWo, p);
  else attention_kernel<1><<<grid, 320, kAttnSmem, st>>>(tmQKV, tmVt, tmWo, p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
