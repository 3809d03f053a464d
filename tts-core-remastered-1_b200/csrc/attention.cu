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
#include <stdlib.h>

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
constexpr int kOffV = kOffK + 3 * kK_BYTES;      // K ring: 3 tiles (S = Q K^T runs two key tiles ahead of P V)
constexpr int kOffP = kOffV + 2 * kV_BYTES;
constexpr int kOffWo = kOffP + 3 * kP_BYTES;      // P ring: 3 tiles (tile j in buffer j % 3), so a softmax group never waits for P V
constexpr int kOffBar = kOffWo + kWO_BYTES;
constexpr int kOffXm = kOffBar + 256;     // row-maximum exchange between the two column halves: [group][tile parity][half][row]
constexpr int kAttnSmem = kOffXm + 2 * 2 * 2 * 128 * 4 + 1024;
constexpr int kColS = 0;       // S[3]: 3 x 128 columns (tile j in buffer j % 3)
constexpr int kColPV = 384;    // PV[2]: 2 x 64 columns
constexpr int kAttnPolyDefault = 6;   // every 6th exponential on the FMA pipe: best of the A/B runs (profiles/r02c_attention_poly.txt)
constexpr int kColF = 0;       // out-projection accumulator: 64 columns over S[0] (all scores consumed by then)

// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error 7.5e-5 -- P is
// rounded to 16 bits (4.9e-4) right after): the MUFU unit (16 exp2 per clock and SM) is what bounds this kernel at d = 64
// (128 x 128 exponentials per 4.2 MFLOP tile = 1024 clk against 524 clk of MMAs), so every POLY-th exponential of a row
// is evaluated here instead (8 FMA-pipe / ALU instructions for 1 MUFU instruction that holds its unit for 8 clocks).
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -126.0f);                        // masked keys (-inf): 2^-126, which the 16-bit P rounds to zero
  const float t = x + 12582912.0f;              // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);        // [-0.5, 0.5]
  float q = fmaf(0.05517167f, f, 0.24261113f);
  q = fmaf(q, f, 0.69326097f);
  q = fmaf(q, f, 0.99992806f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));   // * 2^round(x): add to the exponent field
}

template <int FMT, int POLY>
__global__ void __launch_bounds__(608, 1)
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
  uint64_t* k_full = bars + 1;        // [3]
  uint64_t* k_empty = bars + 4;       // [3]
  uint64_t* v_full = bars + 7;        // [2]
  uint64_t* v_empty = bars + 9;       // [2]
  uint64_t* s_full = bars + 11;       // [3]
  uint64_t* s_empty = bars + 14;      // [3]
  uint64_t* p_full = bars + 17;       // [3]
  uint64_t* p_empty = bars + 20;      // [3]
  uint64_t* o_full = bars + 23;       // [2]
  uint64_t* o_empty = bars + 25;      // [2]
  uint64_t* n_full = bars + 27;       // [1]  normalised O staged in smem
  uint64_t* f_full = bars + 28;       // [1]  out-projection accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, seq = blockIdx.y;
  const int kt0 = (blockIdx.x / p.n_ktiles_window) * p.n_ktiles_window;   // first key tile of this query tile's window
  const int nkt = min(p.n_ktiles_window, p.L / 128 - kt0);                // (the last window of a sequence may be shorter)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmWo);
    mbar_init(q_full, 1);
    for (int b = 0; b < 3; ++b) {
      mbar_init(&k_full[b], 1); mbar_init(&k_empty[b], 1);
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 8);      // 8 = the warps of one softmax group (one arrival per warp)
      mbar_init(&p_full[b], 8); mbar_init(&p_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&v_full[b], 1); mbar_init(&v_empty[b], 1);
      mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 8);
    }
    mbar_init(n_full, 8);
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
      // K tiles (3-deep ring) and V^T tiles (2-deep) are two independent streams -- S = Q K^T runs two key tiles
      // ahead of P V -- served by one thread: non-blocking probes, whichever ring has a free slot goes next
      int jk = 0, jv = 0;
      while (jk < nkt || jv < nkt) {
        if (jk < nkt && mbar_test(&k_empty[jk % 3], ((jk / 3) & 1) ^ 1)) {
          const int b = jk % 3;
          mbar_expect_tx(&k_full[b], kK_BYTES);
          tma_load_3d(sK + b * kK_BYTES, &tmQKV, &k_full[b], kD, (kt0 + jk) * 128, seq);
          ++jk;
        }
        if (jv < nkt && mbar_test(&v_empty[jv & 1], ((jv >> 1) & 1) ^ 1)) {
          const int b = jv & 1, key0 = (kt0 + jv) * 128;
          mbar_expect_tx(&v_full[b], kV_BYTES);
          tma_load_3d(sV + b * kV_BYTES, &tmVt, &v_full[b], key0, 0, seq);
          tma_load_3d(sV + b * kV_BYTES + 64 * 128, &tmVt, &v_full[b], key0 + 64, 0, seq);
          ++jv;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer of S = Q K^T (and of the out projection).
    // Two issuing warps (this one and warp 18 for P V): with one thread issuing all 12 MMAs, 5 commits and 5 barrier waits
    // of a key tile through divergent single-lane code the ISSUER paced the kernel (~2 k clk per tile, the softmax
    // warps spent 36 % of their time waiting for scores).  Warp-uniform control flow, one elected lane issues.
    const uint32_t idesc_s = make_idesc_f16(FMT, 128);
    const uint32_t idesc_o = make_idesc_f16(FMT, 64);
    mbar_wait(q_full, 0);
    const uint64_t q_desc = make_kmajor_desc<128>(smem_u32(sQ));
    // S runs two key tiles ahead of the softmax (three S buffers): the scores of tile j+2 are issued as soon as the
    // softmax of tile j-1 has read its buffer
#pragma unroll 1
    for (int j = 0; j < nkt; ++j) {
      const int b = j % 3;
      const uint32_t ph = (j / 3) & 1;
      mbar_wait(&k_full[b], ph);
      mbar_wait(&s_empty[b], ph ^ 1);
      tc_fence_after();
      const uint64_t k_desc = make_kmajor_desc<128>(smem_u32(sK + b * kK_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColS + b * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[b]);
      }
      __syncwarp();
    }
    // tail: out = x + Wo * (O / l) + bo
    mbar_wait(n_full, 0);
    tc_fence_after();
    const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sP));
    const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sWo));
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + kColF, a_desc + 2 * k, b_desc + 2 * k, idesc_o, k != 0);
      umma_commit(f_full);
    }
    __syncwarp();
  } else if (warp == 18) {
    // ------------------------------------------------------------ MMA issuer of O_j = P_j V_j
    const uint32_t idesc_o = make_idesc_f16(FMT, 64);
#pragma unroll 1
    for (int j = 0; j < nkt; ++j) {
      const int b = j & 1, pb = j % 3;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&v_full[b], ph);
      mbar_wait(&o_empty[b], ph ^ 1);
      mbar_wait(&p_full[pb], (j / 3) & 1);
      tc_fence_after();
      const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sP + pb * kP_BYTES));
      const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sV + b * kV_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + kColPV + b * 64, a_desc + (kb * 128 * 128 >> 4) + 2 * k,
                     b_desc + (kb * 64 * 128 >> 4) + 2 * k, idesc_o, (kb | k) != 0);
        umma_commit(&o_full[b]);
        umma_commit(&p_empty[pb]);
        umma_commit(&v_empty[b]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax / accumulate warps 2..17: sixteen warps =
    // 2 tile groups x 2 column halves x 4 TMEM lane quadrants.  Group g owns the key tiles j = g, g+2, ... (always
    // S / P / PV buffer g) with its own running (max, sum, output); inside a group the two warps of a quadrant split
    // a query row's 128 key columns (half h = keys 64h .. 64h+63 = one k-block of P) and the 64 output channels
    // (d = 32h .. 32h+31), exchanging only the row maximum of their halves per tile.  With one warp per quadrant
    // the kernel was bound by the latency of ~700 dependent instructions per key tile on one warp per scheduler
    // (2.7 k clk per tile against 640 clk of MMAs and 1 k clk of MUFU.EX2); four warps per scheduler hide it.
    const int sw = warp - 2, q = warp & 3, g = (sw >> 2) & 1, h = sw >> 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* sMaxG = reinterpret_cast<float*>(smem + kOffXm) + g * 4 * 128;        // [g][tile parity][h][row]
    const int pair_bar = 3 + g * 4 + q;                                          // named barrier of this (g, q) warp pair
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto accumulate_pv = [&](int j, float alpha) {
      const int b = j & 1;
      mbar_wait(&o_full[b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(lane_addr + kColPV + b * 64 + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(v[i]));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[b]);
    };

    int jprev = -1;
    // j % 3 and j / 3 advance by rule (j grows by 2): no divisions in the loop
    int sb = g, q3 = 0;                                  // S / P buffer j % 3 and j / 3 of the current tile
    for (int j = g; j < nkt; j += 2) {
      const int b = g;                                   // == j & 1: PV buffer
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[sb], q3 & 1);
      tc_fence_after();
      // pass 1: maximum of this thread's 64 columns (scores are already in the log2 domain), then of the row
      float m_half = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kColS + sb * 128 + h * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m_half = fmaxf(m_half, __uint_as_float(v[i]));
      }
      float* sMax = sMaxG + (ph & 1) * 2 * 128;          // double buffered: the partner may still be reading the previous tile's
      sMax[h * 128 + row] = m_half;
      named_bar_sync(pair_bar, 64);
      const float m_new = fmaxf(m_run, fmaxf(m_half, sMax[(h ^ 1) * 128 + row]));
      const float alpha = ex2_approx(m_run - m_new);     // exp2(-inf) = 0 on the first tile
      // pass 2: p = exp2(s - m), partial row sum, 16-bit P into k-block h of the swizzled K-major operand tile
      mbar_wait(&p_empty[sb], (q3 & 1) ^ 1);             // P buffer = j % 3 as well
      float l_tile = 0.f;
      uint8_t* prow = sP + sb * kP_BYTES + h * 128 * 128 + row * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kColS + sb * 128 + h * 64 + c * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float xs = __uint_as_float(v[i]) - m_new;
          pv[i] = (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) ? ex2_poly3(xs) : ex2_approx(xs);
          l_tile += pv[i];
        }
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const int chunk = c * 4 + i8;                  // 8-key chunk inside this half's 64 keys
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack2t<FMT>(pv[i8 * 8 + 0], pv[i8 * 8 + 1]), pack2t<FMT>(pv[i8 * 8 + 2], pv[i8 * 8 + 3]),
                         pack2t<FMT>(pv[i8 * 8 + 4], pv[i8 * 8 + 5]), pack2t<FMT>(pv[i8 * 8 + 6], pv[i8 * 8 + 7]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[sb]);
        mbar_arrive(&s_empty[sb]);
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // fold this group's previous tile's P V product into the running output (its alpha was computed last turn)
      if (jprev >= 0) accumulate_pv(jprev, alpha_prev);
      alpha_prev = alpha;
      jprev = j;
      if (sb >= 1) { sb -= 1; q3 += 1; } else sb += 2;   // (j + 2) % 3, (j + 2) / 3
    }
    if (jprev >= 0) accumulate_pv(jprev, alpha_prev);

    // ---- merge the two groups (and the two halves' row sums) through the now idle K / V ring: every MMA has completed
    // once all softmax warps are past their last o_full
    float* xo = reinterpret_cast<float*>(sK) + row * 67 + h * 32;   // group 1's outputs [128][67] (34 KB <= 64 KB ring)
    float* xm = reinterpret_cast<float*>(sK) + 128 * 67;            // [g][row] running max
    float* xl = xm + 256;                                           // [g][h][row] partial row sums
    named_bar_sync(2, 512);
    if (g == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) xo[i] = o_acc[i];
    }
    if (h == 0) xm[g * 128 + row] = m_run;
    xl[(g * 2 + h) * 128 + row] = l_run;
    named_bar_sync(2, 512);
    if (g == 0) {                                                   // group 0 finishes the tile
      const float m1 = xm[128 + row];
      const float m = fmaxf(m_run, m1);
      const float a0 = ex2_approx(m_run - m), a1 = ex2_approx(m1 - m);     // m1 = -inf (no odd tile): a1 = 0
      const float l_tot = (xl[row] + xl[128 + row]) * a0 + (xl[256 + row] + xl[384 + row]) * a1;
      const float inv_l = 1.0f / l_tot;
      // normalise, stage as the A operand of the out projection (reuses P buffer 0: this thread's 32 channels)
      uint8_t* nrow = sP + row * 128;
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (o_acc[i8 * 8 + e] * a0 + xo[i8 * 8 + e] * a1) * inv_l;
        *reinterpret_cast<uint4*>(nrow + (((h * 4 + i8) ^ (row & 7)) << 4)) =
            make_uint4(pack2t<FMT>(y[0], y[1]), pack2t<FMT>(y[2], y[3]), pack2t<FMT>(y[4], y[5]), pack2t<FMT>(y[6], y[7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(n_full);
      mbar_wait(f_full, 0);
      tc_fence_after();
      const long long roff = ((long long)seq * p.L + q0 + row) * kD + h * 32;
      const uint4* xin = reinterpret_cast<const uint4*>(p.x16 + roff);
      uint4* dst = reinterpret_cast<uint4*>(p.out + roff);
      uint32_t v[32];
      tmem_ld32(lane_addr + kColF + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) {
        const uint4 xa = __ldg(xin + i8);
        const uint32_t xw[4] = {xa.x, xa.y, xa.z, xa.w};
        uint32_t ow[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 xs = unpack2t<FMT>(xw[e2]);
          const int ch = h * 32 + i8 * 8 + e2 * 2;
          const float y0 = xs.x + __uint_as_float(v[i8 * 8 + e2 * 2]) + __ldg(p.bo + ch);
          const float y1 = xs.y + __uint_as_float(v[i8 * 8 + e2 * 2 + 1]) + __ldg(p.bo + ch + 1);
          ow[e2] = pack2t<FMT>(y0, y1);
        }
        dst[i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
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
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  // B200VOC_ATTN_POLY = n: every n-th exponential of a row on the FMA pipe (built: 3, 4, 6, 8; anything else = all on the
  // MUFU unit; A/B switch)
  static const int poly = [] { const char* e = getenv("B200VOC_ATTN_POLY"); return e ? atoi(e) : kAttnPolyDefault; }();
  dim3 grid(L / 128, N);
  auto launch = [&](auto kern) -> int {          // one (format, poly) variant per process and device in practice
    static bool configured[16][2] = {};
    if (!configured[dev & 15][fmt & 1]) {
      B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
      configured[dev & 15][fmt & 1] = true;
    }
    kern<<<grid, 608, kAttnSmem, st>>>(tmQKV, tmVt, tmWo, p);
    return B200VOC_OK;
  };
  int rc;
  switch (poly) {
    case 3: rc = fmt == 0 ? launch(attention_kernel<0, 3>) : launch(attention_kernel<1, 3>); break;
    case 4: rc = fmt == 0 ? launch(attention_kernel<0, 4>) : launch(attention_kernel<1, 4>); break;
    case 6: rc = fmt == 0 ? launch(attention_kernel<0, 6>) : launch(attention_kernel<1, 6>); break;
    case 8: rc = fmt == 0 ? launch(attention_kernel<0, 8>) : launch(attention_kernel<1, 8>); break;
    default: rc = fmt == 0 ? launch(attention_kernel<0, 0>) : launch(attention_kernel<1, 0>); break;
  }
  B200_TRY(rc);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
