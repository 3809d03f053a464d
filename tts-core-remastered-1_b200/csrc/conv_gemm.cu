// K1: multi-tap implicit GEMM on tcgen05/TMEM fed by TMA -- the transposed-convolution stages
// (generator.py:35-38,87) and every plain [rows x Cin] x [Cin x Cout] projection.
//
// GEMM view (time on the M axis, channels-last activations x16[N, L, Cin]):
//   D[m, c] = sum_tap sum_ci  X[n, m + shift_tap, ci] * W[c, tap*Cin + ci]
// A tiles (128 rows x 64 ch) are TMA loads from a 3-D tensor map (ci, l, n): a tap is the same
// tile with the l coordinate shifted, rows outside [0, L) are zero-filled by TMA, which is exactly
// the convolution's zero padding.  B tiles (BN rows x 64) come from the packed weight matrix.
// Accumulators live in TMEM; the epilogue reads them with tcgen05.ld, adds the bias, optionally
// applies leaky-ReLU, rounds to the 16-bit storage format and writes channels-last.
//
// ConvTranspose1d(Cin, Cout, k=2s, stride=s, pad=s/2) is the polyphase GEMM
//   y[m*s + r - p, co] = sum_ci x[m, ci] W[ci, co, r] + x[m-1, ci] W[ci, co, r+s],  m in [0, Lin]
// i.e. 2 taps (shift 0, -1), GEMM column c = r*Cout + co, and because the output is channels-last
// the GEMM output row m IS the contiguous run out[(m*s - p)*Cout ...] -- the pixel shuffle is free.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct GemmTapsParams {
  int rows_per_seq;   // GEMM rows per sequence (Lin + 1 for ConvT)
  int n_taps;
  int tap_shift[4];
  int k_per_tap;      // Cin (multiple of 64)
  int n_total;        // GEMM columns (s * Cout)
  int cout;           // bias period
  int stride;         // s  (1 for a plain projection)
  int pad;            // p = s/2 (0 for a plain projection)
  int fmt;            // B200VOC_FMT_*
  int store_lrelu;
  const float* bias;
  int dbg;            // debug switch (B200VOC_DBG): 4 = skip the TMA stores
  int m_tiles, n_tiles, n_seq;   // tile grid walked by the persistent CTAs
};

constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 x 2B

template <int BN, int STAGES, int FMT, bool LRELU>
__global__ void __launch_bounds__(192, 1)
gemm_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const GemmTapsParams p) {
  // Persistent: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (column tile fastest,
  // so CTAs working on the same input rows run together and share them through L2).  The operand
  // ring runs continuously across tiles; the accumulator is double buffered in TMEM so the
  // epilogue of tile i (TMEM -> registers -> swizzled smem staging -> TMA store) overlaps the
  // MMAs of tile i+1.
  constexpr int B_TILE = BN * 128;
  constexpr int STAGE_BYTES = kATileBytes + B_TILE;
  constexpr int STAGING = 128 * BN * 2;
  constexpr int NSTG = BN <= 128 ? 2 : 1;     // output staging buffers: the TMA store of tile i-1 drains while tile i is converted
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "tile width");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* staging0 = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging0 + NSTG * STAGING);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kpt = p.k_per_tap >> 6;
  const int num_k = p.n_taps * kpt;
  const int total_tiles = p.m_tiles * p.n_tiles * p.n_seq;

  // tile -> (m0, n0, seq).  Column tiles whose output phase r < p map to the previous input row
  // (l = (m-1)*s + r - p + s): they run over GEMM rows m = 1 + 128*i so that the TMA store
  // coordinate (m - 1) is never negative (row m = 0 of those phases is entirely out of range
  // anyway).  The launcher keeps tiles phase-pure.
  auto decode = [&](int t, int& m0, int& n0, int& seq, bool& part_b) {
    const int nt = t % p.n_tiles, rest = t / p.n_tiles;
    n0 = nt * BN;
    seq = rest / p.m_tiles;
    part_b = (n0 / p.cout) < p.pad;
    m0 = (rest - seq * p.m_tiles) * 128 + (part_b ? 1 : 0);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, seq;
        bool part_b;
        decode(t, m0, n0, seq, part_b);
        for (int kb = 0; kb < num_k; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          const int tap = kb / kpt, kk = kb - tap * kpt;
          uint8_t* st = smem + s * STAGE_BYTES;
          tma_load_3d(st, &tmA, &full[s], kk * 64, m0 + p.tap_shift[tap], seq);
          tma_load_2d(st + kATileBytes, &tmB, &full[s], tap * p.k_per_tap + kk * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the uniform control flow; one elected lane issues the MMAs (descriptors in
    // uniform registers); the probe of the next stage overlaps the issue of the current one
    const uint32_t idesc = make_idesc_f16(FMT, BN);
    bool ready = false;
    int g = 0, i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < num_k; ++kb, ++g) {
        const int s = g % STAGES;
        if (!ready) mbar_wait(&full[s], (g / STAGES) & 1);
        tc_fence_after();
        ready = mbar_test(&full[(g + 1) % STAGES], ((g + 1) / STAGES) & 1);
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t a_desc = make_kmajor_desc<128>(a_addr);
        const uint64_t b_desc = make_kmajor_desc<128>(a_addr + kATileBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // UMMA_K = 16 elements = 32 bytes = +2 in the >>4 start field
            umma_f16(tmem_base + buf * BN, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
          if (kb == num_k - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quadrant (warp % 4).  The accumulator tile is converted in
    // registers (+bias, leaky-ReLU, 16-bit) and staged in shared memory as [128 rows x 64 ch]
    // swizzled blocks -- one block per (output phase, 64-channel group) -- which leave through TMA
    // stores on a 4-D view (co, l mod s, l / s, n) of the output: the polyphase pixel shuffle, the -p
    // offset and both sequence ends are handled by the TMA unit.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool narrow = p.cout < 64;                 // Cout = 32: 64-byte rows, SWIZZLE_64B, one block per phase
    const int blk_bytes = narrow ? 128 * 64 : 128 * 128;
    const bool issuer = warp == 2 && lane == 0;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      int m0, n0, seq;
      bool part_b;
      decode(t, m0, n0, seq, part_b);
      const int buf = i & 1;
      mbar_wait(&acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
      uint8_t* staging = staging0 + (i % NSTG) * STAGING;
      if (i >= NSTG) {                               // this staging buffer is free once its previous stores have read it
        if (issuer) {
          if (NSTG == 2) tma_store_wait_read_1(); else tma_store_wait_read();
        }
        named_bar_sync(1, 128);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + (col0 % p.cout));
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          float y0 = __uint_as_float(v[4 * j + 0]) + bb.x, y1 = __uint_as_float(v[4 * j + 1]) + bb.y;
          float y2 = __uint_as_float(v[4 * j + 2]) + bb.z, y3 = __uint_as_float(v[4 * j + 3]) + bb.w;
          if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); y2 = lrelu_fast(y2); y3 = lrelu_fast(y3); }
          w[2 * j] = pack2t<FMT>(y0, y1);
          w[2 * j + 1] = pack2t<FMT>(y2, y3);
        }
        if (narrow) {
          uint8_t* dst = staging + c * blk_bytes + row * 64;        // block c = one phase of 32 channels
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + ((j ^ ((row >> 1) & 3)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        } else {
          uint8_t* dst = staging + (c >> 1) * blk_bytes + row * 128;  // block = 64 channels, this chunk = half of it
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + j) ^ (row & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);                  // accumulator drained: the MMAs of tile i+2 may start
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer && !(p.dbg & 4)) {
        const int bw = narrow ? 32 : 64;             // GEMM columns per block
        for (int j = 0; j < BN / bw; ++j) {
          const int col = n0 + j * bw;
          const int r = col / p.cout, co0 = col % p.cout;
          const int rr = part_b ? r - p.pad + p.stride : r - p.pad;
          const int mm = part_b ? m0 - 1 : m0;
          tma_store_4d(&tmOut, staging + j * blk_bytes, co0, rr, mm, seq);
        }
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- CTA-pair variant, BN = 256
// The single-CTA kernel at BN = 256 loads 48 KB (A 16 + B 32) per 512 clk of MMAs: ~11-13 TB/s of L2 -> shared-memory
// traffic chip-wide (ncu l1tex__m_xbar2l1tex_read_bytes), which is what bounds it (tensor pipe 51-63 %).  Here two
// CTAs (a cluster of 2) own two neighbouring 128-row tiles of the same column tile and the leader issues ONE
// tcgen05.mma.cta_group::2 (M = 256, N = 256) per k-step: each CTA loads its own A tile and only HALF of the weight
// tile (32 KB per stage instead of 48, and the ring gets a fourth stage).  Both CTAs' loads signal the leader's `full`
// barrier (cp.async.bulk.tensor ... cta_group::2), commits are multicast to both CTAs' `empty` / `acc_full` barriers,
// the peer's epilogue warps arrive on the leader's `acc_empty` through mapa.
template <int STAGES, int FMT, bool LRELU>
__global__ void __launch_bounds__(192, 1)
gemm_taps_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const GemmTapsParams p) {
  constexpr int BN = 256;
  constexpr int B_HALF = 128 * 128;                  // this CTA's 128 of the 256 weight rows x 64 k
  constexpr int STAGE_BYTES = kATileBytes + B_HALF;
  constexpr int STAGING = 128 * BN * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + STAGING);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]  (the leader's copy is the one waited on)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int kpt = p.k_per_tap >> 6;
  const int num_k = p.n_taps * kpt;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_tiles = m_pairs * p.n_tiles * p.n_seq;
  const int pair0 = (int)(blockIdx.x >> 1), pair_stride = (int)(gridDim.x >> 1);
  // pair tile -> (m0 of THIS CTA's row tile, n0, seq); a row tile past the end is all out of bounds (TMA zero-fills the
  // loads and clips the stores)
  auto decode = [&](int t, int& m0, int& n0, int& seq, bool& part_b) {
    const int nt = t % p.n_tiles, rest = t / p.n_tiles;
    n0 = nt * BN;
    seq = rest / m_pairs;
    part_b = (n0 / p.cout) < p.pad;
    m0 = (2 * (rest - seq * m_pairs) + (int)rank) * 128 + (part_b ? 1 : 0);
  };
  auto arrive_leader = [&](uint64_t* bar) {
    if (rank == 0) mbar_arrive(bar); else mbar_arrive_remote(mapa_u32(bar, 0));
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * 4);          // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * BN);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int t = pair0; t < total_tiles; t += pair_stride) {
        int m0, n0, seq;
        bool part_b;
        decode(t, m0, n0, seq, part_b);
        for (int kb = 0; kb < num_k; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
          const int tap = kb / kpt, kk = kb - tap * kpt;
          uint8_t* st = smem + s * STAGE_BYTES;
          tma_load_3d_2cta(st, &tmA, &full[s], kk * 64, m0 + p.tap_shift[tap], seq);
          tma_load_2d_2cta(st + kATileBytes, &tmB, &full[s], tap * p.k_per_tap + kk * 64, n0 + (int)rank * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_f16_m(FMT, 256, BN);
      bool ready = false;
      int g = 0, i = 0;
      for (int t = pair0; t < total_tiles; t += pair_stride, ++i) {
        const int buf = i & 1;
        mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_k; ++kb, ++g) {
          const int s = g % STAGES;
          if (!ready) mbar_wait(&full[s], (g / STAGES) & 1);
          tc_fence_after();
          ready = mbar_test(&full[(g + 1) % STAGES], ((g + 1) / STAGES) & 1);
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t a_desc = make_kmajor_desc<128>(a_addr);
          const uint64_t b_desc = make_kmajor_desc<128>(a_addr + kATileBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2cta(tmem_base + buf * BN, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2cta(&empty[s], 0x3);
            if (kb == num_k - 1) umma_commit_2cta(&acc_full[buf], 0x3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int blk_bytes = 128 * 128;
    const bool issuer = warp == 2 && lane == 0;
    int i = 0;
    for (int t = pair0; t < total_tiles; t += pair_stride, ++i) {
      int m0, n0, seq;
      bool part_b;
      decode(t, m0, n0, seq, part_b);
      const int buf = i & 1;
      mbar_wait(&acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
      if (i >= 1) {                                  // the staging buffer is free once its previous stores have read it
        if (issuer) tma_store_wait_read();
        named_bar_sync(1, 128);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + (col0 % p.cout));
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          float y0 = __uint_as_float(v[4 * j + 0]) + bb.x, y1 = __uint_as_float(v[4 * j + 1]) + bb.y;
          float y2 = __uint_as_float(v[4 * j + 2]) + bb.z, y3 = __uint_as_float(v[4 * j + 3]) + bb.w;
          if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); y2 = lrelu_fast(y2); y3 = lrelu_fast(y3); }
          w[2 * j] = pack2t<FMT>(y0, y1);
          w[2 * j + 1] = pack2t<FMT>(y2, y3);
        }
        uint8_t* dst = staging + (c >> 1) * blk_bytes + row * 128;  // block = 64 channels, this chunk = half of it
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + j) ^ (row & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&acc_empty[buf]);   // accumulator drained: the MMAs of pair tile i+2 may start
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer && !(p.dbg & 4)) {
        for (int j = 0; j < BN / 64; ++j) {
          const int col = n0 + j * 64;
          const int r = col / p.cout, co0 = col % p.cout;
          const int rr = part_b ? r - p.pad + p.stride : r - p.pad;
          const int mm = part_b ? m0 - 1 : m0;
          tma_store_4d(&tmOut, staging + j * blk_bytes, co0, rr, mm, seq);
        }
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_read();
  }
  tc_fence_before();
  cluster_sync_all();                                  // the peer may still read our shared memory / TMEM until here
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 2 * BN);
}

template <int FMT, bool LRELU>
static int launch_gemm_taps_pair_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                                   const GemmTapsParams& p, int n_seq, cudaStream_t stream) {
  constexpr int STAGES = 4;
  constexpr int SMEM = STAGES * (kATileBytes + 128 * 128) + 128 * 256 * 2 + 256 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static bool configured[16] = {};
  static int sms[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  auto kernel = gemm_taps_pair_kernel<STAGES, FMT, LRELU>;
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    B200_CUDA(cudaDeviceGetAttribute(&sms[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    configured[dev & 15] = true;
  }
  GemmTapsParams pp = p;
  {
    const char* e = getenv("B200VOC_DBG");
    pp.dbg = e ? atoi(e) : 0;
  }
  pp.m_tiles = ceil_div(p.rows_per_seq, 128);
  pp.n_tiles = p.n_total / 256;
  pp.n_seq = n_seq;
  const long long total = (long long)((pp.m_tiles + 1) / 2) * pp.n_tiles * n_seq;
  const long long cap = sms[dev & 15] / 2;
  const int pairs = (int)(total < cap ? total : cap);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmOut, pp));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

static bool gemm_taps_pair_enabled() {          // B200VOC_TAPS_PAIR=0 keeps the single-CTA kernel (A/B runs)
  static const bool on = [] { const char* e = getenv("B200VOC_TAPS_PAIR"); return !(e && e[0] == '0'); }();
  return on;
}

static int launch_gemm_taps_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                                 const GemmTapsParams& p, int n_seq, cudaStream_t stream) {
  if (p.fmt == 0) {
    if (p.store_lrelu) return launch_gemm_taps_pair_t<0, true>(tmA, tmB, tmOut, p, n_seq, stream);
    return launch_gemm_taps_pair_t<0, false>(tmA, tmB, tmOut, p, n_seq, stream);
  }
  if (p.store_lrelu) return launch_gemm_taps_pair_t<1, true>(tmA, tmB, tmOut, p, n_seq, stream);
  return launch_gemm_taps_pair_t<1, false>(tmA, tmB, tmOut, p, n_seq, stream);
}

template <int BN, int STAGES, int FMT, bool LRELU>
static int launch_gemm_taps_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                              const GemmTapsParams& p, int n_seq, cudaStream_t stream) {
  constexpr int SMEM = STAGES * (kATileBytes + BN * 128) + (BN <= 128 ? 2 : 1) * 128 * BN * 2 + 256 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static bool configured[16] = {};
  static int sms[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(gemm_taps_kernel<BN, STAGES, FMT, LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SMEM));
    B200_CUDA(cudaDeviceGetAttribute(&sms[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    configured[dev & 15] = true;
  }
  GemmTapsParams pp = p;
  {
    const char* e = getenv("B200VOC_DBG");
    pp.dbg = e ? atoi(e) : 0;
  }
  pp.m_tiles = ceil_div(p.rows_per_seq, 128);
  pp.n_tiles = p.n_total / BN;
  pp.n_seq = n_seq;
  const long long total = (long long)pp.m_tiles * pp.n_tiles * n_seq;
  const int per_sm = SMEM <= 110 * 1024 ? 2 : 1;        // CTAs that fit per SM (smem and TMEM: 2 x 2*BN <= 512)
  const long long cap = (long long)sms[dev & 15] * per_sm;
  const int grid = (int)(total < cap ? total : cap);
  gemm_taps_kernel<BN, STAGES, FMT, LRELU><<<grid, 192, SMEM, stream>>>(tmA, tmB, tmOut, pp);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int BN, int STAGES>
static int launch_gemm_taps(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                            const GemmTapsParams& p, int n_seq, cudaStream_t stream) {
  if (p.fmt == 0) {
    if (p.store_lrelu) return launch_gemm_taps_t<BN, STAGES, 0, true>(tmA, tmB, tmOut, p, n_seq, stream);
    return launch_gemm_taps_t<BN, STAGES, 0, false>(tmA, tmB, tmOut, p, n_seq, stream);
  }
  if (p.store_lrelu) return launch_gemm_taps_t<BN, STAGES, 1, true>(tmA, tmB, tmOut, p, n_seq, stream);
  return launch_gemm_taps_t<BN, STAGES, 1, false>(tmA, tmB, tmOut, p, n_seq, stream);
}

// ---------------------------------------------------------------------------- ConvT packing
// reference layout w[Cin][Cout][2s] (fp32) -> packed[(r*Cout + co)][tap*Cin + ci], tap 0 = k=r
// (pairs with x[m]), tap 1 = k=r+s (pairs with x[m-1]).
__global__ void pack_convt_kernel(const float* __restrict__ w, int Cin, int Cout, int s, int fmt,
                                  uint16_t* __restrict__ out) {
  const long long total = (long long)s * Cout * 2 * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kcol = (int)(i % (2 * Cin));
    const int nrow = (int)(i / (2 * Cin));
    const int tap = kcol / Cin, ci = kcol % Cin;
    const int r = nrow / Cout, co = nrow % Cout;
    const float v = w[((long long)ci * Cout + co) * (2 * s) + r + tap * s];
    out[i] = fmt == 0 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int convt1d_launch(const void* x16, const void* w_packed, const float* bias, int N, int Lin, int Cin, int Cout, int s,
                   int fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(Cin % 64 == 0 && Cin >= 64, "convt1d: Cin=%d must be a multiple of 64", Cin);
  B200_CHECK_ARG(Cout % 32 == 0, "convt1d: Cout=%d must be a multiple of 32", Cout);
  B200_CHECK_ARG(s >= 2 && s % 2 == 0, "convt1d: stride %d must be even", s);
  B200_CHECK_ARG(N > 0 && Lin > 0, "convt1d: empty input");
  const int n_total = s * Cout;
  CUtensorMap tmA, tmB, tmOut;
  B200_TRY(make_tmap_3d(&tmA, x16, Cin, Lin, N, (uint64_t)Cin * 2, (uint64_t)Lin * Cin * 2, 64, 128, 128));
  // output [N, s*Lin, Cout] viewed as (co, l mod s, l / s, n)
  const int obox = Cout < 64 ? Cout : 64;
  B200_TRY(make_tmap_4d(&tmOut, out16, Cout, s, Lin, N, (uint64_t)Cout * 2, (uint64_t)s * Cout * 2,
                        (uint64_t)s * Lin * Cout * 2, obox, 1, 128, obox * 2));
  GemmTapsParams p{};
  p.rows_per_seq = Lin;      // phases r >= p use rows m = 0..Lin-1, phases r < p use m = 1..Lin (see kernel)
  p.n_taps = 2;
  p.tap_shift[0] = 0;
  p.tap_shift[1] = -1;
  p.k_per_tap = Cin;
  p.n_total = n_total;
  p.cout = Cout;
  p.stride = s;
  p.pad = s / 2;
  p.fmt = fmt;
  p.store_lrelu = store_lrelu;
  p.bias = bias;
  // tiles must be phase-pure w.r.t. the padding boundary: BN divides p*Cout
  const int pc = p.pad * Cout;
  if (pc % 256 == 0 && n_total % 256 == 0) {
    if (gemm_taps_pair_enabled() && (long long)ceil_div(Lin, 128) * N >= 2) {
      B200_TRY(make_tmap_2d(&tmB, w_packed, 2 * Cin, n_total, (uint64_t)2 * Cin * 2, 64, 128, 128));   // half weight tiles
      return launch_gemm_taps_pair(tmA, tmB, tmOut, p, N, stream);
    }
    B200_TRY(make_tmap_2d(&tmB, w_packed, 2 * Cin, n_total, (uint64_t)2 * Cin * 2, 64, 256, 128));
    return launch_gemm_taps<256, 3>(tmA, tmB, tmOut, p, N, stream);
  } else if (pc % 128 == 0 && n_total % 128 == 0) {
    B200_TRY(make_tmap_2d(&tmB, w_packed, 2 * Cin, n_total, (uint64_t)2 * Cin * 2, 64, 128, 128));
    return launch_gemm_taps<128, 4>(tmA, tmB, tmOut, p, N, stream);
  } else if (pc % 64 == 0 && n_total % 64 == 0) {
    B200_TRY(make_tmap_2d(&tmB, w_packed, 2 * Cin, n_total, (uint64_t)2 * Cin * 2, 64, 64, 128));
    return launch_gemm_taps<64, 3>(tmA, tmB, tmOut, p, N, stream);
  } else if (pc % 32 == 0 && n_total % 32 == 0) {
    B200_TRY(make_tmap_2d(&tmB, w_packed, 2 * Cin, n_total, (uint64_t)2 * Cin * 2, 64, 32, 128));
    return launch_gemm_taps<32, 3>(tmA, tmB, tmOut, p, N, stream);
  }
  set_error("convt1d: (s/2)*Cout=%d must be a multiple of 32", pc);
  return B200VOC_ERR_UNSUPPORTED;
}

// Plain projection y[n, l, :] = W x[n, l, :] + b on channels-last 16-bit (1x1 conv).
int linear_launch(const void* x16, const void* w_packed /*[Cout][Cin]*/, const float* bias, int N, int L, int Cin,
                  int Cout, int fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0, "linear: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
  CUtensorMap tmA, tmB, tmOut;
  B200_TRY(make_tmap_3d(&tmA, x16, Cin, L, N, (uint64_t)Cin * 2, (uint64_t)L * Cin * 2, 64, 128, 128));
  B200_TRY(make_tmap_4d(&tmOut, out16, Cout, 1, L, N, (uint64_t)Cout * 2, (uint64_t)Cout * 2, (uint64_t)L * Cout * 2,
                        64, 1, 128, 128));
  GemmTapsParams p{};
  p.rows_per_seq = L;
  p.n_taps = 1;
  p.k_per_tap = Cin;
  p.n_total = Cout;
  p.cout = Cout;
  p.stride = 1;
  p.pad = 0;
  p.fmt = fmt;
  p.store_lrelu = store_lrelu;
  p.bias = bias;
  if (Cout % 256 == 0) {
    B200_TRY(make_tmap_2d(&tmB, w_packed, Cin, Cout, (uint64_t)Cin * 2, 64, 256, 128));
    return launch_gemm_taps<256, 3>(tmA, tmB, tmOut, p, N, stream);
  } else if (Cout % 128 == 0) {
    B200_TRY(make_tmap_2d(&tmB, w_packed, Cin, Cout, (uint64_t)Cin * 2, 64, 128, 128));
    return launch_gemm_taps<128, 4>(tmA, tmB, tmOut, p, N, stream);
  }
  B200_TRY(make_tmap_2d(&tmB, w_packed, Cin, Cout, (uint64_t)Cin * 2, 64, 64, 128));
  return launch_gemm_taps<64, 3>(tmA, tmB, tmOut, p, N, stream);
}

#ifdef B200VOC_DEV
// ---------------------------------------------------------------------------- experiment
// Can a K-major SWIZZLE_128B descriptor start at an arbitrary 128-byte row of a TMA-written tile?
// variant 0: base_offset field = 0; variant 1: base_offset = (start >> 7) & 7.
__global__ void __launch_bounds__(128, 1)
exp_rowshift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem;                 // 144 rows x 128 B = 18432
  uint8_t* sB = smem + 19 * 1024;     // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 144 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  const uint32_t idesc = make_idesc_f16(0, 64);
  uint32_t parity = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int shift = 0; shift < 16; ++shift) {
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + shift * 128;
        const uint64_t a_desc = make_kmajor_desc<128>(a_addr, variant ? ((a_addr >> 7) & 7) : 0);
        const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sB));
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
        umma_commit(&bars[1]);
      }
      mbar_wait(&bars[1], parity);
      parity ^= 1;
      tc_fence_after();
      float* o = out + ((long long)(variant * 16 + shift) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int exp_rowshift_launch(const void* a16, const void* b16, float* out, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  B200_TRY(make_tmap_2d(&tmA, a16, 64, 144, 128, 64, 144, 128));
  B200_TRY(make_tmap_2d(&tmB, b16, 64, 64, 128, 64, 64, 128));
  const int SMEM = 30 * 1024;
  exp_rowshift_kernel<<<1, 128, SMEM, stream>>>(tmA, tmB, out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

// ---------------------------------------------------------------------------- experiment
// Raw tcgen05.mma issue rate from shared-memory operands (SS mode), M = 128, K = 16 per MMA, for a
// given N: one thread per CTA issues `iters` x 4 MMAs on a fixed (uninitialised) operand tile and
// times them with clock64.  out[blockIdx.x] = cycles.  DESIGN.md uses the result to pick N.
__global__ void __launch_bounds__(512, 1) exp_mma_rate_kernel(int n, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;               // 128 rows x 128 B
  uint8_t* sB = smem + 16384;       // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (n >= 30000) {
    // mbarrier hand-off latency.  variant 0/1: warps 0 and 1 ping-pong through two mbarriers (one lane each),
    // waiting with the suspending try_wait (0) or by polling test_wait (1); variant 2: one thread issues one
    // M128 N64 K16 MMA + tcgen05.commit and polls for its completion (issue -> barrier round trip).
    const int variant = n - 30000;
    uint64_t* ping = bar;                                     // reuse the kernel's barrier + one more
    uint64_t* pong = reinterpret_cast<uint64_t*>(smem + 49152 + 64);
    if (threadIdx.x == 0) {
      mbar_init(pong, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (variant < 2) {
      if ((threadIdx.x & 31) == 0 && warp < 2) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          if (warp == 0) {
            mbar_arrive(ping);
            if (variant == 0) mbar_wait(pong, it & 1); else mbar_wait_poll(pong, it & 1);
          } else {
            if (variant == 0) mbar_wait(ping, it & 1); else mbar_wait_poll(ping, it & 1);
            mbar_arrive(pong);
          }
        }
        if (warp == 0) out[blockIdx.x] = clock64() - t0;
      }
    } else if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_f16(0, 64);
      const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sA)), b_desc = make_kmajor_desc<128>(smem_u32(sB));
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        umma_f16(tmem_base, a_desc, b_desc, idesc, 0);
        umma_commit(ping);
        mbar_wait_poll(ping, it & 1);
      }
      out[blockIdx.x] = clock64() - t0;
    }
  } else
  if (n >= 20000) {
    // TMEM read bandwidth: (n - 20000) warps issue `iters` x (tcgen05.ld 32x32b.x32 + wait) on their lane quadrant
    const int nw = (n - 20000) % 100, shape = (n - 20000) / 100;
    if (warp < nw) {
      uint32_t v[32], acc = 0;
      const uint32_t addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 32;
      __syncwarp();
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (shape == 0) tmem_ld32(addr, v);
        else if (shape == 1) tmem_ld_16x256b_x8(addr, v);
        else if (shape == 2) tmem_ld_16x128b_x16(addr, v);
        else tmem_ld_16x64b_x32(addr, v);
        tmem_ld_wait();
        acc += v[0] ^ v[31];
      }
      const long long t1 = clock64();
      if ((threadIdx.x & 31) == 0 && warp == 0) out[blockIdx.x] = (t1 - t0) + (acc == 0x12345u ? 1 : 0);
    }
  } else
  if (threadIdx.x == 0) {
    // n >= 1000 selects 64-byte rows / SWIZZLE_64B operands (the C=32 layout), 2 k-steps per row
    const int shift_rows = n / 10000;       // n >= 10000: A descriptor starts shift_rows rows into the tile
    n %= 10000;
    const bool sw64 = n >= 1000;
    if (sw64) n -= 1000;
    const uint32_t idesc = make_idesc_f16(0, n);
    const uint64_t a_desc = sw64 ? make_kmajor_desc<64>(smem_u32(sA) + shift_rows * 64)
                                 : make_kmajor_desc<128>(smem_u32(sA) + shift_rows * 128);
    const uint64_t b_desc = sw64 ? make_kmajor_desc<64>(smem_u32(sB)) : make_kmajor_desc<128>(smem_u32(sB));
    const int kmask = sw64 ? 1 : 3;
    // iters >= 1000000: group experiment.  iters = 1000000*mode + 1000*G + reps: issue `reps` rounds of
    // two groups of G MMAs; mode 1: the groups alternate between two accumulators (same shape);
    // mode 2: same accumulator, second group starts with accumulate = 0; mode 3: two accumulators AND
    // the second group has N/2; mode 4: one accumulator, all accumulate (control).
    const int mode = iters / 1000000, G = (iters / 1000) % 1000, reps = iters % 1000;
    const long long t0 = clock64();
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_desc + 2 * (k & kmask), b_desc + 2 * (k & kmask), idesc, 1);
      }
    } else {
      const uint32_t idesc_b = mode == 3 ? make_idesc_f16(0, n / 2) : idesc;
      const uint32_t d_b = (mode == 1 || mode == 3) ? tmem_base + 128 : tmem_base;
      for (int it = 0; it < reps; ++it) {
        for (int k = 0; k < G; ++k) umma_f16(tmem_base, a_desc + 2 * (k & kmask), b_desc + 2 * (k & kmask), idesc, 1);
        for (int k = 0; k < G; ++k)
          umma_f16(d_b, a_desc + 2 * (k & kmask), b_desc + 2 * (k & kmask), idesc_b, (mode == 2 && k == 0) ? 0 : 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int exp_mma_rate_launch(int n, int iters, int blocks, long long* out, cudaStream_t stream) {
  B200_CHECK_ARG(n >= 20000 || ((n % 1000) % 16 == 0 && (n % 1000) >= 16 && (n % 1000) <= 256 && n / 10000 <= 16), "exp_mma_rate: N=%d", n);
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(exp_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
    configured = true;
  }
  exp_mma_rate_kernel<<<blocks, 512, 52 * 1024, stream>>>(n, iters, out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

#endif  // B200VOC_DEV

int pack_convt_launch(const float* w, int Cin, int Cout, int s, int fmt, void* out, cudaStream_t stream) {
  const long long total = (long long)s * Cout * 2 * Cin;
  pack_convt_kernel<<<(int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, stream>>>(
      w, Cin, Cout, s, fmt, reinterpret_cast<uint16_t*>(out));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
