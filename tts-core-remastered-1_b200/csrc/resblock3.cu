// K2c: persistent fused residual block for the wide stages (C = 128, 256) -- the tensor-bound
// part of the network (81 % of all FLOPs live in residual blocks, stages 0-1 hold two thirds).
//
// Same math as resblock2.cu (generator.py:40-41,89-90 / repair R2).  Schedule:
//   * persistent CTA PAIRS (a cluster of 2, template PAIR): the two CTAs own two neighbouring 128-row
//     tiles; the leader issues one tcgen05.mma.cta_group::2 (M = 256, N = 128) per k-step, each CTA
//     supplies its own rows of A and HALF of the B (weight) tile, the accumulators sit at the same
//     TMEM address in both CTAs.  Commits are multicast to both CTAs; the peer's epilogue warps arrive
//     on the leader's barriers (mapa + mbarrier.arrive.shared::cluster); the peer's weight TMA signals
//     the leader's barrier (cp.async.bulk.tensor ... cta_group::2).  PAIR = false is the single-CTA
//     schedule the pair grew out of (B200VOC_RB3_PAIR=0);
//   * the leaky_relu(x) tile (128 rows + 8-row halo each side, all channels) is loaded ONCE per
//     tile; the three dilated taps are row-shifted UMMA descriptors over that tile;
//   * weights do not fit in shared memory (W1 is 192/768 KB), so they stream from L2 through a TMA
//     ring; what paced the single-CTA kernel was the ring (5 x 16 KB = 1.3 k clk of MMAs against a
//     ~1.8 k clk slot round trip) and the single issuing warp's instruction path per slot (as long as
//     the 4 MMAs it fed): a pair stages half tiles, so the same 80 KB hold 5 slots of 8 MMAs;
//   * GEMM1 runs in chunks of 64 value + 64 gate channels into two alternating TMEM accumulators;
//     the GLU/FiLM epilogue of chunk j (8 warps, 32 channels each) overlaps the MMAs of chunk j+1; it
//     writes h as k-block j of GEMM2's A operand; GEMM2's k-block j is issued one chunk later, the
//     last one after the first chunk of the NEXT tile;
//   * the residual/store epilogue (8 more warps) drains D2 while the next tile's GEMM1 runs;
//   * separate producer warps for the weight ring and the input tiles; barriers are arrived on once
//     per warp.  Measured bounds and the experiments behind these choices: DESIGN.md section 4b,
//     profiles/r01_knockouts.txt.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct Resblock3Params {
  int L, dilation, T, P, num_bands;
  int tiles_per_seq, total_tiles, n_seq;
  const uint16_t* a16;   // [N, L, C] leaky_relu(x)
  const float* b_conv;   // [2C]
  const float* b_proj;   // [C]
  const float* film;     // [B, T, film_stride]
  int film_stride;
  uint16_t* out;         // [N, L, C]
  int dbg;               // knock-out experiment switches, only with -DB200VOC_TRACE (B200VOC_DBG)
  long long* trace;      // clock64 timeline of CTA 0, [7][64][4] (trace builds)
};

#ifdef B200VOC_TRACE
#define RB3_DBG(bit) (p.dbg & (bit))
#define RB3_TRACE(slot, i, k)                                                             \
  do {                                                                                    \
    if (p.trace && blockIdx.x == 0 && (i) < 64 && (threadIdx.x & 31) == 0)                 \
      p.trace[(((slot) * 64 + (i)) << 2) + (k)] = clock64();                               \
  } while (0)
#else
#define RB3_DBG(bit) false
#define RB3_TRACE(slot, i, k) do { } while (0)
#endif
extern long long* g_rb2_trace;

#ifndef RB3_BIAS_MMA
#define RB3_BIAS_MMA 1
#endif
#ifndef RB3_WS
#define RB3_WS 1
#endif
template <int C, bool PAIR>
struct Rb3Cfg {
  static constexpr int KPT = C / 64;                 // k-blocks per tap = GEMM1 chunks = GEMM2 k-blocks
  static constexpr int NCH = KPT;
  static constexpr int NH = C / 128;                 // GEMM2 output halves of 128 columns
  static constexpr int HALO = 8;
  static constexpr int A_ROWS = 128 + 2 * HALO;
  static constexpr int A_KB_BYTES = A_ROWS * 128;    // 18 KB per k-block
  static constexpr int A_BYTES = KPT * A_KB_BYTES;
  static constexpr bool INPLACE = C == 128;          // stage the output over the input tile + TMA store
  static constexpr int NA = C == 128 ? 3 : (RB3_WS != 0 && PAIR ? 2 : 1);   // input tiles have their own producer warp (18)
  static constexpr int H_KB_BYTES = 128 * 128;       // 16 KB per k-block
  // PAIR: two CTAs (a cluster of 2) work on two neighbouring row tiles with ONE tcgen05.mma.cta_group::2 per
  // k-step (M = 256): each CTA stages only HALF of every weight tile (64 of its 128 rows), so the same
  // shared-memory budget holds twice as many ring slots -- the weight ring's round trip was what paced the
  // single-CTA kernel (tests/trace_resblock3.py).
  static constexpr int W_ROWS = PAIR ? 64 : 128;
  static constexpr int W_TILE = W_ROWS * 128;        // ring slot: [W_ROWS rows x 64 k]
  // weight ring: a slot's round trip (MMAs complete -> commit -> producer -> TMA from L2 -> issuer) is ~1800 clk
  // against 256 clk of MMA work per slot, measured (tests/trace_resblock3.py): the ring must hold ~7 slots
  // PAIR: a ring slot holds SUBS = 2 sub-tiles (two k-blocks of one tap, or the two column halves of a GEMM2
  // k-block), i.e. 8 MMAs per barrier hand-off: the MMA issuer is a single warp whose per-slot instruction path
  // (~50 dependent instructions, sharing its scheduler with 4 epilogue warps) costs about as much as 4 MMAs.
  static constexpr int SUBS = PAIR ? 2 : 1;
  static constexpr int W_SLOT = SUBS * W_TILE;
  static constexpr int NW = (RB3_WS != 0 && PAIR && C == 256) ? 4 : 5;
  static constexpr int ND2 = C == 128 ? 2 : 1;
  // WS (C = 128 pairs): WEIGHT STATIONARY -- a pair holds all of W1 and W2 (each CTA its half of every tile: 96 + 16
  // KB), loaded once per launch, and h goes back into TMEM over the value columns of the GEMM1 accumulator it came from
  // (packed 16-bit pairs, tcgen05.st) to be GEMM2's A operand (TS-mode MMA), so there is no h buffer and no weight ring:
  // the kernel is bound by the shared-memory data pipe and the ring's TMA writes (112 KB per tile), the
  // h stores and GEMM2's operand reads were a quarter of what that pipe carried; the issuer loses its per-slot barrier
  // wait + commit.  The accumulator buffer is handed back by GEMM2's commit (it holds h until then).
  static constexpr bool TSH = RB3_WS != 0 && PAIR;          // h through TMEM (both widths).  C = 256 keeps its weight ring
                                                           // (W1 is 768 KB) and spends the h buffer's 64 KB on a SECOND input tile
  static constexpr bool WS = TSH && C == 128;
  static constexpr int N_W1_TILES = NCH * 3 * KPT, N_W_TILES = N_W1_TILES + KPT * NH;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_H = OFF_A + NA * A_BYTES;
  static constexpr int OFF_W = OFF_H + (TSH ? 0 : KPT * H_KB_BYTES);
  static constexpr int OFF_BAR = OFF_W + (WS ? N_W_TILES * W_TILE : NW * W_SLOT);
  static constexpr int OFF_PAR = OFF_BAR + 512;
  // BIAS_MMA: the three bias vectors are added on the tensor core -- one more K = 16 MMA per accumulator whose A
  // operand is an all-ones tile and whose B operand carries (b/2)_hi, (b/2)_lo in two K columns (both K core matrices
  // alias the same block, so the product is 2 (b/2) = b at hi + lo precision).  The epilogues' broadcast
  // shared-memory loads of the biases were half of their LDS traffic (a broadcast LDS.128 costs four wavefronts) in
  // a kernel whose shared-memory pipe is ~72 % busy (33 % tensor-core operand reads + 38 % LSU, ncu).
  static constexpr bool BIAS_MMA = RB3_BIAS_MMA != 0 && PAIR;        // (the single-CTA fallback has no room for the tiles)
  static constexpr int B_ROWS = W_ROWS;                     // bias-tile rows this CTA supplies per 128 GEMM columns
  static constexpr int BIAS_TILE = B_ROWS * 16;             // one K core matrix column: [rows][8 x 16 bit]
  static constexpr int N_BIAS_TILES = NCH + NH;             // GEMM1 chunks, then GEMM2 halves
  static constexpr int OFF_ONES = OFF_PAR;                  // 128 bytes of 1.0 (BIAS_MMA: replaces the fp32 bias block)
  static constexpr int OFF_BIAS = OFF_ONES + 128;
  static constexpr int PAR_BYTES = BIAS_MMA ? 128 + N_BIAS_TILES * BIAS_TILE : 3 * C * 4;
  static_assert((2 * NA + 2 * NW + 4 + 2 * KPT + 2 * (C == 128 ? 2 : 1) + NA) * 8 + 8 <= 512, "barrier block");
  static constexpr int OFF_FILM = OFF_PAR + PAR_BYTES;    // per GLU warp: (1+scale | shift) of its 32 channels of a chunk
  static constexpr int SMEM = OFF_FILM + 8 * 64 * 4 + 1024;
  static constexpr int D2_COL = 256;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr int A_PREFETCH_AFTER_CHUNK = NA >= 2 ? 0 : NCH - 1;
  static_assert(OFF_H % 1024 == 0 && OFF_W % 1024 == 0 && A_KB_BYTES % 1024 == 0, "1024B alignment for SWIZZLE_128B");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int C, int FMT, int OFMT, bool LRELU, bool PAIR>
__global__ void __launch_bounds__(608, 1)
resblock3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                 const Resblock3Params p) {
  using K = Rb3Cfg<C, PAIR>;
  constexpr int KPT = K::KPT, NCH = K::NCH, NH = K::NH, NA = K::NA, NW = K::NW, ND2 = K::ND2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem + K::OFF_A;
  uint8_t* sH = smem + K::OFF_H;
  uint8_t* sW = smem + K::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* a_full = bars;                  // [NA]
  uint64_t* a_empty = a_full + NA;          // [NA]
  uint64_t* w_full = a_empty + NA;          // [NW]
  uint64_t* w_empty = w_full + NW;          // [NW]
  uint64_t* d1_full = w_empty + NW;         // [2]
  uint64_t* d1_empty = d1_full + 2;         // [2]
  uint64_t* h_full = d1_empty + 2;          // [KPT]
  uint64_t* h_empty = h_full + KPT;         // [KPT]
  uint64_t* d2_full = h_empty + KPT;        // [ND2]
  uint64_t* d2_empty = d2_full + ND2;       // [ND2]
  uint64_t* a_peer = d2_empty + ND2;        // [NA]  (PAIR, leader) the peer CTA's input tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_peer + NA);
  float* sPar = reinterpret_cast<float*>(smem + K::OFF_PAR);   // [ba | bg/2 | b2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;     // 0 = leader: issues the MMAs, owns the barriers the issuer waits on
  constexpr int kCtas = PAIR ? 2 : 1;
  // multi-thread barriers are arrived on once per WARP (fence, __syncwarp, lane 0): a 128-arrival
  // barrier wakes the waiting MMA issuer ~25 times per phase (measured in resblock2.cu)
  constexpr int kE1Warps = 8;                           // both GLU warp sets share every chunk (32 channels each)
  constexpr int kE2Warps = K::INPLACE ? 4 : 8;          // tile-parity set, or all 8 warps (C=256)

  if (!K::BIAS_MMA) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      sPar[i] = 0.5f * p.b_conv[i];              // W1 is packed pre-scaled by 1/2 (see pack_resblock_kernel)
      sPar[C + i] = 0.5f * p.b_conv[C + i];
      sPar[2 * C + i] = p.b_proj[i];
    }
  } else {
    // all-ones A block and the bias B tiles (SWIZZLE_NONE K-major core matrices: 8 rows x 16 bytes, 8-row groups 128
    // bytes apart).  Tile t < NCH: GEMM1 chunk t, GEMM column n <-> value channel 64 t + n (n < 64) or gate channel
    // 64 t + n - 64; tile NCH + h: GEMM2 output channels 128 h + n.  A pair CTA supplies columns rank*64 .. +63.
    uint16_t* ones = reinterpret_cast<uint16_t*>(smem + K::OFF_ONES);
    uint32_t* bt = reinterpret_cast<uint32_t*>(smem + K::OFF_BIAS);
    const uint32_t rank_s = PAIR ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) ones[i] = FMT == 0 ? 0x3C00 : 0x3F80;
    for (int i = threadIdx.x; i < K::N_BIAS_TILES * K::B_ROWS; i += blockDim.x) {
      const int t = i / K::B_ROWS, r = i - t * K::B_ROWS;
      const int n = (PAIR ? (int)rank_s * 64 : 0) + r;
      float bv;
      if (t < K::NCH) bv = 0.25f * (n < 64 ? p.b_conv[t * 64 + n] : p.b_conv[C + t * 64 + n - 64]);   // (b / 2) / 2: W1 is pre-scaled by 1/2
      else bv = 0.5f * p.b_proj[(t - K::NCH) * 128 + n];
      const float lo = bv - unpack2t<FMT>(pack2t<FMT>(bv, 0.f)).x;
      uint32_t* row = bt + (t * K::BIAS_TILE >> 2) + r * 4;
      row[0] = pack2t<FMT>(bv, lo);
      row[1] = 0u; row[2] = 0u; row[3] = 0u;
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int b = 0; b < NA; ++b) { mbar_init(&a_full[b], 1); mbar_init(&a_empty[b], 1); }
    for (int b = 0; b < NW; ++b) { mbar_init(&w_full[b], 1); mbar_init(&w_empty[b], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&d1_full[b], 1); mbar_init(&d1_empty[b], K::WS ? 1 : kCtas * kE1Warps); }
    for (int b = 0; b < KPT; ++b) { mbar_init(&h_full[b], kCtas * kE1Warps); mbar_init(&h_empty[b], 1); }
    for (int b = 0; b < ND2; ++b) { mbar_init(&d2_full[b], 1); mbar_init(&d2_empty[b], kCtas * kE2Warps); }
    for (int b = 0; b < NA; ++b) mbar_init(&a_peer[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(tmem_slot, K::TMEM_COLS); else tmem_alloc(tmem_slot, K::TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();       // barrier inits (and the pair's TMEM) visible to the peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // work units: a unit is one row tile (single CTA) or two neighbouring row tiles (pair: CTA `rank` takes tile
  // 2u + rank; a tile index past the end is an all-out-of-bounds tile -- TMA zero-fills loads and clips stores)
  const int n_units = PAIR ? (p.total_tiles + 1) / 2 : p.total_tiles;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_my_tiles = unit0 < n_units ? (n_units - unit0 + unit_stride - 1) / unit_stride : 0;
  auto tile_of = [&](int it) { return PAIR ? 2 * (unit0 + it * unit_stride) + (int)rank : unit0 + it * unit_stride; };
  // barriers the issuer waits on live in the leader CTA: arrive there (remote for the peer)
  auto arrive_leader = [&](uint64_t* bar) {
    if (!PAIR || rank == 0) mbar_arrive(bar); else mbar_arrive_remote(mapa_u32(bar, 0));
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer of the weight ring
    if (lane == 0) {
      int wi = 0;
      auto w_slot = [&](int nsub) -> uint8_t* {
        const int s = wi % NW;
        mbar_wait(&w_empty[s], ((wi / NW) & 1) ^ 1);
        if (RB3_DBG(4) && wi >= NW) { if (rank == 0) mbar_arrive(&w_full[s]); return nullptr; }
        (void)nsub;
        if (rank == 0) mbar_expect_tx(&w_full[s], kCtas * nsub * K::W_TILE);   // both CTAs' bytes land on the leader's barrier
        return sW + s * K::W_SLOT;
      };
      auto load_w = [&](uint8_t* dst, const CUtensorMap* tm, int c0, int c1) {
        if (PAIR) tma_load_2d_2cta(dst, tm, &w_full[wi % NW], c0, c1 + rank * 64);
        else tma_load_2d(dst, tm, &w_full[wi % NW], c0, c1);
      };
      auto load_w2 = [&](int kb) {
        if (K::SUBS == 2) {                       // one slot: the NH column halves of this k-block
          uint8_t* dst = w_slot(NH);
          if (dst)
            for (int half = 0; half < NH; ++half) load_w(dst + half * K::W_TILE, &tmW2, kb * 64, half * 128);
          ++wi;
        } else {
          for (int half = 0; half < NH; ++half) {
            uint8_t* dst = w_slot(1);
            if (dst) load_w(dst, &tmW2, kb * 64, half * 128);
            ++wi;
          }
        }
      };
      if (K::WS) {
        // all weights once: tile (chunk j, tap, k-block kb) of W1 at index (j*3 + tap)*KPT + kb, then W2's k-blocks
        if (rank == 0) mbar_expect_tx(&w_full[0], kCtas * K::N_W_TILES * K::W_TILE);
#pragma unroll 1
        for (int t = 0; t < K::N_W1_TILES; ++t) {
          const int j = t / (3 * KPT), tap = (t / KPT) % 3, kb = t % KPT;
          tma_load_2d_2cta(sW + t * K::W_TILE, &tmW1, &w_full[0], tap * C + kb * 64, j * 128 + rank * 64);
        }
#pragma unroll 1
        for (int kb = 0; kb < KPT; ++kb)
          tma_load_2d_2cta(sW + (K::N_W1_TILES + kb) * K::W_TILE, &tmW2, &w_full[0], kb * 64, rank * 64);
      } else {
      // compact code on purpose (no unrolling in the producer / issuer roles): the kernel's instruction footprint
      // exceeds the instruction caches and the issuer's fetch stalls were its largest stall reason (ncu: no_instruction)
      const int total_chunks = n_my_tiles * NCH;
#pragma unroll 1
      for (int gc = 0, j = 0; gc <= total_chunks; ++gc) {
        if (gc < total_chunks) {
#pragma unroll 1
          for (int s3 = 0; s3 < 3 * (KPT / K::SUBS); ++s3) {
            const int tap = s3 / (KPT / K::SUBS), kb = (s3 - tap * (KPT / K::SUBS)) * K::SUBS;
            uint8_t* dst = w_slot(K::SUBS);
            if (dst) {
#pragma unroll
              for (int sub = 0; sub < K::SUBS; ++sub)
                load_w(dst + sub * K::W_TILE, &tmW1, tap * C + (kb + sub) * 64, j * 128);
            }
            ++wi;
          }
        }
        if (gc >= 1) load_w2((gc - 1) % NCH);
        if (++j == NCH) j = 0;
      }
      }
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------ TMA producer of the input tiles (own warp: a
    // wait for a free input slot must not hold up the weight stream)
    if (lane == 0) {
#pragma unroll 1
      for (int it = 0; it < n_my_tiles; ++it) {
        const int tile = tile_of(it);
        const int seq = tile / p.tiles_per_seq, l0 = (tile - seq * p.tiles_per_seq) * 128;
        const int ab = it % NA;
        mbar_wait(&a_empty[ab], ((it / NA) & 1) ^ 1);
        if (RB3_DBG(16) && it >= NA) { mbar_arrive(&a_full[ab]); continue; }
        RB3_TRACE(0, it, 0);
        mbar_expect_tx(&a_full[ab], K::A_BYTES);
#pragma unroll 1
        for (int kb = 0; kb < KPT; ++kb)
          tma_load_3d(sA + ab * K::A_BYTES + kb * K::A_KB_BYTES, &tmX, &a_full[ab], kb * 64, l0 - K::HALO, seq);
      }
    }
  } else if (warp == 1 && PAIR && rank != 0) {
    // ------------------------------------------------------------ peer CTA: no MMAs to issue; tell the leader
    // when each of our input tiles has landed
    if (lane == 0) {
      const uint32_t remote = mapa_u32(&a_peer[0], 0);
      for (int it = 0; it < n_my_tiles; ++it) {
        const int ab = it % NA;
        mbar_wait(&a_full[ab], (it / NA) & 1);
        mbar_arrive_remote(remote + ab * 8);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer.  The whole warp runs the
    // (warp-uniform) control flow so that indices and descriptors live in uniform registers; only the
    // tcgen05.mma / commit instructions are executed by one elected lane.
    {
      const uint32_t idesc = PAIR ? make_idesc_f16_m(FMT, 256, 128) : make_idesc_f16(FMT, 128);
      // descriptors as (low word, common high word): every operand tile here is K-major SWIZZLE_128B
      const uint32_t desc_hi = (uint32_t)(make_kmajor_desc<128>(0) >> 32);
      auto desc_lo = [](uint32_t smem_addr) { return (uint32_t)make_kmajor_desc<128>(smem_addr); };
      auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
        if (PAIR) umma_f16_2cta_lh(d, a_lo, b_lo, desc_hi, idesc, acc); else umma_f16_lh(d, a_lo, b_lo, desc_hi, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if (PAIR) umma_commit_2cta(bar, 0x3); else umma_commit(bar);      // PAIR: same barrier in both CTAs
      };
      // bias MMA: D += ones[rows x 16] * bias_tile[128 x 16]^T.  SWIZZLE_NONE K-major descriptors (layout 0, version
      // 1): the all-ones block is ONE 128-byte core matrix (LBO = SBO = 0: every core matrix of the operand aliases
      // it); the bias tile's 8-row groups are 128 bytes apart (SBO) and both K core matrices alias (LBO = 0)
      const uint64_t ones_desc = (uint64_t)((smem_u32(smem + K::OFF_ONES) & 0x3FFFFu) >> 4) | (1ull << 46);
      auto bias_mma = [&](uint32_t d, int tile) {
        const uint64_t b_desc = (uint64_t)((smem_u32(smem + K::OFF_BIAS + tile * K::BIAS_TILE) & 0x3FFFFu) >> 4) |
                                ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        if (PAIR) umma_f16_2cta(d, ones_desc, b_desc, idesc, 1); else umma_f16(d, ones_desc, b_desc, idesc, 1);
      };
      auto wait_x = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };   // (barriers with remote arrivals)
      int wi = 0;
      // The weight-ring wait is software pipelined: while the MMAs of slot wi are being issued, a
      // non-blocking probe of slot wi+1 is already in flight, so a ready slot costs no round trip.
      bool w_ready = false;
      auto acquire_w = [&]() -> int {
        const int s = wi % NW;
        if (!w_ready) mbar_wait(&w_full[s], (wi / NW) & 1);
        tc_fence_after();
        const int wn = wi + 1;
        w_ready = mbar_test(&w_full[wn % NW], (wn / NW) & 1);
        return s;
      };
      // TSH: GEMM2 k-block of chunk gc2 with h read from TMEM (buffer gc2 & 1, k-step k at columns 16k .. 16k+7 of the
      // accumulator's value half); W2's k-block is resident (WS) or the next ring slot (both column halves)
      auto g2_ts = [&](int gc2) {
        const int it2 = gc2 / NCH, kb = gc2 - it2 * NCH, b2 = gc2 & 1, db = it2 % ND2;
        const uint32_t phh = (gc2 >> 1) & 1, phd = ((it2 / ND2) & 1) ^ 1;
        const bool rh = mbar_test(&h_full[b2], phh), rd = kb == 0 ? mbar_test(&d2_empty[db], phd) : true;
        if (!rh) wait_x(&h_full[b2], phh);
        RB3_TRACE(2, gc2, 0);
        if (!rd) wait_x(&d2_empty[db], phd);
        RB3_TRACE(2, gc2, 1);
        tc_fence_after();
        int s = 0;
        if (!K::WS) s = acquire_w();
        const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(K::WS ? sW + (K::N_W1_TILES + kb) * K::W_TILE : sW + s * K::W_SLOT));
        if (elect_one()) {
          if (!RB3_DBG(8))
#pragma unroll
          for (int half = 0; half < NH; ++half)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2cta_ts(tmem_base + K::D2_COL + db * C + half * 128, tmem_base + b2 * 128 + 16 * k,
                               b_desc + (half * K::W_TILE >> 4) + 2 * k, idesc, (kb | k) != 0);
          if (!K::WS) commit(&w_empty[s]);
          if (kb == NCH - 1) {
            if (K::BIAS_MMA) {
#pragma unroll
              for (int half = 0; half < NH; ++half) bias_mma(tmem_base + K::D2_COL + db * C + half * 128, NCH + half);
            }
            commit(&d2_full[db]);
          }
        }
        __syncwarp();
        if (!K::WS) ++wi;
      };
      auto g2 = [&](int it2, int kb) {
        const int db = it2 % ND2;
        const uint32_t phh = it2 & 1, phd = ((it2 / ND2) & 1) ^ 1;
        const bool rh = mbar_test(&h_full[kb], phh), rd = kb == 0 ? mbar_test(&d2_empty[db], phd) : true;
        if (!rh) wait_x(&h_full[kb], phh);
        RB3_TRACE(2, it2 * NCH + kb, 0);
        if (!rd) wait_x(&d2_empty[db], phd);
        RB3_TRACE(2, it2 * NCH + kb, 1);
        tc_fence_after();
        const uint32_t a_desc = desc_lo(smem_u32(sH + kb * K::H_KB_BYTES));
        if (K::SUBS == 2) {
          const int s = acquire_w();
          const uint32_t b_desc0 = desc_lo(smem_u32(sW + s * K::W_SLOT));
          if (elect_one()) {
            if (!RB3_DBG(8))
#pragma unroll
            for (int half = 0; half < NH; ++half)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma(tmem_base + K::D2_COL + db * C + half * 128, a_desc + 2 * k, b_desc0 + (half * K::W_TILE >> 4) + 2 * k,
                    (kb | k) != 0);
            commit(&w_empty[s]);
          }
          __syncwarp();
          ++wi;
        } else
        for (int half = 0; half < NH; ++half) {
          const int s = acquire_w();
          const uint32_t b_desc = desc_lo(smem_u32(sW + s * K::W_SLOT));
          if (elect_one()) {
            if (!RB3_DBG(8))
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(tmem_base + K::D2_COL + db * C + half * 128, a_desc + 2 * k, b_desc + 2 * k, (kb | k) != 0);
            commit(&w_empty[s]);
          }
          __syncwarp();
          ++wi;
        }
        if (elect_one()) {
          commit(&h_empty[kb]);
          if (kb == NCH - 1) {
            if (K::BIAS_MMA) {
#pragma unroll
              for (int half = 0; half < NH; ++half) bias_mma(tmem_base + K::D2_COL + db * C + half * 128, NCH + half);
            }
            commit(&d2_full[db]);
          }
        }
        __syncwarp();
      };
      const int total_chunks = n_my_tiles * NCH;
#pragma unroll 1
      for (int gc = 0, it = 0, j = 0; gc <= total_chunks; ++gc) {
        if (gc < total_chunks) {
          const int ab = it % NA;
          const int b = gc & 1;
          {
            const uint32_t pha = (it / NA) & 1, phd = ((gc >> 1) & 1) ^ 1;
            // WS: no d1_empty -- GEMM1 of chunk gc overwrites the buffer GEMM2 of chunk gc-2 read h from, and that MMA
            // was issued earlier by this thread (the tensor pipe runs one thread's MMAs in order)
            const bool ra = j == 0 ? mbar_test(&a_full[ab], pha) : true, rd = K::TSH ? true : mbar_test(&d1_empty[b], phd);
            if (!ra) mbar_wait(&a_full[ab], pha);
            if (PAIR && j == 0) wait_x(&a_peer[ab], pha);
            RB3_TRACE(1, gc, 0);
            if (!rd) wait_x(&d1_empty[b], phd);
            RB3_TRACE(1, gc, 1);
          }
          tc_fence_after();
          const uint32_t a_tile = smem_u32(sA + ab * K::A_BYTES) + K::HALO * 128;
          if (K::WS) {
            if (gc == 0) { mbar_wait(&w_full[0], 0); tc_fence_after(); }      // the resident weights have landed (both CTAs' halves)
            const uint32_t w_base = smem_u32(sW + j * 3 * KPT * K::W_TILE);
#pragma unroll 1
            for (int tap = 0; tap < 3; ++tap) {
              const uint32_t a_desc = desc_lo(a_tile + (tap - 1) * p.dilation * 128);
              const uint32_t b_desc = desc_lo(w_base + tap * KPT * K::W_TILE);
              if (elect_one()) {
                if (!RB3_DBG(32))
#pragma unroll
                for (int kb = 0; kb < KPT; ++kb)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    mma(tmem_base + b * 128, a_desc + (kb * K::A_KB_BYTES >> 4) + 2 * k,
                        b_desc + (kb * K::W_TILE >> 4) + 2 * k, (tap | kb | k) != 0);
              }
              __syncwarp();
            }
          } else
#pragma unroll 1
          for (int s3 = 0; s3 < 3 * (KPT / K::SUBS); ++s3) {
            const int tap = s3 / (KPT / K::SUBS), kb = (s3 - tap * (KPT / K::SUBS)) * K::SUBS;
            const int s = acquire_w();
            const uint32_t a_desc = desc_lo(a_tile + kb * K::A_KB_BYTES + (tap - 1) * p.dilation * 128);
            const uint32_t b_desc = desc_lo(smem_u32(sW + s * K::W_SLOT));
            if (elect_one()) {
              if (!RB3_DBG(32))
#pragma unroll
              for (int sub = 0; sub < K::SUBS; ++sub)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma(tmem_base + b * 128, a_desc + (sub * K::A_KB_BYTES >> 4) + 2 * k,
                      b_desc + (sub * K::W_TILE >> 4) + 2 * k, (s3 | sub | k) != 0);
              commit(&w_empty[s]);
            }
            __syncwarp();
            ++wi;
          }
          if (elect_one()) {
            if (K::BIAS_MMA) bias_mma(tmem_base + b * 128, j);
            commit(&d1_full[b]);
            if (!K::INPLACE && j == NCH - 1) commit(&a_empty[ab]);        // INPLACE: released by the store epilogue
          }
          __syncwarp();
          RB3_TRACE(1, gc, 2);
        }
        if (gc >= 1) {
          if (K::TSH) {
            g2_ts(gc - 1);
          } else {
            const int pj = j == 0 ? NCH - 1 : j - 1;
            g2(j == 0 ? it - 1 : it, pj);
          }
        }
        if (gc < total_chunks) RB3_TRACE(1, gc, 3);
        if (++j == NCH) { j = 0; ++it; }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue 1 (warps 2..9): GLU + FiLM -> h
    // two warps per TMEM lane quadrant; BOTH sets work on every chunk, warp set `par` on 32 of its 64
    // value channels: the time an accumulator buffer is held (which gates the MMAs of chunk j+2)
    // halves, and knocking this epilogue out entirely is worth ~9 % (tests/knock_resblock3.py)
    const int q = warp & 3, par = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sBA = reinterpret_cast<const float4*>(sPar);
    const float4* sNB = reinterpret_cast<const float4*>(sPar + C);
    float* scratch = reinterpret_cast<float*>(smem + K::OFF_FILM) + (warp - 2) * 64;     // [S(32) | T(32)]
    int gc = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = tile_of(it);
      int seq = tile / p.tiles_per_seq;
      const int lw = (tile - seq * p.tiles_per_seq) * 128 + q * 32;   // first row of this warp
      const int l = lw + lane;
      if (seq > p.n_seq - 1) seq = p.n_seq - 1;               // pair mode: the tile past the end is out of bounds
      // FiLM: when the warp's 32 rows lie in one frame (P >= 32 and aligned: stage 1 of the generator) the
      // frame's coefficients for this warp's 32 channels are staged once per chunk in shared memory and read
      // back as broadcasts -- a warp-wide LDG.128 of one address costs 4-5 L1 wavefronts (resblock2.cu)
      const int tw = lw / p.P;
      const bool uniform = (lw - tw * p.P) + 31 < p.P;
      int t = uniform ? tw : l / p.P;
      if (t > p.T - 1) t = p.T - 1;
      const float* film = p.film + ((long long)(seq / p.num_bands) * p.T + t) * p.film_stride;
      if (!uniform) {
#pragma unroll
        for (int k = 0; k < (2 * C) / 32; ++k) prefetch_l1(film + k * 32);   // this row's FiLM line(s) -> L1
      }
      for (int j = 0; j < NCH; ++j, ++gc) {
        const int b = gc & 1;
        const int ch0 = j * 64 + par * 32;                    // this warp's first channel in this chunk
        float4 stage = make_float4(0.f, 0.f, 0.f, 0.f);
        if (uniform && lane < 16)                             // lanes 0-7: S, 8-15: T
          stage = __ldg(reinterpret_cast<const float4*>(film + (lane < 8 ? 0 : C) + ch0) + (lane & 7));
        mbar_wait(&d1_full[b], (gc >> 1) & 1);
        if (q == 0 && par == 0) RB3_TRACE(3, gc, 0);
        if (!K::TSH) mbar_wait(&h_empty[j], (it & 1) ^ 1);
        if (q == 0 && par == 0) RB3_TRACE(3, gc, 1);
        tc_fence_after();
        if (uniform) {
          if (lane < 16) reinterpret_cast<float4*>(scratch)[lane] = stage;
          __syncwarp();
        }
        uint8_t* hrow = sH + j * K::H_KB_BYTES + (q * 32 + lane) * 128;
        const int row = q * 32 + lane;
        if (!RB3_DBG(1))
#pragma unroll
        for (int cc = 0; cc < 32; cc += 16) {   // column inside this warp's 32 value channels
          const int cl = par * 32 + cc;
          uint32_t va[16], vg[16];
          if (RB3_DBG(512)) {
#pragma unroll
            for (int z = 0; z < 16; ++z) { va[z] = 0x3f000000u + z; vg[z] = 0x3e000000u + lane; }
          } else {
            tmem_ld16(lane_addr + b * 128 + cl, va);
            tmem_ld16(lane_addr + b * 128 + 64 + cl, vg);
          }
          const int ch = j * 64 + cl;
          float4 S[4], H[4];
          if (RB3_DBG(128)) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) { S[i4] = make_float4(1.f, 1.f, 1.f, 1.f); H[i4] = make_float4(0.f, 0.f, 0.f, 0.f); }
          } else if (uniform) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              S[i4] = reinterpret_cast<const float4*>(scratch + cc)[i4];
              H[i4] = reinterpret_cast<const float4*>(scratch + 32 + cc)[i4];
            }
          } else {
            const float4* fs = reinterpret_cast<const float4*>(film + ch);
            const float4* fh = reinterpret_cast<const float4*>(film + C + ch);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) { S[i4] = __ldg(fs + i4); H[i4] = __ldg(fh + i4); }
          }
          tmem_ld_wait();
          uint32_t hw8[8];
#pragma unroll
          for (int i8 = 0; i8 < 2; ++i8) {
            float hv[8];
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
              const int i4 = i8 * 2 + h4;
              float4 A = make_float4(0.f, 0.f, 0.f, 0.f), G = A;            // BIAS_MMA: already in the accumulator
              if (!K::BIAS_MMA) { A = sBA[(ch >> 2) + i4]; G = sNB[(ch >> 2) + i4]; }
              const float av[4] = {A.x, A.y, A.z, A.w}, gv[4] = {G.x, G.y, G.z, G.w};
              const float sv[4] = {S[i4].x, S[i4].y, S[i4].z, S[i4].w}, tv[4] = {H[i4].x, H[i4].y, H[i4].z, H[i4].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // (x + 0.0f is not a no-op for the compiler: -0 + 0 = +0, so the adds are selected away explicitly)
                const float a = K::BIAS_MMA ? __uint_as_float(va[i4 * 4 + e]) : __uint_as_float(va[i4 * 4 + e]) + av[e];   // (conv_a + b_a) / 2
                const float gg = K::BIAS_MMA ? __uint_as_float(vg[i4 * 4 + e]) : __uint_as_float(vg[i4 * 4 + e]) + gv[e];
                const float th = RB3_DBG(64) ? gg : tanh_approx(gg);                        // tanh(g / 2)
                hv[h4 * 4 + e] = fmaf(fmaf(a, th, a), sv[e], tv[e]);                      // a sigmoid(g) (1+scale) + shift
              }
            }
            if (K::TSH) {
              hw8[i8 * 4 + 0] = pack2t<FMT>(hv[0], hv[1]); hw8[i8 * 4 + 1] = pack2t<FMT>(hv[2], hv[3]);
              hw8[i8 * 4 + 2] = pack2t<FMT>(hv[4], hv[5]); hw8[i8 * 4 + 3] = pack2t<FMT>(hv[6], hv[7]);
            } else {
            const int chunk = (cl >> 3) + i8;
            if (!RB3_DBG(256) || hv[0] == 123.456f)
            *reinterpret_cast<uint4*>(hrow + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pack2t<FMT>(hv[0], hv[1]), pack2t<FMT>(hv[2], hv[3]), pack2t<FMT>(hv[4], hv[5]),
                           pack2t<FMT>(hv[6], hv[7]));
            }
          }
          // WS: the 16 channels just computed, as 8 packed columns over the first 8 of the value columns they came from
          // (k-step cl/16 of GEMM2's TMEM A operand)
          if (K::TSH) tmem_st8(lane_addr + b * 128 + cl, hw8);
        }
        if (K::TSH) {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&h_full[b]);
        } else {
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          arrive_leader(&h_full[j]);
          arrive_leader(&d1_empty[b]);
        }
        }
        if (q == 0 && par == 0) RB3_TRACE(3, gc, 2);
      }
    }
  } else if (K::INPLACE) {
    // ------------------------------------------------------------ epilogue 2, in-place variant (C = 128):
    // warp set `par` owns the tiles with it % 2 == par.  The residual x comes from the shared-memory
    // input tile; the output is staged over the same chunks and leaves through TMA stores.
    const int q = warp & 3, par = (warp - 10) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sB2 = reinterpret_cast<const float4*>(sPar + 2 * C);
    for (int it = par; it < n_my_tiles; it += 2) {
      const int tile = tile_of(it);
      const int seq = tile / p.tiles_per_seq, l0 = (tile - seq * p.tiles_per_seq) * 128;
      const int db = it % ND2, ab = it % NA;
      mbar_wait(&a_full[ab], (it / NA) & 1);      // visibility of the TMA-written tile to this thread
      mbar_wait(&d2_full[db], (it / ND2) & 1);
      if (q == 0) RB3_TRACE(4, it, 0);
      tc_fence_after();
      uint8_t* abase = sA + ab * K::A_BYTES + (row + K::HALO) * 128;
      if (!RB3_DBG(2))
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 16) {
        uint32_t vd[16];
        tmem_ld16(lane_addr + K::D2_COL + db * C + c0, vd);
        uint4 xa[2];
        uint4* xp[2];
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          const int kb = c0 >> 6, chunk = ((c0 & 63) >> 3) + i8;
          xp[i8] = reinterpret_cast<uint4*>(abase + kb * K::A_KB_BYTES + ((chunk ^ (row & 7)) << 4));
          xa[i8] = *xp[i8];
        }
        tmem_ld_wait();
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          const uint32_t xw[4] = {xa[i8].x, xa[i8].y, xa[i8].z, xa[i8].w};
          float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;           // BIAS_MMA: already in the accumulator
          if (!K::BIAS_MMA) { B0 = sB2[(c0 >> 2) + i8 * 2]; B1 = sB2[(c0 >> 2) + i8 * 2 + 1]; }
          const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
          uint32_t ow[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 xs = unpack2t<FMT>(xw[e2]);
            float y0 = (K::BIAS_MMA ? lrelu_inv_fast(xs.x) : lrelu_inv_fast(xs.x) + bv[e2 * 2]) + __uint_as_float(vd[i8 * 8 + e2 * 2]);
            float y1 = (K::BIAS_MMA ? lrelu_inv_fast(xs.y) : lrelu_inv_fast(xs.y) + bv[e2 * 2 + 1]) + __uint_as_float(vd[i8 * 8 + e2 * 2 + 1]);
            if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); }
            ow[e2] = pack2t<OFMT>(y0, y1);
          }
          *xp[i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_leader(&d2_empty[db]);
      named_bar_sync(1 + par, 128);
      if (q == 0 && lane == 0) {
#pragma unroll
        for (int kb = 0; kb < KPT; ++kb)
          tma_store_3d(&tmOut, sA + ab * K::A_BYTES + kb * K::A_KB_BYTES + K::HALO * 128, kb * 64, l0, seq);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&a_empty[ab]);
      }
      if (q == 0) RB3_TRACE(4, it, 1);
    }
  } else {
    // ------------------------------------------------------------ epilogue 2 (warps 10..17), C = 256:
    // residual from global (L2), direct stores; the two warps of a quadrant split the channels
    const int q = warp & 3, hsel = (warp - 10) >> 2;
    const int row = q * 32 + lane;
    constexpr int CW = C / 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sB2 = reinterpret_cast<const float4*>(sPar + 2 * C);
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = tile_of(it);
      const int seq = tile / p.tiles_per_seq, l = (tile - seq * p.tiles_per_seq) * 128 + row;
      const bool valid = l < p.L && seq < p.n_seq;          // (pair mode: the tile past the end is out of bounds)
      const int db = it % ND2;
      const long long roff = ((long long)(seq < p.n_seq ? seq : p.n_seq - 1) * p.L + (valid ? l : 0)) * C;
      const uint4* xin = reinterpret_cast<const uint4*>(p.a16 + roff);
      uint4* dst = reinterpret_cast<uint4*>(p.out + roff);
      // the residual operand comes from global (L2): issue the loads BEFORE waiting for the accumulator
      constexpr int XV = CW / 8 < 8 ? CW / 8 : 8;       // uint4 vectors prefetched per pass (<= 32 registers)
#pragma unroll 1
      for (int pass = 0; pass < CW / (XV * 8); ++pass) {
        const int cp = hsel * CW + pass * XV * 8;
        uint4 xa[XV];
#pragma unroll
        for (int v = 0; v < XV; ++v) xa[v] = __ldg(xin + (cp >> 3) + v);
        if (pass == 0) {
          mbar_wait(&d2_full[db], (it / ND2) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int c0 = cp; c0 < cp + XV * 8; c0 += 16) {
          uint32_t vd[16];
          tmem_ld16(lane_addr + K::D2_COL + db * C + c0, vd);
          tmem_ld_wait();
#pragma unroll
          for (int i8 = 0; i8 < 2; ++i8) {
            const uint4 xv = xa[((c0 - cp) >> 3) + i8];
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
            float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;         // BIAS_MMA: already in the accumulator
            if (!K::BIAS_MMA) { B0 = sB2[(c0 >> 2) + i8 * 2]; B1 = sB2[(c0 >> 2) + i8 * 2 + 1]; }
            const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
            uint32_t ow[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const float2 xs = unpack2t<FMT>(xw[e2]);
              float y0 = (K::BIAS_MMA ? lrelu_inv_fast(xs.x) : lrelu_inv_fast(xs.x) + bv[e2 * 2]) + __uint_as_float(vd[i8 * 8 + e2 * 2]);
              float y1 = (K::BIAS_MMA ? lrelu_inv_fast(xs.y) : lrelu_inv_fast(xs.y) + bv[e2 * 2 + 1]) + __uint_as_float(vd[i8 * 8 + e2 * 2 + 1]);
              if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); }
              ow[e2] = pack2t<OFMT>(y0, y1);
            }
            if (valid) dst[(c0 >> 3) + i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&d2_empty[db]);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();       // the peer may still read our shared memory / TMEM until here
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2cta(tmem_base, K::TMEM_COLS); else tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

static int num_sms3() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!n[dev & 15]) cudaDeviceGetAttribute(&n[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  return n[dev & 15];
}

// B200VOC_RB3_PAIR=0 forces the single-CTA kernel (A/B runs); default is the CTA-pair (cta_group::2) kernel.
static bool use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VOC_RB3_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int C, int FMT, int OFMT, bool LRELU, bool PAIR>
static int launch_resblock3_t(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                              const float* film, int film_stride, int N, int L, int dilation, int T, int num_bands,
                              void* out16, cudaStream_t stream) {
  using K = Rb3Cfg<C, PAIR>;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
  B200_TRY(make_tmap_3d(&tmX, a16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, 64, K::A_ROWS, 128));
  B200_TRY(make_tmap_3d(&tmOut, out16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, 64, 128, 128));
  const uint16_t* w1 = reinterpret_cast<const uint16_t*>(w_packed);
  const uint16_t* w2 = w1 + 2ll * C * 3 * C;
  B200_TRY(make_tmap_2d(&tmW1, w1, 3 * C, 2 * C, (uint64_t)3 * C * 2, 64, K::W_ROWS, 128));
  B200_TRY(make_tmap_2d(&tmW2, w2, C, C, (uint64_t)C * 2, 64, K::W_ROWS, 128));
  Resblock3Params p{};
  p.L = L; p.dilation = dilation; p.T = T; p.P = L / T; p.num_bands = num_bands;
  p.tiles_per_seq = ceil_div(L, 128);
  p.total_tiles = p.tiles_per_seq * N;
  p.n_seq = N;
  p.a16 = reinterpret_cast<const uint16_t*>(a16);
  p.b_conv = b_conv; p.b_proj = b_proj; p.film = film; p.film_stride = film_stride;
  p.out = reinterpret_cast<uint16_t*>(out16);
  {
    const char* e = getenv("B200VOC_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.trace = g_rb2_trace;
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  auto kernel = resblock3_kernel<C, FMT, OFMT, LRELU, PAIR>;
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    configured[dev & 15] = true;
  }
  const int sms = num_sms3();
  if (!PAIR) {
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    kernel<<<grid, 608, K::SMEM, stream>>>(tmX, tmW1, tmW2, tmOut, p);
  } else {
    const int units = (p.total_tiles + 1) / 2, pairs = units < sms / 2 ? units : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(608);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmX, tmW1, tmW2, tmOut, p));
  }
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int C, int FMT, int OFMT, bool LRELU>
static int launch_resblock3(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                            const float* film, int film_stride, int N, int L, int dilation, int T, int num_bands,
                            void* out16, cudaStream_t stream) {
  if (use_pair())
    return launch_resblock3_t<C, FMT, OFMT, LRELU, true>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation,
                                                         T, num_bands, out16, stream);
  return launch_resblock3_t<C, FMT, OFMT, LRELU, false>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation,
                                                        T, num_bands, out16, stream);
}

template <int C>
static int dispatch_resblock3(const void* a16, const void* w, const float* bc, const float* bp, const float* film,
                              int fs, int N, int L, int d, int T, int nb, int fmt, int ofmt, int lrelu, void* out,
                              cudaStream_t st) {
#define RB3(F, O, R) return launch_resblock3<C, F, O, R>(a16, w, bc, bp, film, fs, N, L, d, T, nb, out, st)
  if (fmt == 0) {
    if (ofmt == 0) { if (lrelu) RB3(0, 0, true); else RB3(0, 0, false); }
    else { if (lrelu) RB3(0, 1, true); else RB3(0, 1, false); }
  } else {
    if (ofmt == 0) { if (lrelu) RB3(1, 0, true); else RB3(1, 0, false); }
    else { if (lrelu) RB3(1, 1, true); else RB3(1, 1, false); }
  }
#undef RB3
}

int resblock3_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(dilation >= 1 && dilation <= 8, "resblock3: dilation %d exceeds the 8-row halo", dilation);
  B200_CHECK_ARG((fmt == 0 || fmt == 1) && (out_fmt == 0 || out_fmt == 1), "resblock3: bad format");
  if (C == 128)
    return dispatch_resblock3<128>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                   out_fmt, store_lrelu, out16, stream);
  if (C == 256)
    return dispatch_resblock3<256>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                   out_fmt, store_lrelu, out16, stream);
  set_error("resblock3: C=%d unsupported (128/256)", C);
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
