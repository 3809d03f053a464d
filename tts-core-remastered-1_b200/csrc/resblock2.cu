// K2b: persistent fused residual block for the narrow stages (C = 32, 64).
//
// Same math as resblock.cu (generator.py:40-41,89-90 / repair R2), different schedule.  The late
// stages move 0.9 GB in and 0.9 GB out per block at only 112-224 flop/B, and a 128-row tile holds
// just 4096-8192 elements: measured (profiles/r01_*), the kernel is bound by the number of warp
// instructions issued per tile -- epilogue math plus the fixed per-tile cost of every warp's
// pipeline bookkeeping -- not by the tensor pipe (10-20 tiny MMAs per tile), TMEM reads
// (~690 B/clk/SM measured) or HBM.  So everything linear runs on the tensor core, the epilogues are
// minimal, and barrier traffic is per warp, not per thread:
//   * activations of the narrow stages are stored RAW (x, not leaky_relu(x)); one TMA load per tile
//     brings 144 rows (128 + 8-row halo each side) into the R ring (6-12 slots: the prefetch
//     distance that covers the HBM latency); a slot is held only until GEMM1 has read it;
//   * the residual add is an MMA issued FIRST: D2 = I * x_centre (exact: 1.0 * x in fp32), by a
//     dedicated issuer warp that runs ahead as far as the D2 ring allows; x then lives in TMEM;
//   * 4 "prep" warps then apply leaky_relu IN PLACE in packed 16-bit arithmetic (2 elements per
//     instruction, conflict-free linear pass); GEMM1's three dilated taps are UMMA descriptors
//     over the slot shifted by (8 +- d) rows; GEMM1's commit hands the slot back to the TMA warp;
//   * W1 and the conv bias are pre-scaled by 1/2 at pack time (exact in fp16/bf16), so that
//     GLU + FiLM is  h = fma(fma(a', tanh(g'), a'), S, T)  (sigmoid(g) = 1/2 + 1/2 tanh(g/2));
//   * GEMM2 accumulates W2 * h onto the x already in D2; epilogue 2 is TMEM -> +bias -> pack ->
//     staged over the (now dead) h slot -> one TMA store, whose shared-memory read completion is
//     awaited one tile later (off the critical path) before the h slot is recycled;
//   * W1 (3 taps), W2 and I stay resident in shared memory; the roles are specialised warps:
//     TMA | 2 GEMM issuers (even/odd tiles) | identity-MMA issuer | 4 prep | 4 store | 8 GLU+FiLM;
//   * FiLM coefficients of a tile (one frame when P % 128 == 0) are staged per warp in shared memory;
//   * multi-thread barriers are arrived on once per WARP (fence, __syncwarp, lane 0 arrives): a
//     128-arrival barrier wakes its waiter ~25 times per phase, which alone was ~20 % of all
//     issued instructions.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct Resblock2Params {
  int L, dilation, T, P, num_bands;
  int tiles_per_seq, total_tiles;
  const float* b_conv;   // [2C]
  const float* b_proj;   // [C]
  const float* film;     // [B, T, film_stride]
  int film_stride;
  long long* trace;      // clock64 timeline of CTA 0, [7 roles][64 tiles][4] (only with -DB200VOC_TRACE)
};

#ifdef B200VOC_TRACE
#define RB2_TRACE(slot, i, k)                                                             \
  do {                                                                                    \
    if (p.trace && blockIdx.x == 0 && (i) < 64 && (threadIdx.x & 31) == 0)                 \
      p.trace[(((slot) * 64 + (i)) << 2) + (k)] = clock64();                               \
  } while (0)
#else
#define RB2_TRACE(slot, i, k) do { } while (0)
#endif

template <int C>
struct Rb2Cfg {
  static constexpr int KB = C;                      // C in {32, 64}: one k-block per tap
  static constexpr int ROWB = KB * 2;               // 64 or 128 bytes per row = swizzle span
  static constexpr int N1 = 2 * C;
  static constexpr int HALO = 8;
  static constexpr int A_ROWS = 128 + 2 * HALO;
  static constexpr int A_BYTES = A_ROWS * ROWB;
  static constexpr int A_SLOT = (A_BYTES + 1023) & ~1023;
  static constexpr int W1_TILE = N1 * ROWB;
  static constexpr int W2_TILE = C * ROWB;
  static constexpr int H_BYTES = 128 * ROWB;
  static constexpr int ND1 = C == 32 ? 4 : 2;        // ring depth of the GEMM1 accumulators
  static constexpr int ND2 = C == 32 ? 8 : 4;        // ring depth of the GEMM2 accumulators (hold x from the identity MMA on)
  static constexpr int NH = C == 32 ? 6 : 3;         // ring depth of h (also the store staging)
  static constexpr int NR = C == 32 ? 12 : 6;        // input tiles (TMA prefetch distance)
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = 3 * W1_TILE;
  static constexpr int OFF_ID = OFF_W2 + W2_TILE;
  static constexpr int OFF_R = (OFF_ID + W2_TILE + 1023) & ~1023;
  static constexpr int OFF_H = OFF_R + NR * A_SLOT;
  static constexpr int OFF_BAR = OFF_H + NH * H_BYTES;
  static constexpr int OFF_PAR = OFF_BAR + 1024;
  static constexpr bool SPLIT = C == 64;             // GLU epilogue: both warp sets share a tile (column halves)
  static constexpr int E1_CH = SPLIT ? C / 2 : C;    // channels per GLU warp
  static constexpr int OFF_FILM = OFF_PAR + 3 * C * 4;     // per GLU warp: (1+scale | shift) of its channels, 2*E1_CH floats
  static constexpr int SMEM = OFF_FILM + 8 * 2 * E1_CH * 4 + 1024;
  static constexpr int NBARS = 1 + 4 * NR + 2 * ND1 + 2 * NH + 2 * ND2;
  static constexpr int D2_COL = ND1 * N1;
  static constexpr int TMEM_NEED = ND1 * N1 + ND2 * C;
  static_assert(NBARS * 8 + 8 <= 1024, "barrier block");
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(OFF_W2 % 1024 == 0 && OFF_ID % 1024 == 0 && A_SLOT % 1024 == 0 && H_BYTES % 1024 == 0, "swizzle alignment");
  static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
};

// (sequence, first row) of the tiles a persistent CTA walks, without a division per tile.
// `stride` = tiles between two visits (gridDim.x, or 2 * gridDim.x for a role that owns every other tile).
struct TileWalk {
  int seq, l0, step, span;
  __device__ __forceinline__ TileWalk(int tile0, int tiles_per_seq, int stride) {
    seq = tile0 / tiles_per_seq;
    l0 = (tile0 - seq * tiles_per_seq) * 128;
    step = stride * 128;
    span = tiles_per_seq * 128;
  }
  __device__ __forceinline__ void next() {
    l0 += step;
    while (l0 >= span) { l0 -= span; ++seq; }
  }
};

// The same walk plus the FiLM coordinates of the tile's first row -- frame t0 = l0 / P, offset
// rem0 = l0 % P, batch element b = seq / num_bands -- all kept incrementally (the per-tile integer
// divisions were a measurable part of the GLU epilogue's fixed cost).
struct FilmWalk {
  int seq, l0, t0, rem0, b, sb;
  int step, span, P, nb, qstep, rstep, qspan, rspan;
  __device__ __forceinline__ FilmWalk(int tile0, int tiles_per_seq, int stride, int P_, int nb_) {
    P = P_; nb = nb_;
    seq = tile0 / tiles_per_seq;
    l0 = (tile0 - seq * tiles_per_seq) * 128;
    step = stride * 128;
    span = tiles_per_seq * 128;
    t0 = l0 / P; rem0 = l0 - t0 * P;
    b = seq / nb; sb = seq - b * nb;
    qstep = step / P; rstep = step - qstep * P;
    qspan = span / P; rspan = span - qspan * P;
  }
  __device__ __forceinline__ void next() {
    l0 += step; t0 += qstep; rem0 += rstep;
    if (rem0 >= P) { rem0 -= P; ++t0; }
    while (l0 >= span) {
      l0 -= span; ++seq;
      t0 -= qspan; rem0 -= rspan;
      if (rem0 < 0) { rem0 += P; --t0; }
      if (++sb == nb) { sb = 0; ++b; }
    }
  }
};

// leaky_relu on two packed 16-bit values: max(x, 0.1 x).  The slope constant is rounded to the
// storage format (fp16: 2.4e-4 relative), which only touches the negative side, i.e. values
// already scaled down 10x -- 20x below the rounding step of the stored activation itself.
template <int FMT>
__device__ __forceinline__ uint32_t lrelu2_packed(uint32_t w) {
  if constexpr (FMT == 0) {
    const __half2 x = *reinterpret_cast<const __half2*>(&w);
    const __half2 y = __hmax2(x, __hmul2(x, __float2half2_rn(kLreluSlope)));
    return *reinterpret_cast<const uint32_t*>(&y);
  } else {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&w);
    const __nv_bfloat162 y = __hmax2(x, __hmul2(x, __float2bfloat162_rn(kLreluSlope)));
    return *reinterpret_cast<const uint32_t*>(&y);
  }
}

template <int C, int FMT, int OFMT, bool LRELU>
__global__ void __launch_bounds__(640, 1)
resblock2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                 const Resblock2Params p) {
  using K = Rb2Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sW1 = smem + K::OFF_W1;
  uint8_t* sW2 = smem + K::OFF_W2;
  uint8_t* sID = smem + K::OFF_ID;
  uint8_t* sR = smem + K::OFF_R;
  uint8_t* sH = smem + K::OFF_H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  constexpr int ND1 = K::ND1, ND2 = K::ND2, NR = K::NR, NH = K::NH;
  uint64_t* w_full = bars;                 // [1]
  uint64_t* r_full = bars + 1;             // [NR]  TMA bytes landed            -> identity issuer, prep
  uint64_t* x_done = r_full + NR;          // [NR]  identity MMA has read x     -> prep
  uint64_t* p_full = x_done + NR;          // [NR]  leaky_relu applied in place -> GEMM issuers
  uint64_t* r_empty = p_full + NR;         // [NR]  GEMM1 has read the slot     -> TMA
  uint64_t* d1_full = r_empty + NR;        // [ND1]
  uint64_t* d1_empty = d1_full + ND1;      // [ND1]
  uint64_t* h_full = d1_empty + ND1;       // [NH]
  uint64_t* h_empty = h_full + NH;         // [NH]  the tile's TMA store has read the slot -> epilogue 1
  uint64_t* d2_full = h_empty + NH;        // [ND2]
  uint64_t* d2_empty = d2_full + ND2;      // [ND2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + ND2);
  float* sPar = reinterpret_cast<float*>(smem + K::OFF_PAR);   // [ba/2 | bg/2 | b2], C floats each

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kSetWarps = 4;              // the 4 warps (one per TMEM lane quadrant) that share a tile
  constexpr int kSetThreads = 128;

  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sPar[i] = 0.5f * p.b_conv[i];
    sPar[C + i] = 0.5f * p.b_conv[C + i];
    sPar[2 * C + i] = p.b_proj[i];
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    mbar_init(w_full, 1);
    for (int b = 0; b < NR; ++b) {
      mbar_init(&r_full[b], 1);
      mbar_init(&x_done[b], 1);
      mbar_init(&p_full[b], kSetWarps);
      mbar_init(&r_empty[b], 1);
    }
    for (int b = 0; b < ND1; ++b) {
      mbar_init(&d1_full[b], 1);
      mbar_init(&d1_empty[b], K::SPLIT ? 2 * kSetWarps : kSetWarps);
    }
    for (int b = 0; b < NH; ++b) {
      mbar_init(&h_full[b], K::SPLIT ? 2 * kSetWarps : kSetWarps);
      mbar_init(&h_empty[b], 1);
    }
    for (int b = 0; b < ND2; ++b) {
      mbar_init(&d2_full[b], 1);
      mbar_init(&d2_empty[b], kSetWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, K::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int grid = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(w_full, 3 * K::W1_TILE + 2 * K::W2_TILE);
      for (int tap = 0; tap < 3; ++tap) tma_load_2d(sW1 + tap * K::W1_TILE, &tmW1, w_full, tap * C, 0);
      tma_load_2d(sW2, &tmW2, w_full, 0, 0);
      tma_load_2d(sID, &tmW2, w_full, C, 0);
      TileWalk tw(blockIdx.x, p.tiles_per_seq, grid);
      int i = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++i, tw.next()) {
        const int rb = i % NR;
        mbar_wait(&r_empty[rb], ((i / NR) & 1) ^ 1);
        RB2_TRACE(0, i, 0);
        mbar_expect_tx(&r_full[rb], K::A_BYTES);
        tma_load_3d(sR + rb * K::A_SLOT, &tmX, &r_full[rb], 0, tw.l0 - K::HALO, tw.seq);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ identity-MMA issuer: D2[i] = I * x_centre(i)
    // (the residual).  Runs ahead of the GEMM issuers as far as the D2 ring allows; once it has
    // completed, the slot may be leaky_relu'ed in place.
    const uint32_t idesc2 = make_idesc_f16(FMT, C);
    mbar_wait(w_full, 0);
    const uint64_t id_desc = make_kmajor_desc<K::ROWB>(smem_u32(sID));
    int i = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++i) {
      const int rb = i % NR, b2 = i % ND2;
      const uint32_t phr = (i / NR) & 1, ph2 = ((i / ND2) & 1) ^ 1;
      const bool r1 = mbar_test(&r_full[rb], phr), r2 = mbar_test(&d2_empty[b2], ph2);
      if (!r1) mbar_wait(&r_full[rb], phr);
      RB2_TRACE(5, i, 0);
      if (!r2) mbar_wait(&d2_empty[b2], ph2);
      RB2_TRACE(5, i, 1);
      tc_fence_after();
      const uint64_t x_desc = make_kmajor_desc<K::ROWB>(smem_u32(sR + rb * K::A_SLOT) + K::HALO * K::ROWB);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < K::KB / 16; ++k)
          umma_f16(tmem_base + K::D2_COL + b2 * C, x_desc + 2 * k, id_desc + 2 * k, idesc2, k != 0);
        umma_commit(&x_done[rb]);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ GEMM issuers: warp 1 owns the even tiles,
    // warp 2 the odd ones (tcgen05.mma may be issued by any thread).  The whole warp runs the
    // (warp-uniform) control flow, one elected lane issues tcgen05.mma / commit: indices and
    // descriptors stay in uniform registers; independent barrier probes are issued together so
    // their round trips overlap.
    const int mpar = warp - 1;
    const uint32_t idesc1 = make_idesc_f16(FMT, K::N1);
    const uint32_t idesc2 = make_idesc_f16(FMT, C);
    mbar_wait(w_full, 0);
    auto issue_g2 = [&](int i) {
      const int hb = i % NH, b2 = i % ND2;
      mbar_wait(&h_full[hb], (i / NH) & 1);
      RB2_TRACE(2, i, 0);
      tc_fence_after();
      const uint64_t h_desc = make_kmajor_desc<K::ROWB>(smem_u32(sH + hb * K::H_BYTES));
      const uint64_t w2_desc = make_kmajor_desc<K::ROWB>(smem_u32(sW2));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < K::KB / 16; ++k)
          umma_f16(tmem_base + K::D2_COL + b2 * C, h_desc + 2 * k, w2_desc + 2 * k, idesc2, 1);   // x + W2 h
        umma_commit(&d2_full[b2]);
      }
      __syncwarp();
    };
    int prev = -1;
    for (int i = mpar; blockIdx.x + (long long)i * grid < p.total_tiles; i += 2) {
      const int b = i % ND1, rb = i % NR;
      const uint32_t php = (i / NR) & 1, phd = ((i / ND1) & 1) ^ 1;
      const bool ra = mbar_test(&p_full[rb], php), rd = mbar_test(&d1_empty[b], phd);
      if (!ra) mbar_wait(&p_full[rb], php);
      RB2_TRACE(1, i, 0);
      if (!rd) mbar_wait(&d1_empty[b], phd);
      RB2_TRACE(1, i, 1);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sR + rb * K::A_SLOT);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
          const uint64_t a_desc = make_kmajor_desc<K::ROWB>(a_base + (K::HALO + (tap - 1) * p.dilation) * K::ROWB);
          const uint64_t b_desc = make_kmajor_desc<K::ROWB>(smem_u32(sW1 + tap * K::W1_TILE));
#pragma unroll
          for (int k = 0; k < K::KB / 16; ++k)
            umma_f16(tmem_base + b * K::N1, a_desc + 2 * k, b_desc + 2 * k, idesc1, (tap | k) != 0);
        }
        umma_commit(&d1_full[b]);
        umma_commit(&r_empty[rb]);
      }
      __syncwarp();
      RB2_TRACE(1, i, 2);
      if (prev >= 0) issue_g2(prev);        // GEMM2 trails GEMM1 by one own tile (= 2 tiles)
      prev = i;
    }
    if (prev >= 0) issue_g2(prev);
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue 1 (warps 12..19): GLU + FiLM -> h
    // Two warps per TMEM lane quadrant.  C = 32: warp set `set` owns the tiles with i % 2 == set, so
    // two tiles' GLU epilogues are in flight per quadrant.  C = 64 (only two GEMM1 accumulators fit
    // in TMEM next to the D2 ring): both sets work on EVERY tile, one half of the channels each,
    // which halves the time an accumulator is held.
    const int q = warp & 3, set = (warp - 12) >> 2;
    const int row = q * 32 + lane;
    constexpr int CH = K::E1_CH;
    const int cbeg = K::SPLIT ? set * CH : 0;
    const int first = K::SPLIT ? 0 : set, stride = K::SPLIT ? 1 : 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* scratch = reinterpret_cast<float*>(smem + K::OFF_FILM) + (warp - 12) * 2 * CH;   // [S(CH) | T(CH)]
    FilmWalk fw(blockIdx.x + first * grid, p.tiles_per_seq, stride * grid, p.P, p.num_bands);
    for (int i = first; blockIdx.x + (long long)i * grid < p.total_tiles; i += stride, fw.next()) {
      const int b = i % ND1, hb = i % NH;
      // FiLM: when all 128 rows of the tile lie in one frame (always, in the generator: P is a
      // multiple of 128) the frame's coefficients are staged once per warp in shared memory and
      // read back as broadcasts; a warp-wide LDG.128 of one address costs 4-5 L1 wavefronts, and
      // 2 per 4 elements of them made the L1 data pipe the busiest unit of this kernel.
      const bool uniform = fw.rem0 + 127 < fw.P;
      float4 stage = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* film_row;
      if (uniform) {
        film_row = p.film + ((long long)fw.b * p.T + fw.t0) * p.film_stride;
        if (lane < CH / 2) {                       // lanes [0, CH/4): S, [CH/4, CH/2): T
          const int j = lane < CH / 4 ? lane : lane - CH / 4;
          stage = __ldg(reinterpret_cast<const float4*>(film_row + (lane < CH / 4 ? 0 : C) + cbeg) + j);
        }
      } else {
        int t = fw.t0 + (fw.rem0 + row) / fw.P;
        if (t > p.T - 1) t = p.T - 1;
        film_row = p.film + ((long long)fw.b * p.T + t) * p.film_stride;
      }
      mbar_wait(&d1_full[b], (i / ND1) & 1);
      if (q == 0) RB2_TRACE(3, i, 0);
      mbar_wait(&h_empty[hb], ((i / NH) & 1) ^ 1);
      if (q == 0) RB2_TRACE(3, i, 1);
      tc_fence_after();
      if (uniform) {
        if (lane < CH / 2) reinterpret_cast<float4*>(scratch)[lane] = stage;
        __syncwarp();
      }
      uint8_t* hrow = sH + hb * K::H_BYTES + row * K::ROWB;
      const float* srcS = film_row + cbeg;
      const float* srcT = film_row + C + cbeg;
#pragma unroll
      for (int cc = 0; cc < CH; cc += 16) {
        const int c0 = cbeg + cc;
        uint32_t va[16], vg[16];
        tmem_ld16(lane_addr + b * K::N1 + c0, va);
        tmem_ld16(lane_addr + b * K::N1 + C + c0, vg);
        float4 S[4], H[4];
        if (uniform) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            S[i4] = reinterpret_cast<const float4*>(scratch + cc)[i4];
            H[i4] = reinterpret_cast<const float4*>(scratch + CH + cc)[i4];
          }
        } else {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            S[i4] = __ldg(reinterpret_cast<const float4*>(srcS + cc) + i4);
            H[i4] = __ldg(reinterpret_cast<const float4*>(srcT + cc) + i4);
          }
        }
        tmem_ld_wait();
        if (q == 0 && cc == 0) RB2_TRACE(3, i, 3);
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          float hv[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const int i4 = i8 * 2 + h4;
            const float4 A = *reinterpret_cast<const float4*>(sPar + c0 + i4 * 4);
            const float4 G = *reinterpret_cast<const float4*>(sPar + C + c0 + i4 * 4);
            const float av[4] = {A.x, A.y, A.z, A.w}, gv[4] = {G.x, G.y, G.z, G.w};
            const float sv[4] = {S[i4].x, S[i4].y, S[i4].z, S[i4].w}, tv[4] = {H[i4].x, H[i4].y, H[i4].z, H[i4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(va[i4 * 4 + e]) + av[e];                  // (conv_a + b_a) / 2
              const float th = tanh_approx(__uint_as_float(vg[i4 * 4 + e]) + gv[e]);    // tanh(g / 2)
              hv[h4 * 4 + e] = fmaf(fmaf(a, th, a), sv[e], tv[e]);                      // a * sigmoid(g) * (1+scale) + shift
            }
          }
          const int chunk = (c0 >> 3) + i8;
          const int phys = K::ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
          *reinterpret_cast<uint4*>(hrow + phys * 16) =
              make_uint4(pack2t<FMT>(hv[0], hv[1]), pack2t<FMT>(hv[2], hv[3]), pack2t<FMT>(hv[4], hv[5]),
                         pack2t<FMT>(hv[6], hv[7]));
        }
        if (q == 0 && cc == 0) RB2_TRACE(6, i, 2);
      }
      if (q == 0) RB2_TRACE(6, i, 3);
      tc_fence_before();             // TMEM reads of D1 are complete
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the UMMA (async proxy)
      __syncwarp();                  // (also: every lane is done with the FiLM scratch)
      if (lane == 0) {
        mbar_arrive(&h_full[hb]);
        mbar_arrive(&d1_empty[b]);
      }
      if (q == 0) RB2_TRACE(3, i, 2);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ prep (warps 4..7): slot <- leaky_relu(slot)
    const int tid = threadIdx.x - 128;
    constexpr int NV = K::A_BYTES / 16;                       // 16-byte vectors per tile (576 / 1152)
    constexpr int PER = (NV + kSetThreads - 1) / kSetThreads;
    constexpr int HALF = (PER + 1) / 2;
    int i = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++i) {
      const int rb = i % NR;
      mbar_wait(&r_full[rb], (i / NR) & 1);     // acquire the TMA writes
      mbar_wait(&x_done[rb], (i / NR) & 1);     // the identity MMA has read the raw rows
      if (tid == 0) RB2_TRACE(6, i, 0);
      tc_fence_after();
      uint4* slot = reinterpret_cast<uint4*>(sR + rb * K::A_SLOT);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 v[HALF];
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const int idx = tid + (h * HALF + j) * kSetThreads;
          if (idx < NV) v[j] = slot[idx];
        }
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
          const int idx = tid + (h * HALF + j) * kSetThreads;
          if (idx < NV)
            slot[idx] = make_uint4(lrelu2_packed<FMT>(v[j].x), lrelu2_packed<FMT>(v[j].y), lrelu2_packed<FMT>(v[j].z),
                                   lrelu2_packed<FMT>(v[j].w));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[rb]);
      if (tid == 0) RB2_TRACE(6, i, 1);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------ epilogue 2 (warps 8..11): + bias, store
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sB2 = reinterpret_cast<const float4*>(sPar + 2 * C);
    const bool storer = warp == 8 && lane == 0;
    TileWalk tw(blockIdx.x, p.tiles_per_seq, grid);
    int i = 0, pending_hb = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++i, tw.next()) {
      const int b = i % ND2, hb = i % NH;
      mbar_wait(&d2_full[b], (i / ND2) & 1);
      if (q == 0) RB2_TRACE(4, i, 0);
      tc_fence_after();
      // The output tile is staged over the tile's h slot (its last reader, GEMM2, has completed:
      // d2_full) and leaves through one TMA store: no strided per-thread global stores, and rows
      // past the end of the sequence are clipped by TMA.
      uint8_t* orow = sH + hb * K::H_BYTES + row * K::ROWB;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t vd[32];
        tmem_ld32(lane_addr + K::D2_COL + b * C + c0, vd);
        tmem_ld_wait();
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const float4 B0 = sB2[(c0 >> 2) + i8 * 2], B1 = sB2[(c0 >> 2) + i8 * 2 + 1];
          const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
          uint32_t ow[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            float y0 = __uint_as_float(vd[i8 * 8 + e2 * 2]) + bv[e2 * 2];
            float y1 = __uint_as_float(vd[i8 * 8 + e2 * 2 + 1]) + bv[e2 * 2 + 1];
            if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); }
            ow[e2] = pack2t<OFMT>(y0, y1);
          }
          const int chunk = (c0 >> 3) + i8;
          const int phys = K::ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
          *reinterpret_cast<uint4*>(orow + phys * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
      tc_fence_before();                            // TMEM reads of D2 are complete
      fence_proxy_async_smem();                     // staged tile -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) mbar_arrive(&d2_empty[b]);
      named_bar_sync(1, kSetThreads);
      if (storer) {
        tma_store_3d(&tmOut, sH + hb * K::H_BYTES, 0, tw.l0, tw.seq);
        tma_store_commit();
        if (pending_hb >= 0) {
          tma_store_wait_read_1();                  // the PREVIOUS tile's store has read its slot
          mbar_arrive(&h_empty[pending_hb]);
        }
        pending_hb = hb;
      }
      if (q == 0) RB2_TRACE(4, i, 1);
    }
    if (storer) {
      if (pending_hb >= 0) {
        tma_store_wait_read();
        mbar_arrive(&h_empty[pending_hb]);
      }
      tma_store_wait_all();                         // global writes complete before the CTA retires
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

long long* g_rb2_trace = nullptr;   // set through b200voc_debug_set_trace (used by -DB200VOC_TRACE builds)

static int num_sms() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!n[dev & 15]) cudaDeviceGetAttribute(&n[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  return n[dev & 15];
}

template <int C, int FMT, int OFMT, bool LRELU>
static int launch_resblock2(const void* x16, const void* w_packed, const float* b_conv, const float* b_proj,
                            const float* film, int film_stride, int N, int L, int dilation, int T, int num_bands,
                            void* out16, cudaStream_t stream) {
  using K = Rb2Cfg<C>;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
  B200_TRY(make_tmap_3d(&tmX, x16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, K::KB, K::A_ROWS, K::ROWB));
  B200_TRY(make_tmap_3d(&tmOut, out16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, K::KB, 128, K::ROWB));
  const uint16_t* w1 = reinterpret_cast<const uint16_t*>(w_packed);
  const uint16_t* w2 = w1 + 2ll * C * 3 * C;                      // [C][2C] = [W2 | I]
  B200_TRY(make_tmap_2d(&tmW1, w1, 3 * C, 2 * C, (uint64_t)3 * C * 2, K::KB, K::N1, K::ROWB));
  B200_TRY(make_tmap_2d(&tmW2, w2, 2 * C, C, (uint64_t)2 * C * 2, K::KB, C, K::ROWB));
  Resblock2Params p{};
  p.L = L; p.dilation = dilation; p.T = T; p.P = L / T; p.num_bands = num_bands;
  p.tiles_per_seq = ceil_div(L, 128);
  p.total_tiles = p.tiles_per_seq * N;
  p.b_conv = b_conv; p.b_proj = b_proj; p.film = film; p.film_stride = film_stride;
  p.trace = g_rb2_trace;
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(resblock2_kernel<C, FMT, OFMT, LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   K::SMEM));
    configured[dev & 15] = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  resblock2_kernel<C, FMT, OFMT, LRELU><<<grid, 640, K::SMEM, stream>>>(tmX, tmW1, tmW2, tmOut, p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int C>
static int dispatch_resblock2(const void* a16, const void* w, const float* bc, const float* bp, const float* film,
                              int fs, int N, int L, int d, int T, int nb, int fmt, int ofmt, int lrelu, void* out,
                              cudaStream_t st) {
#define RB2(F, O, R) return launch_resblock2<C, F, O, R>(a16, w, bc, bp, film, fs, N, L, d, T, nb, out, st)
  if (fmt == 0) {
    if (ofmt == 0) { if (lrelu) RB2(0, 0, true); else RB2(0, 0, false); }
    else { if (lrelu) RB2(0, 1, true); else RB2(0, 1, false); }
  } else {
    if (ofmt == 0) { if (lrelu) RB2(1, 0, true); else RB2(1, 0, false); }
    else { if (lrelu) RB2(1, 1, true); else RB2(1, 1, false); }
  }
#undef RB2
}

// x16 holds RAW x (narrow-stage storage convention); store_lrelu selects leaky_relu(y) for the output.
int resblock2_launch(const void* x16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(dilation >= 1 && dilation <= 8, "resblock2: dilation %d exceeds the 8-row halo", dilation);
  B200_CHECK_ARG((fmt == 0 || fmt == 1) && (out_fmt == 0 || out_fmt == 1), "resblock2: bad format");
  if (C == 32)
    return dispatch_resblock2<32>(x16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                  out_fmt, store_lrelu, out16, stream);
  if (C == 64)
    return dispatch_resblock2<64>(x16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                  out_fmt, store_lrelu, out16, stream);
  set_error("resblock2: C=%d unsupported (32/64)", C);
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
