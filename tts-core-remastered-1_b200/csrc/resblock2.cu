// K2b: persistent fused residual block for the narrow stages (C = 32, 64).
//
// Same math as resblock.cu (generator.py:40-41,89-90 / repair R2), different schedule.  The late
// stages are bandwidth/epilogue bound (224 -> 112 flop/B), so this kernel is organised to stream:
//   * one persistent CTA per SM loops over (sequence, 128-row) tiles;
//   * W1 (3 taps) and W2 stay resident in shared memory for the whole kernel (<= 56 KB);
//   * ONE TMA load per tile brings 144 rows (128 + 8-row halo each side) of leaky_relu(x); the
//     three dilated taps are the same tile addressed through UMMA descriptors whose start address
//     is shifted by (8 +- d) rows -- legal because the 128B/64B swizzle is a function of the
//     absolute shared-memory address (tests/test_gpu_generator.py::test_rowshifted_umma_descriptors);
//   * the residual x is recovered from the same shared-memory tile (no second global read);
//   * A tiles, the h operand and both TMEM accumulators are ring buffered, and the work is
//     split over specialised warps: TMA producer | MMA issuer | 8 warps GLU+FiLM epilogue |
//     8 warps residual+store epilogue, so tile i+1's GEMM1 and GLU overlap tile i's GEMM2/store;
//     operand format / output format / stored activation are template parameters (no per-element
//     branches in the epilogues, which are the instruction-issue bottleneck of these stages).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct Resblock2Params {
  int L, dilation, T, P, num_bands, fmt, out_fmt, store_lrelu;
  int tiles_per_seq, total_tiles;
  const float* b_conv;   // [2C]
  const float* b_proj;   // [C]
  const float* film;     // [B, T, film_stride]
  int film_stride;
  uint16_t* out;         // [N, L, C]
  long long* trace;      // optional clock64 timeline of CTA 0 (debug), [5 roles][64 tiles][4]
  int dbg;               // debug-only experiment switches (B200VOC_DBG): 1 = no FiLM loads, 2 = no E2 stores
};

#define RB2_TRACE(slot, i, k)                                                             \
  do {                                                                                    \
    if (p.trace && blockIdx.x == 0 && (i) < 64 && (threadIdx.x & 31) == 0)                 \
      p.trace[(((slot) * 64 + (i)) << 2) + (k)] = clock64();                               \
  } while (0)

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
  static constexpr int W2_BYTES = C * ROWB;
  static constexpr int H_BYTES = 128 * ROWB;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = 3 * W1_TILE;
  static constexpr int OFF_A = (OFF_W2 + W2_BYTES + 1023) & ~1023;
  static constexpr int ND1 = C == 32 ? 6 : 3;        // ring depth of the GEMM1 accumulators and of h
  static constexpr int ND2 = C == 32 ? 4 : 2;        // ring depth of the GEMM2 accumulators
  static constexpr int NA = C == 32 ? 10 : ND1 + 2;  // ring depth of the input tiles (TMA prefetch distance)
  static constexpr int OFF_H = OFF_A + NA * A_SLOT;
  static constexpr int OFF_BAR = OFF_H + ND1 * H_BYTES;
  static constexpr int OFF_PAR = OFF_BAR + 512;
  static constexpr int SMEM = OFF_PAR + 3 * C * 4 + 1024;
  static constexpr int D2_COL = ND1 * N1;
  static constexpr int TMEM_NEED = ND1 * N1 + ND2 * C;
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
};

template <int C, int FMT, int OFMT, bool LRELU>
__global__ void __launch_bounds__(608, 1)
resblock2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                 const Resblock2Params p) {
  using K = Rb2Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sW1 = smem + K::OFF_W1;
  uint8_t* sW2 = smem + K::OFF_W2;
  uint8_t* sA = smem + K::OFF_A;
  uint8_t* sH = smem + K::OFF_H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  constexpr int ND1 = K::ND1, ND2 = K::ND2, NA = K::NA;
  uint64_t* w_full = bars;                 // [1]
  uint64_t* a_full = bars + 1;             // [NA]
  uint64_t* a_empty = a_full + NA;         // [NA]
  uint64_t* d1_full = a_empty + NA;        // [ND1]
  uint64_t* d1_empty = d1_full + ND1;      // [ND1]
  uint64_t* h_full = d1_empty + ND1;       // [ND1]
  uint64_t* h_empty = h_full + ND1;        // [ND1]
  uint64_t* d2_full = h_empty + ND1;       // [ND2]
  uint64_t* d2_empty = d2_full + ND2;      // [ND2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + ND2);
  float* sPar = reinterpret_cast<float*>(smem + K::OFF_PAR);   // [ba | bg/2 | b2], C floats each

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kEpiThreads = 128;          // per tile: the 4 warps (one per TMEM lane quadrant) of one parity set

  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sPar[i] = p.b_conv[i];
    sPar[C + i] = 0.5f * p.b_conv[C + i];
    sPar[2 * C + i] = p.b_proj[i];
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    mbar_init(w_full, 1);
    for (int b = 0; b < NA; ++b) {
      mbar_init(&a_full[b], 1);
      mbar_init(&a_empty[b], 1);
    }
    for (int b = 0; b < ND1; ++b) {
      mbar_init(&d1_full[b], 1);
      mbar_init(&d1_empty[b], kEpiThreads);
      mbar_init(&h_full[b], kEpiThreads);
      mbar_init(&h_empty[b], 1);
    }
    for (int b = 0; b < ND2; ++b) {
      mbar_init(&d2_full[b], 1);
      mbar_init(&d2_empty[b], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, K::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(w_full, 3 * K::W1_TILE + K::W2_BYTES);
      for (int tap = 0; tap < 3; ++tap) tma_load_2d(sW1 + tap * K::W1_TILE, &tmW1, w_full, tap * C, 0);
      tma_load_2d(sW2, &tmW2, w_full, 0, 0);
      int i = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
        const int ab = i % NA;
        const int seq = tile / p.tiles_per_seq, l0 = (tile - seq * p.tiles_per_seq) * 128;
        mbar_wait(&a_empty[ab], ((i / NA) & 1) ^ 1);
        RB2_TRACE(0, i, 0);
        mbar_expect_tx(&a_full[ab], K::A_BYTES);
        tma_load_3d(sA + ab * K::A_SLOT, &tmX, &a_full[ab], 0, l0 - K::HALO, seq);
      }
    }
  } else if (warp == 1 || warp == 18) {
    // ------------------------------------------------------------ MMA issuers: warp 1 owns the even tiles,
    // warp 18 the odd ones (tcgen05.mma may be issued by any thread; the per-tile barrier round trips
    // of a single issuing thread were the bottleneck of the C=32 stage, see profiles/).
    // Whole warp runs the (warp-uniform) control flow, one elected lane issues tcgen05.mma / commit:
    // indices and descriptors stay in uniform registers; independent barrier probes are issued
    // together so their round trips overlap.
    {
      const int mpar = warp == 1 ? 0 : 1;
      const uint32_t idesc1 = make_idesc_f16(FMT, K::N1);
      const uint32_t idesc2 = make_idesc_f16(FMT, C);
      mbar_wait(w_full, 0);
      auto issue_g2 = [&](int i) {
        const int b1 = i % ND1, b2 = i % ND2;
        const uint32_t ph1 = (i / ND1) & 1, ph2 = ((i / ND2) & 1) ^ 1;
        const bool r1 = mbar_test(&h_full[b1], ph1), r2 = mbar_test(&d2_empty[b2], ph2);
        if (!r1) mbar_wait(&h_full[b1], ph1);
        RB2_TRACE(2, i, 0);
        if (!r2) mbar_wait(&d2_empty[b2], ph2);
        RB2_TRACE(2, i, 1);
        tc_fence_after();
        const uint64_t a_desc = make_kmajor_desc<K::ROWB>(smem_u32(sH + b1 * K::H_BYTES));
        const uint64_t b_desc = make_kmajor_desc<K::ROWB>(smem_u32(sW2));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < K::KB / 16; ++k)
            umma_f16(tmem_base + K::D2_COL + b2 * C, a_desc + 2 * k, b_desc + 2 * k, idesc2, k != 0);
          umma_commit(&d2_full[b2]);
          umma_commit(&h_empty[b1]);
        }
        __syncwarp();
      };
      int i = 0, prev = -1;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
        if ((i & 1) != mpar) continue;
        const int b = i % ND1, ab = i % NA;
        const uint32_t pha = (i / NA) & 1, phd = ((i / ND1) & 1) ^ 1;
        const bool ra = mbar_test(&a_full[ab], pha), rd = mbar_test(&d1_empty[b], phd);
        if (!ra) mbar_wait(&a_full[ab], pha);
        RB2_TRACE(1, i, 0);
        if (!rd) mbar_wait(&d1_empty[b], phd);
        RB2_TRACE(1, i, 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + ab * K::A_SLOT);
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
        }
        __syncwarp();
        RB2_TRACE(1, i, 2);
        if (prev >= 0) issue_g2(prev);        // GEMM2 trails GEMM1 by one own tile (= 2 tiles)
        prev = i;
      }
      if (prev >= 0) issue_g2(prev);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue 1 (warps 2..9): GLU + FiLM -> h
    // two warps per TMEM lane quadrant; warp set `par` owns the tiles with i % 2 == par, so two
    // tiles' GLU epilogues are in flight per quadrant and their latencies overlap
    const int q = warp & 3, par = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    constexpr int CW = C;
    const int cbase = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sBA = reinterpret_cast<const float4*>(sPar);
    const float4* sNB = reinterpret_cast<const float4*>(sPar + C);
    int i = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
      if ((i & 1) != par) continue;
      const int b = i % ND1;
      const uint32_t ph = (i / ND1) & 1;
      const int seq = tile / p.tiles_per_seq, l = (tile - seq * p.tiles_per_seq) * 128 + row;
      int t = l / p.P;
      if (t > p.T - 1) t = p.T - 1;
      const float* film = p.film + ((long long)(seq / p.num_bands) * p.T + t) * p.film_stride;
#pragma unroll
      for (int k = 0; k < (2 * C) / 32; ++k) prefetch_l1(film + k * 32);     // this row's FiLM line(s) -> L1
      mbar_wait(&d1_full[b], ph);
      if (q == 0) RB2_TRACE(3, i, 0);
      mbar_wait(&h_empty[b], ph ^ 1);
      if (q == 0) RB2_TRACE(3, i, 1);
      tc_fence_after();
      uint8_t* hrow = sH + b * K::H_BYTES + row * K::ROWB;
#pragma unroll
      for (int c0 = cbase; c0 < cbase + CW; c0 += 16) {
        uint32_t va[16], vg[16];
        tmem_ld16(lane_addr + b * K::N1 + c0, va);
        tmem_ld16(lane_addr + b * K::N1 + C + c0, vg);
        const float4* fs = reinterpret_cast<const float4*>(film + c0);
        const float4* fh = reinterpret_cast<const float4*>(film + C + c0);
        float4 S[4], H[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          if (p.dbg & 1) { S[i4] = make_float4(1.f, 1.f, 1.f, 1.f); H[i4] = make_float4(0.f, 0.f, 0.f, 0.f); }
          else { S[i4] = __ldg(fs + i4); H[i4] = __ldg(fh + i4); }
        }
        tmem_ld_wait();
        if (q == 0 && c0 == 0) RB2_TRACE(3, i, 3);
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          float hv[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const int i4 = i8 * 2 + h4;
            const float4 A = sBA[(c0 >> 2) + i4], G = sNB[(c0 >> 2) + i4];
            const float av[4] = {A.x, A.y, A.z, A.w}, gv[4] = {G.x, G.y, G.z, G.w};
            const float sv[4] = {S[i4].x, S[i4].y, S[i4].z, S[i4].w}, tv[4] = {H[i4].x, H[i4].y, H[i4].z, H[i4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(va[i4 * 4 + e]) + av[e];
              const float sg = sigmoid_from_half_g(fmaf(__uint_as_float(vg[i4 * 4 + e]), 0.5f, gv[e]));
              hv[h4 * 4 + e] = fmaf(a * sg, sv[e], tv[e]);
            }
          }
          const int chunk = (c0 >> 3) + i8;
          const int phys = K::ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
          *reinterpret_cast<uint4*>(hrow + phys * 16) =
              make_uint4(pack2t<FMT>(hv[0], hv[1]), pack2t<FMT>(hv[2], hv[3]), pack2t<FMT>(hv[4], hv[5]),
                         pack2t<FMT>(hv[6], hv[7]));
        }
        if (q == 0 && c0 == 0) RB2_TRACE(4, i, 2);
      }
      if (q == 0) RB2_TRACE(4, i, 3);
      tc_fence_before();
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the UMMA (async proxy)
      mbar_arrive(&h_full[b]);
      mbar_arrive(&d1_empty[b]);
      if (q == 0) RB2_TRACE(3, i, 2);
    }
  } else if (warp < 18) {
    // ------------------------------------------------------------ epilogue 2 (warps 10..17): residual + store
    const int q = warp & 3, par = (warp - 10) >> 2;
    const int row = q * 32 + lane;
    constexpr int CW = C;
    const int cbase = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float4* sB2 = reinterpret_cast<const float4*>(sPar + 2 * C);
    int i = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
      if ((i & 1) != par) continue;
      const int b = i % ND2, ab = i % NA;
      const uint32_t ph = (i / ND2) & 1;
      const int seq = tile / p.tiles_per_seq, l = (tile - seq * p.tiles_per_seq) * 128 + row;
      mbar_wait(&a_full[ab], (i / NA) & 1);   // visibility of the TMA-written tile to this thread
      mbar_wait(&d2_full[b], ph);
      if (q == 0) RB2_TRACE(4, i, 0);
      tc_fence_after();
      // The output tile is staged IN PLACE over the centre rows of the input tile (each thread
      // overwrites exactly the 16-byte chunks it read) and leaves through one TMA store: no
      // strided per-thread global stores, and rows past the end of the sequence are clipped by TMA.
      uint8_t* xrow = sA + ab * K::A_SLOT + (row + K::HALO) * K::ROWB;
#pragma unroll
      for (int c0 = cbase; c0 < cbase + CW; c0 += 16) {
        uint32_t vd[16];
        tmem_ld16(lane_addr + K::D2_COL + b * C + c0, vd);
        uint4 xa[2];
        uint4* xp[2];
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          const int chunk = (c0 >> 3) + i8;
          const int phys = K::ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
          xp[i8] = reinterpret_cast<uint4*>(xrow + phys * 16);
          xa[i8] = *xp[i8];
        }
        tmem_ld_wait();
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {
          const uint32_t xw[4] = {xa[i8].x, xa[i8].y, xa[i8].z, xa[i8].w};
          const float4 B0 = sB2[(c0 >> 2) + i8 * 2], B1 = sB2[(c0 >> 2) + i8 * 2 + 1];
          const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
          uint32_t ow[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 xs = unpack2t<FMT>(xw[e2]);
            float y0 = (lrelu_inv_fast(xs.x) + bv[e2 * 2]) + __uint_as_float(vd[i8 * 8 + e2 * 2]);
            float y1 = (lrelu_inv_fast(xs.y) + bv[e2 * 2 + 1]) + __uint_as_float(vd[i8 * 8 + e2 * 2 + 1]);
            if (LRELU) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); }
            ow[e2] = pack2t<OFMT>(y0, y1);
          }
          *xp[i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&d2_empty[b]);
      fence_proxy_async_smem();                     // staged tile -> visible to the TMA engine
      named_bar_sync(1 + par, 128);                 // the 4 warps of this parity set
      if (q == 0 && lane == 0 && !(p.dbg & 2)) {
        tma_store_3d(&tmOut, sA + ab * K::A_SLOT + K::HALO * K::ROWB, 0, l - row, seq);
        tma_store_commit();
        tma_store_wait_read();                      // smem has been read: the A slot may be refilled
      }
      if (q == 0 && lane == 0) mbar_arrive(&a_empty[ab]);
      if (q == 0) RB2_TRACE(4, i, 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

long long* g_rb2_trace = nullptr;   // set through b200voc_debug_set_trace (debug only)

static int num_sms() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!n[dev & 15]) cudaDeviceGetAttribute(&n[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  return n[dev & 15];
}

template <int C, int FMT, int OFMT, bool LRELU>
static int launch_resblock2(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                            const float* film, int film_stride, int N, int L, int dilation, int T, int num_bands,
                            void* out16, cudaStream_t stream) {
  using K = Rb2Cfg<C>;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
  B200_TRY(make_tmap_3d(&tmX, a16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, K::KB, K::A_ROWS, K::ROWB));
  B200_TRY(make_tmap_3d(&tmOut, out16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, K::KB, 128, K::ROWB));
  const uint16_t* w1 = reinterpret_cast<const uint16_t*>(w_packed);
  const uint16_t* w2 = w1 + 2ll * C * 3 * C;
  B200_TRY(make_tmap_2d(&tmW1, w1, 3 * C, 2 * C, (uint64_t)3 * C * 2, K::KB, K::N1, K::ROWB));
  B200_TRY(make_tmap_2d(&tmW2, w2, C, C, (uint64_t)C * 2, K::KB, C, K::ROWB));
  Resblock2Params p{};
  p.L = L; p.dilation = dilation; p.T = T; p.P = L / T; p.num_bands = num_bands;
  p.fmt = FMT; p.out_fmt = OFMT; p.store_lrelu = LRELU;
  p.tiles_per_seq = ceil_div(L, 128);
  p.total_tiles = p.tiles_per_seq * N;
  p.b_conv = b_conv; p.b_proj = b_proj; p.film = film; p.film_stride = film_stride;
  p.out = reinterpret_cast<uint16_t*>(out16);
  p.trace = g_rb2_trace;
  {
    const char* e = getenv("B200VOC_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(resblock2_kernel<C, FMT, OFMT, LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   K::SMEM));
    configured[dev & 15] = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  resblock2_kernel<C, FMT, OFMT, LRELU><<<grid, 608, K::SMEM, stream>>>(tmX, tmW1, tmW2, tmOut, p);
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

int resblock2_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(dilation >= 1 && dilation <= 8, "resblock2: dilation %d exceeds the 8-row halo", dilation);
  B200_CHECK_ARG((fmt == 0 || fmt == 1) && (out_fmt == 0 || out_fmt == 1), "resblock2: bad format");
  if (C == 32)
    return dispatch_resblock2<32>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                  out_fmt, store_lrelu, out16, stream);
  if (C == 64)
    return dispatch_resblock2<64>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt,
                                  out_fmt, store_lrelu, out16, stream);
  set_error("resblock2: C=%d unsupported (32/64)", C);
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
