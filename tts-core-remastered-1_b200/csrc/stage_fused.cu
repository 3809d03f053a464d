// K2d: one kernel per NARROW stage (C = 32, 64): ConvTranspose1d (stride 2) + up to three dilated residual
// blocks (+ band_merge + tanh after the last stage) with every intermediate activation kept on chip
// (generator.py:85-98; ResidualBlock body = repair R2).
//
// Why: run block by block, stages 2-3 move 15 GB per step through HBM at 112-224 flop/B and each block pays a full
// trip through a 6-role pipeline per 128 rows (resblock2.cu, profiles/r01_knockouts.txt).  Here a CTA owns a STRIP of
// NT x 128 consecutive time steps of one sequence and carries it through all layers of the stage:
//   * the strip's activations live in ONE shared-memory buffer X[8 guard | R rows | 8 guard] (channels-last 16-bit,
//     K-major swizzled = directly a UMMA A operand) holding leaky_relu(x); a dilated tap is a row-shifted descriptor
//     over X and crosses m-tile borders freely, so only the strip ends lose a halo (sum of dilations, recomputed by the
//     neighbouring strip);
//   * ConvT is the polyphase GEMM of conv_gemm.cu: GEMM row m -> output rows 2m-1, 2m, so a strip starts at an odd
//     output row 2*m0 - 1 and 128 GEMM rows fill 256 strip rows; its epilogue writes leaky_relu(x) straight into X;
//   * GEMM1 (3 taps, N = 2C) -> GLU + FiLM epilogue -> h goes back to TENSOR MEMORY as packed 16-bit pairs over the
//     value columns just read (tcgen05.st) and GEMM2 reads it from there (tcgen05.mma with the A operand in TMEM):
//     no shared-memory h buffer, none of its store / operand-read traffic (the shared-memory crossbar, 128 B/clk, is
//     what bounds M128 x N<=64 MMAs), and the N = C MMA runs at N/2 instead of (A + B bytes) / 128 clocks;
//   * GEMM2's accumulator aliases the gate columns of GEMM1's (dead after the GLU epilogue): 2C TMEM columns per m-tile;
//   * the residual x is recovered from X itself (x = min(a, 10 a), exact inverse of leaky_relu up to the 16-bit rounding
//     x would have had anyway); the epilogue rewrites its own row of X in place -- safe because all GEMM1s of a block
//     are issued before its GEMM2s and the tensor pipe executes in order;
//   * m-tile j of the next block needs only tiles j-1, j, j+1 of this one: the MMA issuer walks
//     [GEMM1 x NT, GEMM2 x NT] per block and waits per tile, so the epilogues of one tile overlap the MMAs of the others;
//   * a CTA carries NCTX strips ("contexts": sequences seq, seq+1 of the same utterance over the same span of time)
//     through the stage together: one strip alone leaves the tensor pipe idle for most of the
//     GEMM1 -> GLU -> GEMM2 -> epilogue 2 -> next GEMM1 chain (measured, profiles/r02_stage_fused_trace.txt: 4.5 k clk
//     per block against 1.9 k clk of MMAs, the epilogue warps busy 30 % of the time); with two strips all 512 TMEM
//     columns hold accumulators (8 m-tiles x 2C columns), the issuer walks [GEMM1 x 8, GEMM2 x 8] and every epilogue
//     warp alternates between its tile of strip 0 and of strip 1.  FiLM coefficients are shared by the contexts;
//   * stage 3: the 4 bands of an utterance run on the same strip of time (two at a time); after the last block the strip
//     (raw x, 16-bit, as band_merge reads it today) is multiplied by the 7 merge taps as ONE N = 16 MMA ([hi | lo] split
//     of the fp32 taps: z_k = m_k . x[l]), and y[l] = sum_k z_k[l + k - 3] is gathered from shared memory into a register
//     that accumulates over the bands; tanh (+ PCM16 / length mask) and the store follow the 4th band.  The 0.9 GB
//     stage-3 output is never written.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 = 4 epilogue groups of 4 warps (one per TMEM lane
// quadrant); C = 32: group g owns m-tile g; C = 64: two groups share an m-tile (32 channels each).
#include <stdlib.h>

// this kernel keeps the inlined mbarrier wait loop (ptx.cuh): its 16 epilogue warps wait often and briefly, and the
// out-of-line slow path costs more there than the smaller code saves (A/B on one box: stage2.a 1.09 vs 1.25 ms)
#ifndef B200VOC_INLINE_WAIT
#define B200VOC_INLINE_WAIT 1
#endif
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct StageFusedParams {
  CUtensorMap tmIn;        // IN_CT: stage input [N, Lin, 2C] raw, box (64 ch, 136 rows); else [N, L, C] leaky_relu(x), box (C, 136)
  CUtensorMap tmCtW;       // packed ConvT weights [2C][4C] (pack_convt_kernel), box (64, 2C)
  CUtensorMap tmW1[3];     // [2C][3C] per block, box (C, 2C)
  CUtensorMap tmW2[3];     // [C][2C] = [W2 | I] per block (resblock2 packing), box (C, C) at column 0
  CUtensorMap tmMW;        // merge taps [nb * 16][32]: rows k = hi(tap k), 8 + k = lo(tap k), box (32, 16)
  int L, Lin, T, P, num_bands, n_seq;
  int strips_per_seq, total_units, V, HL, HR;
  int dil[3];
  const float* b_ct;       // [C]
  const float* b_conv[3];  // [2C]
  const float* b_proj[3];  // [C]
  const float* film;       // [B, T, film_stride]: (1 + scale | shift) per block at film_col
  int film_col[3];
  int film_stride;
  uint16_t* out16;         // OUT 0 / 1: [N, L, C]
  void* wav;               // OUT 2: [B, L] fp32 or int16
  const float* merge_bias;
  const int* valid_samples;
  int pcm16;
  long long* trace;        // clock64 timeline of CTA 0: [16 band-strips][5 roles][32 events] (dev build only)
};

// dev build (-DB200VOC_TRACE): role 0 = MMA issuer, 1..4 = epilogue groups 0..3 (their quadrant-0 warp)
#ifdef B200VOC_TRACE
#define SF_TRACE(role, ev)                                                                  \
  do {                                                                                      \
    if (p.trace && blockIdx.x == 0 && bs < 16 && (threadIdx.x & 31) == 0)                   \
      p.trace[((bs) * 5 + (role)) * 32 + (ev)] = clock64();                                 \
  } while (0)
#else
#define SF_TRACE(role, ev) do { } while (0)
#endif
extern long long* g_rb2_trace;

enum { SF_OUT_LRELU = 0, SF_OUT_RAW = 1, SF_OUT_MERGE = 2 };

#ifndef SF_BIAS_MMA
#define SF_BIAS_MMA 1
#endif
template <int C, bool IN_CT, int NBLK, int OUT, int NCTX>
struct SfCfg {
  static constexpr int ROWB = 2 * C;                  // bytes per X row = swizzle span (64 / 128)
  static constexpr int NT = C == 32 ? 4 : 2;          // m-tiles per strip
  static constexpr int NTT = NCTX * NT;               // m-tiles in flight per CTA
  static constexpr int R = NT * 128;
  static constexpr int G = 8;                         // guard rows on each side of X (>= max dilation)
  static constexpr int NXB = (IN_CT || NCTX == 2) ? NCTX : 2;   // X buffers: one per context; a single TMA-loaded strip is double buffered
  static constexpr int X_BYTES = (R + 2 * G) * ROWB;
  static constexpr int CT_KB = (2 * C) / 64;          // 64-channel k-blocks of the ConvT input
  static constexpr int NCHUNK = R / 256;              // ConvT chunks per strip: 128 GEMM rows -> 256 strip rows
  static constexpr int NSLOT = NCTX * NCHUNK;         // input chunk slots: every (context, chunk) has its own (two issuers must
                                                      // not share a ring: a parity wait is only valid on a slot whose previous phase
                                                      // the waiter has seen complete)
  static constexpr int IN_ROWS = 136;
  static constexpr int IN_KB_BYTES = IN_ROWS * 128;
  static constexpr int IN_CHUNK_BYTES = CT_KB * IN_KB_BYTES;
  static constexpr int CTW_TILE = 2 * C * 128;        // [2C rows][64 k]
  static constexpr int CTW_BYTES = IN_CT ? 2 * CT_KB * CTW_TILE : 0;
  static constexpr int W1_TILE = 2 * C * ROWB;        // one tap
  static constexpr int W2_TILE = C * ROWB;
  static constexpr int BLK_W = 3 * W1_TILE + W2_TILE;
  static constexpr int MW_TILE = 16 * 64;
  static constexpr int NB = 4;                        // bands (merge mode)
  static constexpr int OFF_CTW = 0;
  static constexpr int OFF_W = OFF_CTW + CTW_BYTES;
  static constexpr int OFF_MW = OFF_W + NBLK * BLK_W;
  static constexpr int OFF_X = OFF_MW + (OUT == SF_OUT_MERGE ? NB * MW_TILE : 0);
  static constexpr int OFF_IN = OFF_X + NXB * X_BYTES;
  static constexpr int OFF_Z = OFF_IN + (IN_CT ? NSLOT * IN_CHUNK_BYTES : 0);
  static constexpr int OFF_FILM = OFF_Z + (OUT == SF_OUT_MERGE ? 7 * R * 4 : 0);
  static constexpr int OFF_PAR = OFF_FILM + 16 * 512;                       // per epilogue warp: 2 frames x (S | T) x 32 ch of the current block
  // BIAS_MMA (as in resblock3.cu): biases are added on the tensor core by one more K = 16 MMA per accumulator -- A = an
  // all-ones block (one aliased 128-byte core matrix), B = [b/2 hi, b/2 lo, 0 ...] per GEMM column (both K core
  // matrices alias, so the sum is b) -- and leave the epilogues, which are instruction-issue bound: 3 of the 7.4
  // instructions per element of the GLU epilogue and 1.3 of the 8.7 of epilogue 2 were bias loads / adds.
  static constexpr bool BIAS_MMA = SF_BIAS_MMA != 0;
  static constexpr int OFF_ONES = OFF_PAR;                                  // [128 B ones][ConvT tile 2C x 16 B][per block: GEMM1 2C x 16 B, GEMM2 C x 16 B]
  static constexpr int BT_CT = 128, BT_BLK0 = BT_CT + 2 * C * 16, BT_BLK = 3 * C * 16;
  static constexpr int PAR_BYTES = BIAS_MMA ? BT_BLK0 + NBLK * BT_BLK : (C + NBLK * 3 * C) * 4;
  static constexpr int OFF_BAR = (OFF_PAR + PAR_BYTES + 15) & ~15;
  static constexpr int NBARS = 1 + 2 * NSLOT + NCTX * NCHUNK + 4 + 5 * NTT;
  static constexpr int SMEM = ((OFF_BAR + NBARS * 8 + 16 + 1023) & ~1023) + 1024;
  // TMEM: 2C columns per m-tile: GEMM1 accumulator = [0, 2C) (value | gate); h (packed 16-bit) over the value columns
  // at 16k; GEMM2 accumulator = the gate columns [C, 2C); merge accumulator = [C, C + 16) once epilogue 2 has read them;
  // the ConvT accumulator of chunk c (2C columns) = tile 2c of its context
  static constexpr int N1 = 2 * C;
  // PIPE_CT (one context, room in TMEM): the ConvT accumulators get their own columns and the ConvT MMAs of strip
  // s + 1 are issued behind the GEMM1s of strip s, so they run under its epilogues
  static constexpr bool PIPE_CT = IN_CT && NCTX == 1 && (NTT + NCHUNK) * N1 <= 512;
  static constexpr int CT_COL = NTT * N1;
  static constexpr int TMEM_NEED = NTT * N1 + (PIPE_CT ? NCHUNK * N1 : 0);
  static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int XR_COUNT = C == 32 ? 4 : 8;    // warp arrivals that complete x_ready / h_full of one m-tile
  // x_ready completions per iteration: [ConvT epilogue], one per block, [merge epilogue]
  static constexpr int XPB = NBLK + (IN_CT ? 1 : 0) + (OUT == SF_OUT_MERGE ? 1 : 0);
  static_assert(C == 32 || C == 64, "narrow stages");
  static_assert(NCTX == 1 || NCTX == 2, "contexts");
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(OFF_W % 1024 == 0 && OFF_MW % 1024 == 0 && OFF_X % 1024 == 0 && OFF_IN % 1024 == 0 &&
                X_BYTES % 1024 == 0 && W1_TILE % 1024 == 0 && W2_TILE % 1024 == 0 && IN_KB_BYTES % 1024 == 0,
                "swizzle alignment");
  static_assert(OUT != SF_OUT_MERGE || (C == 32 && IN_CT), "merge follows the last (C = 32) stage");
};

template <int C, bool IN_CT, int NBLK, int OUT, int NCTX, int FMT>
__global__ void __launch_bounds__(576 + 32 * (NCTX - 1), 1)
stage_fused_kernel(const __grid_constant__ StageFusedParams p) {
  using K = SfCfg<C, IN_CT, NBLK, OUT, NCTX>;
  constexpr int NT = K::NT, NTT = K::NTT, R = K::R, G = K::G, ROWB = K::ROWB, N1 = K::N1, NCHUNK = K::NCHUNK, XPB = K::XPB;
  constexpr int NSLOT = K::NSLOT;
  constexpr int NSUB = OUT == SF_OUT_MERGE ? K::NB / NCTX : 1;      // iterations per unit (merge: all bands of the utterance)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sCtW = smem + K::OFF_CTW;
  uint8_t* sW = smem + K::OFF_W;
  uint8_t* sMW = smem + K::OFF_MW;
  uint8_t* sX = smem + K::OFF_X;
  uint8_t* sIn = smem + K::OFF_IN;
  float* sZ = reinterpret_cast<float*>(smem + K::OFF_Z);
  float* sPar = reinterpret_cast<float*>(smem + K::OFF_PAR);      // [b_ct (C)] then per block [ba/2 | bg/2 | b2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* w_full = bars;                      // [1]
  uint64_t* in_full = w_full + 1;               // [NSLOT]          TMA: input chunk landed          -> issuer
  uint64_t* in_empty = in_full + NSLOT;         // [NSLOT]          ConvT MMAs have read the chunk   -> producer
  uint64_t* ct_full = in_empty + NSLOT;         // [NCTX * NCHUNK]  ConvT accumulator complete       -> epilogue
  uint64_t* xin_full = ct_full + NCTX * NCHUNK; // [2]              TMA: lrelu(x) strip landed in X[b] -> issuer, epilogue
  uint64_t* x_free = xin_full + 2;              // [2]              last block's epilogue is done with X[b] -> producer
  uint64_t* x_ready = x_free + 2;               // [NTT]  X rows of the m-tile written, its TMEM columns drained -> issuer
  uint64_t* d1_full = x_ready + NTT;            // [NTT]
  uint64_t* h_full = d1_full + NTT;             // [NTT]  h in TMEM, D1 drained                      -> issuer
  uint64_t* d2_full = h_full + NTT;             // [NTT]
  uint64_t* z_full = d2_full + NTT;             // [NTT]  merge accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_full + NTT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---------------------------------------------------------------- one-time setup
  for (int i = threadIdx.x; i < K::NXB * K::X_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sX)[i] = make_uint4(0u, 0u, 0u, 0u);              // guard rows stay zero for good
  if (!K::BIAS_MMA) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      sPar[i] = IN_CT ? p.b_ct[i] : 0.f;
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        float* q = sPar + C + b * 3 * C;
        q[i] = 0.5f * p.b_conv[b][i];            // W1 is packed pre-scaled by 1/2 (pack_resblock_kernel)
        q[C + i] = 0.5f * p.b_conv[b][C + i];
        q[2 * C + i] = p.b_proj[b][i];
      }
    }
  } else {
    // bias tiles, SWIZZLE_NONE K-major core matrices: GEMM column n = 16 bytes at n * 16: (b/2)_hi, (b/2)_lo, 0 x 6
    uint8_t* sB = smem + K::OFF_ONES;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) reinterpret_cast<uint16_t*>(sB)[i] = FMT == 0 ? 0x3C00 : 0x3F80;
    auto put = [&](uint8_t* tile, int n, float bhalf) {
      const float lo = bhalf - unpack2t<FMT>(pack2t<FMT>(bhalf, 0.f)).x;
      uint32_t* row = reinterpret_cast<uint32_t*>(tile) + n * 4;
      row[0] = pack2t<FMT>(bhalf, lo);
      row[1] = 0u; row[2] = 0u; row[3] = 0u;
    };
    for (int n = threadIdx.x; n < 2 * C; n += blockDim.x) {
      put(sB + K::BT_CT, n, IN_CT ? 0.5f * p.b_ct[n % C] : 0.f);              // ConvT column n = phase * C + channel
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        put(sB + K::BT_BLK0 + b * K::BT_BLK, n, 0.25f * p.b_conv[b][n]);      // value | gate; W1 is packed pre-scaled by 1/2
        if (n < C) put(sB + K::BT_BLK0 + b * K::BT_BLK + 2 * C * 16, n, 0.5f * p.b_proj[b][n]);
      }
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmIn);
    mbar_init(w_full, 1);
    for (int c = 0; c < NSLOT; ++c) { mbar_init(&in_full[c], 1); mbar_init(&in_empty[c], 1); }
    for (int c = 0; c < NCTX * NCHUNK; ++c) mbar_init(&ct_full[c], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&xin_full[b], 1); mbar_init(&x_free[b], 16); }
    for (int t = 0; t < NTT; ++t) {
      mbar_init(&x_ready[t], K::XR_COUNT);
      mbar_init(&d1_full[t], 1);
      mbar_init(&h_full[t], K::XR_COUNT);
      mbar_init(&d2_full[t], 1);
      mbar_init(&z_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, K::TMEM_COLS);
  fence_proxy_async_smem();                      // the zeroed X is visible to the async proxy (UMMA reads the guards)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int grid = gridDim.x;

  // unit -> (sequence group, first strip row).  A unit is one span of time of NCTX consecutive sequences (NSUB = 1) or of
  // all num_bands sequences of one utterance (merge: NSUB iterations of NCTX bands); iteration `it` of the CTA handles
  // sequences seq0 .. seq0 + NCTX - 1
  auto unit_geom = [&](int u, int& sg, int& s0) {
    sg = u / p.strips_per_seq;
    s0 = (u - sg * p.strips_per_seq) * p.V - p.HL;
  };
  // X buffer of context cx in iteration it; TMA-loaded strips: barrier index and phase of xin_full / x_free
  auto xbuf = [&](int cx, int it) { return (IN_CT || NCTX == 2) ? cx : (it & 1); };
  auto xin_idx = [&](int cx, int it) { return NCTX == 2 ? cx : (it & 1); };
  auto xin_par = [&](int it) { return (uint32_t)(NCTX == 2 ? (it & 1) : ((it >> 1) & 1)); };

  if (warp == 0) {
    // ============================================================== TMA producer
    if (lane == 0) {
      uint32_t wbytes = K::CTW_BYTES + NBLK * K::BLK_W + (OUT == SF_OUT_MERGE ? K::NB * K::MW_TILE : 0);
      mbar_expect_tx(w_full, wbytes);
      if (IN_CT)
        for (int t = 0; t < 2 * K::CT_KB; ++t) tma_load_2d(sCtW + t * K::CTW_TILE, &p.tmCtW, w_full, t * 64, 0);
      for (int b = 0; b < NBLK; ++b) {
        for (int tap = 0; tap < 3; ++tap) tma_load_2d(sW + b * K::BLK_W + tap * K::W1_TILE, &p.tmW1[b], w_full, tap * C, 0);
        tma_load_2d(sW + b * K::BLK_W + 3 * K::W1_TILE, &p.tmW2[b], w_full, 0, 0);
      }
      if (OUT == SF_OUT_MERGE)
        for (int b = 0; b < K::NB; ++b) tma_load_2d(sMW + b * K::MW_TILE, &p.tmMW, w_full, 0, b * 16);
      int it = 0;
      for (int u = blockIdx.x; u < p.total_units; u += grid) {
        int sg, s0;
        unit_geom(u, sg, s0);
        for (int sub = 0; sub < NSUB; ++sub, ++it) {
          const int seq0 = (sg * NSUB + sub) * NCTX;
          if (IN_CT) {
            const int m0 = (s0 + 1) / 2;          // s0 is odd: strip row 0 = output row 2*m0 - 1
            for (int cx = 0; cx < NCTX; ++cx)
              for (int c = 0; c < NCHUNK; ++c) {
                const int sl = cx * NCHUNK + c;
                mbar_wait(&in_empty[sl], (it & 1) ^ 1);
                mbar_expect_tx(&in_full[sl], K::IN_CHUNK_BYTES);
                for (int kb = 0; kb < K::CT_KB; ++kb)
                  tma_load_3d(sIn + sl * K::IN_CHUNK_BYTES + kb * K::IN_KB_BYTES, &p.tmIn, &in_full[sl], kb * 64,
                              m0 - 1 + 128 * c, seq0 + cx);
              }
          } else {
            for (int cx = 0; cx < NCTX; ++cx) {
              const int bi = xin_idx(cx, it);
              mbar_wait(&x_free[bi], xin_par(it) ^ 1);
              mbar_expect_tx(&xin_full[bi], K::X_BYTES);
              for (int j = 0; j < (R + 2 * G) / K::IN_ROWS; ++j)
                tma_load_3d(sX + xbuf(cx, it) * K::X_BYTES + j * K::IN_ROWS * ROWB, &p.tmIn, &xin_full[bi], 0,
                            s0 - G + j * K::IN_ROWS, seq0 + cx);
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 18) {
    // ============================================================== MMA issuers: warp 1 runs context 0, warp 18 (NCTX = 2)
    // context 1 -- two independent in-order programs.  (One warp issuing for both contexts was the bottleneck: its
    // ~300 clk of scalar work per op -- barrier probe, fence, descriptors through the uniform datapath, commit -- is not
    // hidden behind anything, 8 GEMM1 + 8 GEMM2 ops cost 8 k clk per block against 3.8 k clk of MMAs.)  Warp-uniform
    // control flow, one elected lane issues.
    const int cx = warp == 1 ? 0 : 1;
    const uint32_t idesc1 = make_idesc_f16(FMT, N1);     // GEMM1 and ConvT: N = 2C
    const uint32_t idesc2 = make_idesc_f16(FMT, C);
    const uint32_t idesc_z = make_idesc_f16(FMT, 16);
    // bias MMA: D += ones[128 x 16] * tile[N x 16]^T, SWIZZLE_NONE descriptors (ones: LBO = SBO = 0; tile: SBO = 128, LBO = 0)
    const uint64_t ones_desc = (uint64_t)((smem_u32(smem + K::OFF_ONES) & 0x3FFFFu) >> 4) | (1ull << 46);
    auto bias_desc = [&](int off) {
      return (uint64_t)((smem_u32(smem + K::OFF_ONES + off) & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    };
    mbar_wait(w_full, 0);
    int it = 0;
    for (int u = blockIdx.x; u < p.total_units; u += grid) {
      for (int sub = 0; sub < NSUB; ++sub, ++it) {
        const int bs = it;
        const uint32_t xbase = smem_u32(sX + xbuf(cx, it) * K::X_BYTES);
        if (it > 0) {
          // the previous iteration's last epilogues have left this context's tensor memory and X
          for (int t = 0; t < NT; ++t) mbar_wait(&x_ready[cx * NT + t], (it * XPB - 1) & 1);
        }
        if (cx == 0) SF_TRACE(0, 0);
        auto issue_ct = [&](int it_ct) {
          for (int c = 0; c < NCHUNK; ++c) {
            const int sl = cx * NCHUNK + c;
            mbar_wait(&in_full[sl], it_ct & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sIn + sl * K::IN_CHUNK_BYTES);
            const uint32_t d_ct = tmem_base + (K::PIPE_CT ? K::CT_COL + c * N1 : (cx * NT + 2 * c) * N1);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 2; ++tap)      // tap 0: x[m] (tile row i + 1), tap 1: x[m - 1] (tile row i)
#pragma unroll
                for (int kb = 0; kb < K::CT_KB; ++kb) {
                  const uint64_t a_desc = make_kmajor_desc<128>(a0 + kb * K::IN_KB_BYTES + (1 - tap) * 128);
                  const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sCtW + (tap * K::CT_KB + kb) * K::CTW_TILE));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16(d_ct, a_desc + 2 * k, b_desc + 2 * k, idesc1, (tap | kb | k) != 0);
                }
              if (K::BIAS_MMA) umma_f16(d_ct, ones_desc, bias_desc(K::BT_CT), idesc1, 1);
              umma_commit(&ct_full[cx * NCHUNK + c]);
              umma_commit(&in_empty[sl]);
            }
            __syncwarp();
          }
        };
        if (IN_CT && (!K::PIPE_CT || it == 0)) {
          issue_ct(it);
          if (cx == 0) SF_TRACE(0, 2);
        }
#pragma unroll 1
        for (int blk = 0; blk < NBLK; ++blk) {
          // [GEMM1 x NT, GEMM2 x NT] in this order: the epilogue behind GEMM2(t) rewrites tile t's rows of X in place,
          // which GEMM1(t +- 1) read -- the tensor pipe runs one thread's MMAs in issue order.  (A polling scheduler that
          // issued whichever of the two queues was ready was SLOWER: every mbarrier probe costs ~150 clk.)
          const int d = p.dil[blk];
          const uint32_t w1 = smem_u32(sW + blk * K::BLK_W), w2 = w1 + 3 * K::W1_TILE;
          for (int t = 0; t < NT; ++t) {
            const int T = cx * NT + t;
            // inputs of this block for tiles t-1, t, t+1 of the strip
            if (!IN_CT && blk == 0) {
              if (t == 0) mbar_wait(&xin_full[xin_idx(cx, it)], xin_par(it));
            } else {
              const uint32_t par = (it * XPB + blk - (IN_CT ? 0 : 1)) & 1;
              if (t == 0) { mbar_wait(&x_ready[T], par); if (NT > 1) mbar_wait(&x_ready[T + 1], par); }
              else if (t + 1 < NT) mbar_wait(&x_ready[T + 1], par);
            }
            tc_fence_after();
            if (cx == 0) SF_TRACE(0, 3 + blk * 8 + t);
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 3; ++tap) {
                const uint64_t a_desc = make_kmajor_desc<ROWB>(xbase + (G + 128 * t + (tap - 1) * d) * ROWB);
                const uint64_t b_desc = make_kmajor_desc<ROWB>(w1 + tap * K::W1_TILE);
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                  umma_f16(tmem_base + T * N1, a_desc + 2 * k, b_desc + 2 * k, idesc1, (tap | k) != 0);
              }
              if (K::BIAS_MMA) umma_f16(tmem_base + T * N1, ones_desc, bias_desc(K::BT_BLK0 + blk * K::BT_BLK), idesc1, 1);
              umma_commit(&d1_full[T]);
            }
            __syncwarp();
          }
          if (K::PIPE_CT && blk == NBLK - 1) {
            // next strip's ConvT behind this strip's last GEMM1s: its accumulator columns are free (the ConvT epilogue of
            // this strip arrived on x_ready long ago) and its input chunk was prefetched
            const bool more = sub + 1 < NSUB || u + grid < p.total_units;
            if (more) issue_ct(it + 1);
          }
          for (int t = 0; t < NT; ++t) {
            const int T = cx * NT + t;
            mbar_wait(&h_full[T], (it * NBLK + blk) & 1);
            tc_fence_after();
            if (cx == 0) SF_TRACE(0, 3 + blk * 8 + 4 + t);
            if (elect_one()) {
              const uint64_t b_desc = make_kmajor_desc<ROWB>(w2);
#pragma unroll
              for (int k = 0; k < C / 16; ++k)      // A = h in TMEM: k-step k sits on the value columns 16k .. 16k+7
                umma_f16_ts(tmem_base + T * N1 + C, tmem_base + T * N1 + 16 * k, b_desc + 2 * k, idesc2, k != 0);
              if (K::BIAS_MMA)
                umma_f16(tmem_base + T * N1 + C, ones_desc, bias_desc(K::BT_BLK0 + blk * K::BT_BLK + 2 * C * 16), idesc2, 1);
              umma_commit(&d2_full[T]);
            }
            __syncwarp();
          }
        }
        if (OUT == SF_OUT_MERGE) {
          for (int t = 0; t < NT; ++t) {
            const int T = cx * NT + t;
            mbar_wait(&x_ready[T], (it * XPB + NBLK) & 1);
            tc_fence_after();
            if (cx == 0) SF_TRACE(0, 27 + t);
            if (elect_one()) {
              const uint64_t a_desc = make_kmajor_desc<ROWB>(xbase + (G + 128 * t) * ROWB);
              const uint64_t b_desc = make_kmajor_desc<ROWB>(smem_u32(sMW + (sub * NCTX + cx) * K::MW_TILE));
#pragma unroll
              for (int k = 0; k < C / 16; ++k)
                umma_f16(tmem_base + T * N1 + C, a_desc + 2 * k, b_desc + 2 * k, idesc_z, k != 0);
              umma_commit(&z_full[T]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ============================================================== epilogue groups (warps 2..17)
    const int eg = (warp - 2) >> 2, q = warp & 3, ew = warp - 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int mt = C == 32 ? eg : (eg >> 1);               // m-tile (within a strip) of the block epilogues
    const int c_lo = C == 32 ? 0 : 32 * (eg & 1);          // this thread's 32 channels
    // ConvT epilogue geometry: chunk, output phase r, accumulator columns, m-tile the written rows fall into
    const int ct_c = C == 32 ? (eg >> 1) : 0;
    const int ct_r = C == 32 ? (eg & 1) : (eg >> 1);
    const int ct_col = C == 32 ? 32 * ct_r : 64 * ct_r + c_lo;
    const int ct_mt = 2 * ct_c + (q >> 1);
    const int ct_sr = 256 * ct_c + 2 * (32 * q + lane) + ct_r;   // strip row this thread writes
    const int sr = 128 * mt + 32 * q + lane;                     // strip row of the block epilogues
    float* scratch = reinterpret_cast<float*>(smem + K::OFF_FILM) + ew * 128;
    auto sw_chunk = [](int row, int j) { return ROWB == 128 ? (j ^ (row & 7)) : (j ^ ((row >> 1) & 3)); };
    float y_acc = 0.f;
    int it = 0;
    for (int u = blockIdx.x; u < p.total_units; u += grid) {
      int sg, s0;
      unit_geom(u, sg, s0);
      for (int sub = 0; sub < NSUB; ++sub, ++it) {
        const int bs = it;
        const int seq0 = (sg * NSUB + sub) * NCTX;
        const int bidx = seq0 / p.num_bands;                    // utterance (FiLM row): the same for all contexts
        const int l = s0 + sr;
        const bool in_seq = l >= 0 && l < p.L;
        const bool warp_in_seq = __all_sync(0xffffffffu, in_seq);
        // FiLM coefficients of this warp's 32 rows (at most two frames: P >= 32) for every block of the iteration,
        // fetched now so that the L2 latency hides behind the ConvT phase
        const int lw = s0 + 128 * mt + 32 * q;
        const int t_first = lw < 0 ? 0 : min(lw / p.P, p.T - 1);
        const int t_last = lw + 31 < 0 ? 0 : min((lw + 31) / p.P, p.T - 1);
        const int t_mine = l < 0 ? 0 : min(l / p.P, p.T - 1);
        const float* my_film = scratch + (t_mine - t_first) * 64;
        float4 film_st[NBLK];                                   // in flight during the ConvT phase, parked in shared memory after it
        {
          const int f = lane >> 4, which = (lane >> 3) & 1, j = lane & 7;
          const float* src = p.film + ((long long)bidx * p.T + (f ? t_last : t_first)) * p.film_stride + which * C + c_lo;
#pragma unroll
          for (int b = 0; b < NBLK; ++b) film_st[b] = __ldg(reinterpret_cast<const float4*>(src + p.film_col[b]) + j);
        }
        if (IN_CT) {
          // ---------------------------------------------------------- ConvT epilogue: + bias, leaky_relu -> X
#pragma unroll 1
          for (int cx = 0; cx < NCTX; ++cx) {
            uint8_t* X = sX + xbuf(cx, it) * K::X_BYTES;
            if (it > 0) mbar_wait(&x_ready[cx * NT + ct_mt], (it * XPB - 1) & 1);   // the tile's previous owner is done with its rows
            mbar_wait(&ct_full[cx * NCHUNK + ct_c], it & 1);
            tc_fence_after();
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 0);
            uint32_t v[32];
            tmem_ld32(lane_addr + (K::PIPE_CT ? K::CT_COL + ct_c * N1 : (cx * NT + 2 * ct_c) * N1) + ct_col, v);
            tmem_ld_wait();
            const int cl = s0 + ct_sr;
            const float keep = (cl >= 0 && cl < p.L) ? 1.f : 0.f;           // rows outside the sequence are the next conv's zero padding
            const int crow = G + ct_sr;
            uint8_t* cxrow = X + crow * ROWB;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;       // BIAS_MMA: already in the accumulator
              if (!K::BIAS_MMA) {
                B0 = *reinterpret_cast<const float4*>(sPar + c_lo + 8 * j);
                B1 = *reinterpret_cast<const float4*>(sPar + c_lo + 8 * j + 4);
              }
              const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
              uint32_t w[4];
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                // (x + 0.0f is not a no-op for the compiler: -0 + 0 = +0, so the adds are selected away explicitly)
                const float x0 = (K::BIAS_MMA ? __uint_as_float(v[8 * j + 2 * e2]) : __uint_as_float(v[8 * j + 2 * e2]) + bv[2 * e2]) * keep;
                const float x1 = (K::BIAS_MMA ? __uint_as_float(v[8 * j + 2 * e2 + 1]) : __uint_as_float(v[8 * j + 2 * e2 + 1]) + bv[2 * e2 + 1]) * keep;
                w[e2] = pack2t<FMT>(lrelu_fast(x0), lrelu_fast(x1));
              }
              *reinterpret_cast<uint4*>(cxrow + (sw_chunk(crow, (c_lo >> 3) + j) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_ready[cx * NT + ct_mt]);
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 1);
          }
        } else {
          for (int cx = 0; cx < NCTX; ++cx) mbar_wait(&xin_full[xin_idx(cx, it)], xin_par(it));   // acquire the TMA-written strips (residual reads)
        }
        const int row = G + sr;
#pragma unroll 1
        for (int blk = 0; blk < NBLK; ++blk) {
          const float* par = sPar + C + blk * 3 * C;
          const float* film_b = my_film;
          {   // this block's FiLM coefficients -> the warp's scratch (the previous block's readers are through: __syncwarp below)
            float4 f4 = film_st[0];
#pragma unroll
            for (int b = 1; b < NBLK; ++b) if (blk == b) f4 = film_st[b];
            reinterpret_cast<float4*>(scratch)[(lane >> 4) * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)] = f4;
            __syncwarp();
          }
          const bool last = blk == NBLK - 1;
          // ---------------------------------------------------------- GLU + FiLM epilogue: D1 -> h (TMEM)
#pragma unroll 1
          for (int cx = 0; cx < NCTX; ++cx) {
            const int T = cx * NT + mt;
            mbar_wait(&d1_full[T], (it * NBLK + blk) & 1);
            tc_fence_after();
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 2 + blk * 4);
#pragma unroll
            for (int cc = 0; cc < 32; cc += 16) {
              uint32_t va[16], vg[16];
              tmem_ld16(lane_addr + T * N1 + c_lo + cc, va);
              tmem_ld16(lane_addr + T * N1 + C + c_lo + cc, vg);
              tmem_ld_wait();
              uint32_t hw[8];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                float4 A = make_float4(0.f, 0.f, 0.f, 0.f), Gt = A;        // BIAS_MMA: already in the accumulator
                if (!K::BIAS_MMA) {
                  A = *reinterpret_cast<const float4*>(par + c_lo + cc + 4 * i4);
                  Gt = *reinterpret_cast<const float4*>(par + C + c_lo + cc + 4 * i4);
                }
                const float4 S = *reinterpret_cast<const float4*>(film_b + cc + 4 * i4);
                const float4 Tt = *reinterpret_cast<const float4*>(film_b + 32 + cc + 4 * i4);
                const float av[4] = {A.x, A.y, A.z, A.w}, gv[4] = {Gt.x, Gt.y, Gt.z, Gt.w};
                const float sv[4] = {S.x, S.y, S.z, S.w}, tv[4] = {Tt.x, Tt.y, Tt.z, Tt.w};
                float hv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = K::BIAS_MMA ? __uint_as_float(va[4 * i4 + e]) : __uint_as_float(va[4 * i4 + e]) + av[e];   // (conv_a + b_a) / 2
                  const float th = tanh_approx(K::BIAS_MMA ? __uint_as_float(vg[4 * i4 + e]) : __uint_as_float(vg[4 * i4 + e]) + gv[e]);   // tanh(g / 2)
                  hv[e] = fmaf(fmaf(a, th, a), sv[e], tv[e]);                              // a sigmoid(g) (1 + scale) + shift
                }
                hw[2 * i4] = pack2t<FMT>(hv[0], hv[1]);
                hw[2 * i4 + 1] = pack2t<FMT>(hv[2], hv[3]);
              }
              tmem_st8(lane_addr + T * N1 + c_lo + cc, hw);     // over the value columns just read
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_full[T]);
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 3 + blk * 4);
          }
          // ---------------------------------------------------------- epilogue 2: x + W2 h + b2
#pragma unroll 1
          for (int cx = 0; cx < NCTX; ++cx) {
            const int T = cx * NT + mt;
            uint8_t* xrow = sX + xbuf(cx, it) * K::X_BYTES + row * ROWB;
            mbar_wait(&d2_full[T], (it * NBLK + blk) & 1);
            tc_fence_after();
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 4 + blk * 4);
            uint32_t vd[32];
            tmem_ld32(lane_addr + T * N1 + C + c_lo, vd);
            uint4 xa[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xa[j] = *reinterpret_cast<const uint4*>(xrow + (sw_chunk(row, (c_lo >> 3) + j) << 4));
            tmem_ld_wait();
            uint4 ow[4];
            const bool act = !last || OUT == SF_OUT_LRELU;      // store leaky_relu(y) (the next GEMM1's operand) or raw y
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t xw[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w};
              float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;       // BIAS_MMA: already in the accumulator
              if (!K::BIAS_MMA) {
                B0 = *reinterpret_cast<const float4*>(par + 2 * C + c_lo + 8 * j);
                B1 = *reinterpret_cast<const float4*>(par + 2 * C + c_lo + 8 * j + 4);
              }
              const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
              uint32_t o[4];
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                // (packed 16-bit leaky_relu / inverse here saves ~1.5 instructions per element but was measured to cost
                // 3.5 dB of SNR (70.1 -> 66.6) for no change in the epilogue's latency: fp32 it is)
                const float2 xs = unpack2t<FMT>(xw[e2]);
                float y0 = (K::BIAS_MMA ? lrelu_inv_fast(xs.x) : lrelu_inv_fast(xs.x) + bv[2 * e2]) + __uint_as_float(vd[8 * j + 2 * e2]);
                float y1 = (K::BIAS_MMA ? lrelu_inv_fast(xs.y) : lrelu_inv_fast(xs.y) + bv[2 * e2 + 1]) + __uint_as_float(vd[8 * j + 2 * e2 + 1]);
                if (act) { y0 = lrelu_fast(y0); y1 = lrelu_fast(y1); }
                o[e2] = pack2t<FMT>(y0, y1);
              }
              ow[j] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (!warp_in_seq && !in_seq) {                      // rows outside the sequence: the next conv's zero padding
#pragma unroll
              for (int j = 0; j < 4; ++j) ow[j] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (!last || OUT == SF_OUT_MERGE) {
#pragma unroll
              for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(xrow + (sw_chunk(row, (c_lo >> 3) + j) << 4)) = ow[j];
              fence_proxy_async_smem();
            } else if (in_seq && sr >= p.HL && sr < R - p.HR) {
              uint4* dst = reinterpret_cast<uint4*>(p.out16 + ((long long)(seq0 + cx) * p.L + l) * C + c_lo);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = ow[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&x_ready[T]);
              if (!IN_CT && last) mbar_arrive(&x_free[xin_idx(cx, it)]);
            }
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 5 + blk * 4);
          }
        }
        if (OUT == SF_OUT_MERGE) {
          // ---------------------------------------------------------- merge: z_k[l] = m_k . x[l]  ->  y[l] += sum_k z_k[l+k-3]
#pragma unroll 1
          for (int cx = 0; cx < NCTX; ++cx) {
            const int T = cx * NT + mt;
            mbar_wait(&z_full[T], it & 1);
            tc_fence_after();
            if (q == 0 && cx == 0) SF_TRACE(1 + eg, 14);
            uint32_t vz[16];
            tmem_ld16(lane_addr + T * N1 + C, vz);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_ready[T]);              // the tile's TMEM columns are free for the next iteration
            named_bar_sync(1, 512);                               // the previous band's gather is complete
#pragma unroll
            for (int k = 0; k < 7; ++k) sZ[k * R + sr] = __uint_as_float(vz[k]) + __uint_as_float(vz[8 + k]);
            named_bar_sync(1, 512);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              const int rr = sr + k - 3;
              if (rr >= 0 && rr < R) y_acc += sZ[k * R + rr];
            }
          }
          if (sub == NSUB - 1) {
            if (in_seq && sr >= p.HL && sr < R - p.HR) {
              float y = tanhf(y_acc + __ldg(p.merge_bias));
              if (p.valid_samples != nullptr && l >= __ldg(p.valid_samples + sg)) y = 0.f;
              if (p.pcm16) {
                const float cl = fminf(fmaxf(y, -1.f), 1.f) * 32767.f;
                reinterpret_cast<int16_t*>(p.wav)[(long long)sg * p.L + l] = (int16_t)__float2int_rn(cl);
              } else {
                reinterpret_cast<float*>(p.wav)[(long long)sg * p.L + l] = y;
              }
            }
            y_acc = 0.f;
          }
          if (q == 0) SF_TRACE(1 + eg, 15);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

static int sf_num_sms() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!n[dev & 15]) cudaDeviceGetAttribute(&n[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  return n[dev & 15];
}

template <int C, bool IN_CT, int NBLK, int OUT, int NCTX, int FMT>
static int launch_stage_fused_t(const StageFusedArgs& a, cudaStream_t stream) {
  using K = SfCfg<C, IN_CT, NBLK, OUT, NCTX>;
  StageFusedParams p{};
  const int L = IN_CT ? 2 * a.Lin : a.Lin;
  p.L = L; p.Lin = a.Lin; p.T = a.T; p.P = L / a.T; p.num_bands = a.num_bands; p.n_seq = a.N;
  int halo = OUT == SF_OUT_MERGE ? 3 : 0;
  for (int b = 0; b < NBLK; ++b) {
    B200_CHECK_ARG(a.dil[b] >= 1 && a.dil[b] <= K::G, "stage_fused: dilation %d exceeds the %d-row guard", a.dil[b], K::G);
    halo += a.dil[b];
    p.dil[b] = a.dil[b];
    p.b_conv[b] = a.b_conv[b]; p.b_proj[b] = a.b_proj[b]; p.film_col[b] = a.film_col[b];
  }
  if (IN_CT) halo |= 1;                          // strips start at an odd output row (ConvT phase alignment)
  p.HL = halo; p.HR = halo; p.V = K::R - 2 * halo;
  p.strips_per_seq = ceil_div(L, p.V);
  B200_CHECK_ARG(a.num_bands % NCTX == 0, "stage_fused: %d contexts need num_bands %% %d == 0", NCTX, NCTX);
  const int groups = OUT == SF_OUT_MERGE ? a.N / a.num_bands : a.N / NCTX;
  p.total_units = p.strips_per_seq * groups;
  B200_CHECK_ARG(p.P >= 32 && L % a.T == 0, "stage_fused: L=%d T=%d", L, a.T);
  B200_CHECK_ARG(OUT != SF_OUT_MERGE || a.num_bands == K::NB, "stage_fused: merge needs %d bands", K::NB);
  p.b_ct = a.ct_b; p.film = a.film; p.film_stride = a.film_stride;
  p.out16 = reinterpret_cast<uint16_t*>(a.out16);
  p.wav = a.wav; p.merge_bias = a.merge_b; p.valid_samples = a.valid_samples; p.pcm16 = a.pcm16;
  {   // dev build: B200VOC_SF_TRACE = 32 | 64a | 64b selects which launch of the step records its timeline
    const char* e = getenv("B200VOC_SF_TRACE");
    const char* me = C == 32 ? "32" : (IN_CT ? "64a" : "64b");
    p.trace = (e ? strcmp(e, me) == 0 : C == 32) ? g_rb2_trace : nullptr;
  }
  if (IN_CT) {
    B200_TRY(make_tmap_3d(&p.tmIn, a.x_in, 2 * C, a.Lin, a.N, (uint64_t)2 * C * 2, (uint64_t)a.Lin * 2 * C * 2, 64,
                          K::IN_ROWS, 128));
    B200_TRY(make_tmap_2d(&p.tmCtW, a.ct_w, 4 * C, 2 * C, (uint64_t)4 * C * 2, 64, 2 * C, 128));
  } else {
    B200_TRY(make_tmap_3d(&p.tmIn, a.x_in, C, L, a.N, (uint64_t)C * 2, (uint64_t)L * C * 2, C, K::IN_ROWS, K::ROWB));
    p.tmCtW = p.tmIn;
  }
  for (int b = 0; b < 3; ++b) {
    const int bb = b < NBLK ? b : 0;
    const uint16_t* w1 = reinterpret_cast<const uint16_t*>(a.blk_w[bb]);
    const uint16_t* w2 = w1 + 2ll * C * 3 * C;
    B200_TRY(make_tmap_2d(&p.tmW1[b], w1, 3 * C, 2 * C, (uint64_t)3 * C * 2, C, 2 * C, K::ROWB));
    B200_TRY(make_tmap_2d(&p.tmW2[b], w2, 2 * C, C, (uint64_t)2 * C * 2, C, C, K::ROWB));
  }
  if (OUT == SF_OUT_MERGE) B200_TRY(make_tmap_2d(&p.tmMW, a.merge_w16, 32, 16 * K::NB, 64, 32, 16, 64));
  else p.tmMW = p.tmIn;
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  auto kernel = stage_fused_kernel<C, IN_CT, NBLK, OUT, NCTX, FMT>;
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    configured[dev & 15] = true;
  }
  const int sms = sf_num_sms();
  const int grid = p.total_units < sms ? p.total_units : sms;
  kernel<<<grid, 576 + 32 * (NCTX - 1), K::SMEM, stream>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int stage_fused_launch(const StageFusedArgs& a, cudaStream_t stream) {
  B200_CHECK_ARG(a.fmt == 0 || a.fmt == 1, "stage_fused: bad format");
  B200_CHECK_ARG(a.N > 0 && a.Lin > 0 && a.T > 0 && a.N % a.num_bands == 0, "stage_fused: bad shape");
  // B200VOC_SF_NCTX=1: one strip per CTA instead of two for the C = 32 stage (A/B runs)
  static const bool one_ctx = [] { const char* e = getenv("B200VOC_SF_NCTX"); return e && e[0] == '1'; }();
#define SF(CC, CT, NB_, O, NC)                                                                 \
  if (a.C == CC && (a.in_ct != 0) == CT && a.nblk == NB_ && a.out_mode == O)                   \
    return a.fmt == 0 ? launch_stage_fused_t<CC, CT, NB_, O, NC, 0>(a, stream) : launch_stage_fused_t<CC, CT, NB_, O, NC, 1>(a, stream)
  if (!one_ctx && a.num_bands % 2 == 0) {
    SF(32, true, 3, SF_OUT_MERGE, 2);
    SF(32, true, 3, SF_OUT_RAW, 2);
  }
  SF(32, true, 3, SF_OUT_MERGE, 1);
  SF(32, true, 3, SF_OUT_RAW, 1);
  SF(64, true, 1, SF_OUT_LRELU, 1);
  if (!one_ctx && a.num_bands % 2 == 0) SF(64, false, 2, SF_OUT_RAW, 2);
  SF(64, false, 2, SF_OUT_RAW, 1);
#undef SF
  set_error("stage_fused: unsupported configuration (C=%d in_ct=%d nblk=%d out=%d)", a.C, a.in_ct, a.nblk, a.out_mode);
  return B200VOC_ERR_UNSUPPORTED;
}

// merge taps [1][nb * 32][7] fp32 -> [nb][16][32] 16-bit: row k = hi(w[band*32 + c][k]), row 8 + k = w - hi
__global__ void pack_merge_kernel(const float* __restrict__ w, int nb, int fmt, uint16_t* __restrict__ out) {
  const int total = nb * 16 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 31, r = (i >> 5) & 15, band = i >> 9;
    const int k = r & 7;
    float v = 0.f;
    if (k < 7) {
      const float full = w[(band * 32 + c) * 7 + k];
      const float hi = fmt == 0 ? __half2float(__float2half_rn(full)) : __bfloat162float(__float2bfloat16_rn(full));
      v = r < 8 ? hi : full - hi;
    }
    out[i] = fmt == 0 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}
int pack_merge_launch(const float* w, int nb, int fmt, void* out, cudaStream_t st) {
  pack_merge_kernel<<<8, 256, 0, st>>>(w, nb, fmt, reinterpret_cast<uint16_t*>(out));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
