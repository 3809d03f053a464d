// K2: one fused kernel per residual block (generator.py:40-41,89-90; body = repair R2):
//
//   h   = GLU( Conv1d(C->2C, k=3, dilation d)(leaky_relu(x)) )      GEMM1 on tcgen05, 3 TMA taps
//   h   = h * (1 + scale[t]) + shift[t]                             FiLM at frame rate, fp32 regs
//   out = x + Conv1d(C->C, 1)(h)                                    GEMM2 on tcgen05, A = h in SMEM
//
// One CTA owns 128 consecutive time steps of one sequence.  The input tensor holds leaky_relu(x)
// (the producing epilogue stores it that way), so GEMM1's A operand is a plain TMA load; the raw x
// needed by the residual add is recovered in the epilogue with the exact inverse (x = a>=0?a:10a).
// GEMM1 runs in chunks of 64 value + 64 gate channels (weights are packed GLU-interleaved), each
// chunk's GLU/FiLM epilogue writes 64 channels of h as a K-major swizzled SMEM tile, which is
// k-block j of GEMM2 -- h never touches HBM.  Chunk j+1's MMAs overlap chunk j's epilogue
// (two TMEM accumulator buffers), GEMM2 k-block j is issued as soon as h chunk j is ready.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct ResblockParams {
  int L;            // time steps per sequence
  int dilation;
  int T;            // frames (film rows per batch element)
  int P;            // L / T
  int num_bands;
  int fmt;
  int out_fmt;      // storage format of the output (the consumer's operand format)
  int store_lrelu;
  const uint16_t* a16;   // [N, L, C] leaky_relu(x), 16-bit
  const float* b_conv;   // [2C] reference order (value half | gate half)
  const float* b_proj;   // [C]
  const float* film;     // [B, T, film_stride]: (1+scale | shift) in the first 2C columns
  int film_stride;
  uint16_t* out;         // [N, L, C]
};

template <int C>
struct RbCfg {
  static constexpr int KB = C >= 64 ? 64 : 32;        // k-block elements = swizzle span / 2
  static constexpr int ROWB = KB * 2;                  // bytes per smem row
  static constexpr int KPT = C / KB;                   // k-blocks per tap == GEMM1 chunks == GEMM2 k-blocks
  static constexpr int CH = KB;                        // value channels per GEMM1 chunk
  static constexpr int N1 = 2 * CH;                    // UMMA N of GEMM1
  static constexpr int NB1 = KPT > 1 ? 2 : 1;          // D1 accumulator buffers
  static constexpr int A_TILE = 128 * ROWB;
  static constexpr int B1_TILE = N1 * ROWB;
  static constexpr int B2_TILE = C * ROWB;
  static constexpr int STAGE = (A_TILE + B1_TILE) > B2_TILE ? (A_TILE + B1_TILE) : B2_TILE;
  static constexpr int B2_OFF = (B2_TILE > STAGE - A_TILE) ? 0 : A_TILE;
  static constexpr int STAGES = C == 256 ? 4 : (C == 128 ? 3 : 4);
  static constexpr int H_BYTES = KPT * A_TILE;
  static constexpr int D2_COL = NB1 * N1;
  static constexpr int TMEM_NEED = D2_COL + C;
  static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128
                                        : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE + H_BYTES + 512 + 1024;
  static_assert(STAGE % 1024 == 0 && A_TILE % 1024 == 0, "tiles must keep 1024B alignment");
};

template <int C>
__global__ void __launch_bounds__(192, 1)
resblock_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const ResblockParams p) {
  using K = RbCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* stages = smem;
  uint8_t* h_smem = smem + K::STAGES * K::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_smem + K::H_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = full + K::STAGES;          // [STAGES]
  uint64_t* d1_full = empty + K::STAGES;       // [2]
  uint64_t* d1_empty = d1_full + 2;            // [2]
  uint64_t* h_ready = d1_empty + 2;            // [KPT <= 4]
  uint64_t* d2_full = h_ready + 4;             // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l0 = blockIdx.x * 128, seq = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < K::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d1_full[b], 1);
      mbar_init(&d1_empty[b], 128);
    }
    for (int j = 0; j < 4; ++j) mbar_init(&h_ready[j], 128);
    mbar_init(d2_full, 1);
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
      int it = 0;
      auto acquire = [&](uint32_t bytes) -> uint8_t* {
        const int s = it % K::STAGES;
        mbar_wait(&empty[s], ((it / K::STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s], bytes);
        return stages + s * K::STAGE;
      };
      auto load_g2 = [&](int kk) {
        uint8_t* st = acquire(K::B2_TILE);
        tma_load_2d(st + K::B2_OFF, &tmW2, &full[it % K::STAGES], kk * K::KB, 0);
        ++it;
      };
      for (int j = 0; j < K::KPT; ++j) {
        for (int tap = 0; tap < 3; ++tap)
          for (int kk = 0; kk < K::KPT; ++kk) {
            uint8_t* st = acquire(K::A_TILE + K::B1_TILE);
            tma_load_3d(st, &tmX, &full[it % K::STAGES], kk * K::KB, l0 + (tap - 1) * p.dilation, seq);
            tma_load_2d(st + K::A_TILE, &tmW1, &full[it % K::STAGES], tap * C + kk * K::KB, j * K::N1);
            ++it;
          }
        if (j >= 1) load_g2(j - 1);
      }
      load_g2(K::KPT - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(p.fmt, K::N1);
      const uint32_t idesc2 = make_idesc_f16(p.fmt, C);
      int it = 0;
      auto issue_g2 = [&](int kk) {
        mbar_wait(&h_ready[kk], 0);
        const int s = it % K::STAGES;
        mbar_wait(&full[s], (it / K::STAGES) & 1);
        tc_fence_after();
        const uint64_t a_desc = make_kmajor_desc<K::ROWB>(smem_u32(h_smem + kk * K::A_TILE));
        const uint64_t b_desc = make_kmajor_desc<K::ROWB>(smem_u32(stages + s * K::STAGE + K::B2_OFF));
#pragma unroll
        for (int k = 0; k < K::KB / 16; ++k)
          umma_f16(tmem_base + K::D2_COL, a_desc + 2 * k, b_desc + 2 * k, idesc2, (kk | k) != 0);
        umma_commit(&empty[s]);
        ++it;
      };
      for (int j = 0; j < K::KPT; ++j) {
        const int b = j % K::NB1;
        if (j >= K::NB1) mbar_wait(&d1_empty[b], ((j / K::NB1) - 1) & 1);
        tc_fence_after();
        for (int step = 0; step < 3 * K::KPT; ++step) {
          const int s = it % K::STAGES;
          mbar_wait(&full[s], (it / K::STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(stages + s * K::STAGE);
          const uint64_t a_desc = make_kmajor_desc<K::ROWB>(st);
          const uint64_t b_desc = make_kmajor_desc<K::ROWB>(st + K::A_TILE);
#pragma unroll
          for (int k = 0; k < K::KB / 16; ++k)
            umma_f16(tmem_base + b * K::N1, a_desc + 2 * k, b_desc + 2 * k, idesc1, (step | k) != 0);
          umma_commit(&empty[s]);
          ++it;
        }
        umma_commit(&d1_full[b]);
        if (j >= 1) issue_g2(j - 1);
      }
      issue_g2(K::KPT - 1);
      umma_commit(d2_full);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int l = l0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int bidx = seq / p.num_bands;
    int t = l / p.P;
    if (t > p.T - 1) t = p.T - 1;
    const float* film = p.film + ((long long)bidx * p.T + t) * p.film_stride;
    const int fmt = p.fmt;

    // ---- epilogue 1: GLU + FiLM -> h (16-bit, swizzled K-major smem tile)
    for (int j = 0; j < K::KPT; ++j) {
      const int b = j % K::NB1;
      mbar_wait(&d1_full[b], (j / K::NB1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < K::CH / 32; ++cc) {
        uint32_t va[32], vg[32];
        tmem_ld32(lane_addr + b * K::N1 + cc * 32, va);
        tmem_ld32(lane_addr + b * K::N1 + K::CH + cc * 32, vg);
        tmem_ld_wait();
        const int ch0 = j * K::CH + cc * 32;
        const float4* ba = reinterpret_cast<const float4*>(p.b_conv + ch0);
        const float4* bg = reinterpret_cast<const float4*>(p.b_conv + C + ch0);
        const float4* fs = reinterpret_cast<const float4*>(film + ch0);
        const float4* fh = reinterpret_cast<const float4*>(film + C + ch0);
        uint8_t* hrow = h_smem + j * K::A_TILE + row * K::ROWB;
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {   // 8 channels = one 16-byte chunk
          float hv[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const int i4 = i8 * 2 + h4;
            const float4 A = __ldg(ba + i4), G = __ldg(bg + i4), S = __ldg(fs + i4), H = __ldg(fh + i4);
            const float av[4] = {A.x, A.y, A.z, A.w}, gv[4] = {G.x, G.y, G.z, G.w};
            const float sv[4] = {S.x, S.y, S.z, S.w}, tv[4] = {H.x, H.y, H.z, H.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(va[i4 * 4 + e]) + av[e];
              const float g = __uint_as_float(vg[i4 * 4 + e]) + gv[e];
              const float sg = __fdividef(1.0f, 1.0f + __expf(-g));
              hv[h4 * 4 + e] = fmaf(a * sg, sv[e], tv[e]);
            }
          }
          const int chunk = cc * 4 + i8;
          const int phys = K::ROWB == 128 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3));
          *reinterpret_cast<uint4*>(hrow + phys * 16) =
              make_uint4(pack2(hv[0], hv[1], fmt), pack2(hv[2], hv[3], fmt), pack2(hv[4], hv[5], fmt),
                         pack2(hv[6], hv[7], fmt));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the UMMA (async proxy)
      mbar_arrive(&h_ready[j]);
      mbar_arrive(&d1_empty[b]);
    }

    // ---- epilogue 2: residual add + bias -> 16-bit channels-last store
    mbar_wait(d2_full, 0);
    tc_fence_after();
    const bool valid = l < p.L;
    const long long roff = ((long long)seq * p.L + (valid ? l : 0)) * C;
    const uint4* xin = reinterpret_cast<const uint4*>(p.a16 + roff);
    uint4* dst = reinterpret_cast<uint4*>(p.out + roff);
#pragma unroll 1
    for (int cc = 0; cc < C / 32; ++cc) {
      uint32_t vd[32];
      tmem_ld32(lane_addr + K::D2_COL + cc * 32, vd);
      tmem_ld_wait();
      if (valid) {
        const float4* b2 = reinterpret_cast<const float4*>(p.b_proj + cc * 32);
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const uint4 xa = __ldg(xin + cc * 4 + i8);
          const uint32_t xw[4] = {xa.x, xa.y, xa.z, xa.w};
          const float4 B0 = __ldg(b2 + i8 * 2), B1 = __ldg(b2 + i8 * 2 + 1);
          const float bv[8] = {B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, B1.z, B1.w};
          uint32_t ow[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 xs = unpack2(xw[e2], fmt);
            float y0 = lrelu_inv(xs.x) + __uint_as_float(vd[i8 * 8 + e2 * 2]) + bv[e2 * 2];
            float y1 = lrelu_inv(xs.y) + __uint_as_float(vd[i8 * 8 + e2 * 2 + 1]) + bv[e2 * 2 + 1];
            if (p.store_lrelu) { y0 = lrelu(y0); y1 = lrelu(y1); }
            ow[e2] = pack2(y0, y1, p.out_fmt);
          }
          dst[cc * 4 + i8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, K::TMEM_COLS);
}

// ---------------------------------------------------------------------------- packing
// w_conv [2C][C][3] (value rows 0..C-1, gate rows C..2C-1)  ->  w1[row][tap*C + ci] with rows
// GLU-interleaved per chunk of CH channels: row = j*2CH + {0..CH-1: value ch j*CH+i | CH..: gate}.
// w_proj [C][C][1] -> w2[co][ci].  Layout of w_packed: w1 (2C*3C elems) then w2.
// Narrow stages (C <= 64, resblock2.cu): w1 is pre-scaled by 1/2 (exact; the kernel evaluates
// sigmoid through tanh(g/2) and folds the other 1/2 into the value half), and w2 is [C][2C] =
// [W_proj | I]: the identity block adds the residual x on the tensor core.
__global__ void pack_resblock_kernel(const float* __restrict__ w_conv, const float* __restrict__ w_proj, int C,
                                     int CH, int fmt, uint16_t* __restrict__ out) {
  const bool narrow = C <= 64;
  const int w2_cols = narrow ? 2 * C : C;
  const long long n1 = 2ll * C * 3 * C, total = n1 + (long long)C * w2_cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v;
    if (i < n1) {
      const int col = (int)(i % (3 * C)), row = (int)(i / (3 * C));
      const int tap = col / C, ci = col % C;
      const int j = row / (2 * CH), within = row % (2 * CH);
      const int src_row = within < CH ? (j * CH + within) : (C + j * CH + within - CH);
      v = w_conv[((long long)src_row * C + ci) * 3 + tap];
      if (narrow) v *= 0.5f;
    } else {
      const int col = (int)((i - n1) % w2_cols), row = (int)((i - n1) / w2_cols);
      v = col < C ? w_proj[(long long)row * C + col] : (col - C == row ? 1.0f : 0.0f);
    }
    out[i] = fmt == 0 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

int pack_resblock_launch(const float* w_conv, const float* w_proj, int C, int fmt, void* out, cudaStream_t stream) {
  B200_CHECK_ARG(C == 32 || C == 64 || C == 128 || C == 256, "resblock: C=%d unsupported (32/64/128/256)", C);
  const int CH = C >= 64 ? 64 : 32;
  pack_resblock_kernel<<<1024, 256, 0, stream>>>(w_conv, w_proj, C, CH, fmt, reinterpret_cast<uint16_t*>(out));
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int C>
static int launch_resblock(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                           const float* film, int film_stride, int N, int L, int dilation, int T, int num_bands,
                           int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  using K = RbCfg<C>;
  CUtensorMap tmX, tmW1, tmW2;
  B200_TRY(make_tmap_3d(&tmX, a16, C, L, N, (uint64_t)C * 2, (uint64_t)L * C * 2, K::KB, 128, K::ROWB));
  const uint16_t* w1 = reinterpret_cast<const uint16_t*>(w_packed);
  const uint16_t* w2 = w1 + 2ll * C * 3 * C;
  B200_TRY(make_tmap_2d(&tmW1, w1, 3 * C, 2 * C, (uint64_t)3 * C * 2, K::KB, K::N1, K::ROWB));
  B200_TRY(make_tmap_2d(&tmW2, w2, C, C, (uint64_t)C * 2, K::KB, C, K::ROWB));
  ResblockParams p{};
  p.L = L;
  p.dilation = dilation;
  p.T = T;
  p.P = L / T;
  p.num_bands = num_bands;
  p.fmt = fmt;
  p.out_fmt = out_fmt;
  p.store_lrelu = store_lrelu;
  p.a16 = reinterpret_cast<const uint16_t*>(a16);
  p.b_conv = b_conv;
  p.b_proj = b_proj;
  p.film = film;
  p.film_stride = film_stride;
  p.out = reinterpret_cast<uint16_t*>(out16);
  static bool configured[16] = {};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 15]) {
    B200_CUDA(cudaFuncSetAttribute(resblock_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    configured[dev & 15] = true;
  }
  dim3 grid(ceil_div(L, 128), N);
  resblock_kernel<C><<<grid, 192, K::SMEM, stream>>>(tmX, tmW1, tmW2, p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int resblock2_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream);

int resblock3_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                     const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                     int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream);

// B200VOC_RESBLOCK_V1=1 forces the first-generation one-tile-per-CTA kernel (A/B measurements).
static bool force_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VOC_RESBLOCK_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int resblock_launch(const void* a16, const void* w_packed, const float* b_conv, const float* b_proj,
                    const float* film, int film_stride, int N, int L, int C, int dilation, int T, int num_bands,
                    int fmt, int out_fmt, int store_lrelu, void* out16, cudaStream_t stream) {
  B200_CHECK_ARG(N > 0 && L > 0 && T > 0 && L % T == 0, "resblock: L=%d must be a multiple of T=%d", L, T);
  B200_CHECK_ARG(N % num_bands == 0, "resblock: N=%d not a multiple of num_bands=%d", N, num_bands);
  B200_CHECK_ARG(a16 != out16, "resblock: in-place is not supported (neighbour tiles read the halo)");
  switch (C) {
    case 32:
    case 64:
      // narrow stages: RAW input convention, weights packed for resblock2.cu only
      return resblock2_launch(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, C, dilation, T, num_bands, fmt,
                              out_fmt, store_lrelu, out16, stream);
    case 128:
    case 256:
      if (dilation <= 8 && !force_v1())
        return resblock3_launch(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, C, dilation, T, num_bands, fmt,
                                out_fmt, store_lrelu, out16, stream);
      if (C == 256)
        return launch_resblock<256>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt, out_fmt, store_lrelu, out16, stream);
      return launch_resblock<128>(a16, w_packed, b_conv, b_proj, film, film_stride, N, L, dilation, T, num_bands, fmt, out_fmt, store_lrelu, out16, stream);
  }
  set_error("resblock: C=%d unsupported (32/64/128/256)", C);
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
