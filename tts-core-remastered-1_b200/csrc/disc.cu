// Discriminator forward kernels (forward half): the three waveform critics of
// vocoder7/discriminators.py are stacks of STRIDED 1-D convolutions with tiny channel counts at the input end
// (1 -> 4 -> 16 -> 64 -> 256 [-> 1024]) followed by LeakyReLU(0.2):
//   * MultiPeriodDiscriminator  discriminators.py:8-60   Conv2d (5,1) stride (3,1) over [B, C, T/p, p]: a 1-D conv
//     along T/p, the p period columns are independent (here: P innermost columns the conv does not touch);
//   * MultiScaleDiscriminator   discriminators.py:63-108 Conv1d k in {15, 41, 41}, stride 2,2,2,1,1, on x and on
//     avg_pool1d(x, 4, 2, 1) (twice: both pooled scales are pooled from x, discriminators.py:99);
//   * MultiBandDiscriminator    discriminators.py:111-157 Conv1d k15 stride 2 on the four torch.chunk()s of the
//     TIME axis (discriminators.py:147).
// Every layer output AND every activation output is a returned feature map (feat_maps / out_feats lists), so the
// conv kernel writes both in one pass.  Weights are spectral-normalised: weight = weight_orig / (u . (W v))
// (torch.nn.utils.spectral_norm, eval-mode: no power iteration), evaluated once at load time on the GPU.
//
// First correct CUDA path for this row: fp32 direct convolution on the CUDA cores (each thread = one output
// position x CO output channels, weights of an 8-input-channel slab staged in shared memory).  The two wide
// layers (256 -> 1024, k41) are GEMM-shaped and belong on tcgen05 like the Generator's convolutions; that and the
// backward kernels are the rest of rank 4.
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

struct DiscConvParams {
  const float* x;
  const float* w;      // [Cout][Cin][K], already spectral-normalised
  const float* bias;   // [Cout]
  float* y_pre;        // conv + bias          (may be null)
  float* y_act;        // LeakyReLU(conv+bias) (may be null)
  int B, Cin, Cout, Lin, Lout, P, K, stride, pad;
  long long in_batch_stride;   // elements between batch items of x
  long long in_valid;          // elements of one (b, ci) row that exist; reads past it are zero (MPD's F.pad)
  float slope;
};

constexpr int kDcThreads = 128;
constexpr int kDcCi = 8;   // input channels per shared-memory weight slab

template <int CO>
__global__ void __launch_bounds__(kDcThreads) disc_conv_kernel(const DiscConvParams p) {
  extern __shared__ __align__(16) float w_s[];   // [kDcCi * K][CO]
  const long long pos = (long long)blockIdx.x * kDcThreads + threadIdx.x;   // flattened (lo, column)
  const int co0 = blockIdx.y * CO, b = blockIdx.z;
  const bool active = pos < (long long)p.Lout * p.P;
  const int lo = active ? (int)(pos / p.P) : 0, col = active ? (int)(pos - (long long)lo * p.P) : 0;
  const int li0 = lo * p.stride - p.pad;
  const long long chan_stride = (long long)p.Lin * p.P;
  const float* xb = p.x + (long long)b * p.in_batch_stride;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += kDcCi) {
    const int nci = min(kDcCi, p.Cin - ci0);
    __syncthreads();   // the previous slab has been consumed
    for (int i = threadIdx.x; i < nci * p.K * CO; i += kDcThreads) {
      const int c = i % CO, r = i / CO;   // r = cc * K + k
      const int cc = r / p.K, k = r - cc * p.K, co = co0 + c;
      w_s[i] = co < p.Cout ? __ldg(p.w + ((long long)co * p.Cin + ci0 + cc) * p.K + k) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int cc = 0; cc < nci; ++cc) {
        const float* xc = xb + (long long)(ci0 + cc) * chan_stride;
        const float* wr = w_s + cc * p.K * CO;
        for (int k = 0; k < p.K; ++k) {
          const int li = li0 + k;
          float xv = 0.f;
          if (li >= 0 && li < p.Lin) {
            const long long idx = (long long)li * p.P + col;
            if (idx < p.in_valid) xv = __ldg(xc + idx);
          }
#pragma unroll
          for (int c = 0; c < CO; ++c) acc[c] = fmaf(xv, wr[k * CO + c], acc[c]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const int co = co0 + c;
      if (co < p.Cout) {
        const float y = acc[c] + __ldg(p.bias + co);
        const long long o = (((long long)b * p.Cout + co) * p.Lout + lo) * p.P + col;
        if (p.y_pre) p.y_pre[o] = y;
        if (p.y_act) p.y_act[o] = y > 0.f ? y : p.slope * y;
      }
    }
  }
}

// Stride-1, single-column variant (the two layers that hold > 99 % of the critics' FLOPs, MSD's 64->256 and
// 256->1024 k15/k41 convolutions): a thread owns NP = 4 CONSECUTIVE output positions x CO = 16 channels, so that the
// four input samples it needs for tap k+1 are three of tap k's plus one new load (sliding window in registers):
// 64 FMAs per (1 global load + 4 shared-memory vector loads).
constexpr int kDcNp = 4;
template <int CO>
__global__ void __launch_bounds__(kDcThreads) disc_conv_s1_kernel(const DiscConvParams p) {
  extern __shared__ __align__(16) float w_s[];   // [kDcCi * K][CO]
  const int lo0 = (blockIdx.x * kDcThreads + threadIdx.x) * kDcNp;
  const int co0 = blockIdx.y * CO, b = blockIdx.z;
  const bool active = lo0 < p.Lout;
  const int lim = (int)min((long long)p.Lin, p.in_valid);   // samples of a row that exist
  const int base = lo0 - p.pad;                              // input index of (position lo0, tap 0)
  const float* xb = p.x + (long long)b * p.in_batch_stride;
  float acc[kDcNp][CO];
#pragma unroll
  for (int j = 0; j < kDcNp; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += kDcCi) {
    const int nci = min(kDcCi, p.Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < nci * p.K * CO; i += kDcThreads) {
      const int c = i % CO, r = i / CO;
      const int cc = r / p.K, k = r - cc * p.K, co = co0 + c;
      w_s[i] = co < p.Cout ? __ldg(p.w + ((long long)co * p.Cin + ci0 + cc) * p.K + k) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int cc = 0; cc < nci; ++cc) {
        const float* xc = xb + (long long)(ci0 + cc) * p.Lin;
        const float* wr = w_s + cc * p.K * CO;
        float xw[kDcNp];
#pragma unroll
        for (int j = 0; j < kDcNp; ++j) {
          const int li = base + j;
          xw[j] = (li >= 0 && li < lim) ? __ldg(xc + li) : 0.f;
        }
        for (int k = 0; k < p.K; ++k) {
          const int ln = base + k + kDcNp;                    // the sample tap k+1 adds to the window
          const float xn = (ln >= 0 && ln < lim) ? __ldg(xc + ln) : 0.f;
#pragma unroll
          for (int c = 0; c < CO; ++c) {
            const float wv = wr[k * CO + c];
#pragma unroll
            for (int j = 0; j < kDcNp; ++j) acc[j][c] = fmaf(xw[j], wv, acc[j][c]);
          }
#pragma unroll
          for (int j = 0; j + 1 < kDcNp; ++j) xw[j] = xw[j + 1];
          xw[kDcNp - 1] = xn;
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const int co = co0 + c;
      if (co < p.Cout) {
        const float bv = __ldg(p.bias + co);
        const long long o = ((long long)b * p.Cout + co) * p.Lout + lo0;
#pragma unroll
        for (int j = 0; j < kDcNp; ++j) {
          if (lo0 + j < p.Lout) {
            const float y = acc[j][c] + bv;
            if (p.y_pre) p.y_pre[o + j] = y;
            if (p.y_act) p.y_act[o + j] = y > 0.f ? y : p.slope * y;
          }
        }
      }
    }
  }
}

// Score layer: Conv(C -> 1, k3, stride 1) at the end of every critic stack (discriminators.py:27-31, 85-89, 134-138):
// one output channel, i.e. a reduction over Cin x K per position -- the generic kernel gives it ONE thread per position
// (0.79 ms for MSD's 1024-channel map).  Here a CTA owns 64 positions and 16 channel groups split the
// input channels (reads stay coalesced along time); partial sums are combined in shared memory in a fixed order
// (deterministic).
constexpr int kC1Pos = 64, kC1Groups = 16;
__global__ void __launch_bounds__(kC1Pos * kC1Groups) disc_conv_cout1_kernel(const DiscConvParams p) {
  __shared__ float part[kC1Groups][kC1Pos];
  const int px = threadIdx.x & (kC1Pos - 1), cg = threadIdx.x / kC1Pos, b = blockIdx.z;
  const long long pos = (long long)blockIdx.x * kC1Pos + px;          // flattened (lo, column)
  const bool active = pos < (long long)p.Lout * p.P;
  const int lo = active ? (int)(pos / p.P) : 0, col = active ? (int)(pos - (long long)lo * p.P) : 0;
  const long long chan_stride = (long long)p.Lin * p.P;
  const float* xb = p.x + (long long)b * p.in_batch_stride;
  float acc = 0.f;
  if (active) {
    for (int ci = cg; ci < p.Cin; ci += kC1Groups) {
      const float* xc = xb + (long long)ci * chan_stride;
      const float* wr = p.w + (long long)ci * p.K;
      for (int k = 0; k < p.K; ++k) {
        const int li = lo - p.pad + k;
        if (li >= 0 && li < p.Lin) {
          const long long idx = (long long)li * p.P + col;
          if (idx < p.in_valid) acc = fmaf(__ldg(xc + idx), __ldg(wr + k), acc);
        }
      }
    }
  }
  part[cg][px] = acc;
  __syncthreads();
  if (cg == 0 && active) {
    float y = __ldg(p.bias);
#pragma unroll
    for (int g = 0; g < kC1Groups; ++g) y += part[g][px];
    const long long o = (long long)b * p.Lout * p.P + pos;
    if (p.y_pre) p.y_pre[o] = y;
    if (p.y_act) p.y_act[o] = y > 0.f ? y : p.slope * y;
  }
}

template <int CO>
static int launch_disc_conv(const DiscConvParams& p, cudaStream_t st) {
  const long long npos = (long long)p.Lout * p.P;
  const size_t smem = (size_t)kDcCi * p.K * CO * sizeof(float);
  B200_CHECK_ARG(smem <= 48 * 1024, "disc_conv: kernel size %d too large for the weight slab", p.K);
  dim3 grid((unsigned)((npos + kDcThreads - 1) / kDcThreads), (unsigned)ceil_div(p.Cout, CO), (unsigned)p.B);
  disc_conv_kernel<CO><<<grid, kDcThreads, smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

static bool disc_s1_enabled() {
  static const bool on = [] { const char* e = getenv("B200VOC_DISC_GENERIC"); return !(e && e[0] == '1'); }();
  return on;
}
int disc_conv_launch(const DiscConvParams& p, cudaStream_t st) {
  if (p.Cout == 1 && p.stride == 1 && p.Cin >= 64 && disc_s1_enabled()) {
    const long long npos = (long long)p.Lout * p.P;
    dim3 grid((unsigned)((npos + kC1Pos - 1) / kC1Pos), 1, (unsigned)p.B);
    disc_conv_cout1_kernel<<<grid, kC1Pos * kC1Groups, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return B200VOC_OK;
  }
  if (p.stride == 1 && p.P == 1 && p.Cout >= 16 && disc_s1_enabled()) {
    constexpr int CO = 16;
    const size_t smem = (size_t)kDcCi * p.K * CO * sizeof(float);
    B200_CHECK_ARG(smem <= 48 * 1024, "disc_conv: kernel size %d too large for the weight slab", p.K);
    dim3 grid((unsigned)ceil_div(p.Lout, kDcThreads * kDcNp), (unsigned)ceil_div(p.Cout, CO), (unsigned)p.B);
    disc_conv_s1_kernel<CO><<<grid, kDcThreads, smem, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return B200VOC_OK;
  }
  if (p.Cout >= 16) return launch_disc_conv<16>(p, st);
  if (p.Cout >= 4) return launch_disc_conv<4>(p, st);
  return launch_disc_conv<1>(p, st);
}

// sigma = u . (W v) for W = weight_orig viewed as [rows][cols] (torch.nn.utils.spectral_norm.compute_weight with
// do_power_iteration=False), then w_out = weight_orig / sigma.  One CTA, fixed reduction order (deterministic).
// With freshly initialised u, v the sum nearly cancels (sigma ~ 1e-4 .. 1e-2 for the default init, so that fp32
// round-off of the reduction is up to 4e-5 of sigma and every later map inherits it): the products are fp32 like
// the reference's, the accumulation is fp64 -- a load-time kernel, its speed does not matter.
__global__ void __launch_bounds__(1024) sn_sigma_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                        const float* __restrict__ v, int rows, int cols,
                                                        float* __restrict__ sigma) {
  __shared__ double part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double mine = 0.0;
  for (int r = warp; r < rows; r += 32) {
    const float* wr = w + (long long)r * cols;
    double d = 0.0;
    for (int c = lane; c < cols; c += 32) d += (double)__ldg(wr + c) * (double)__ldg(v + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    mine += (double)__ldg(u + r) * d;   // every lane holds the same value
  }
  if (lane == 0) part[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 32; ++i) s += part[i];
    *sigma = (float)s;
  }
}
__global__ void __launch_bounds__(256) sn_scale_kernel(const float* __restrict__ w, const float* __restrict__ sigma,
                                                       long long n, float* __restrict__ out) {
  const float s = __ldg(sigma);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = w[i] / s;
}
int spectral_norm_launch(const float* w_orig, const float* u, const float* v, int rows, int cols, float* w_out,
                         float* sigma, cudaStream_t st) {
  sn_sigma_kernel<<<1, 1024, 0, st>>>(w_orig, u, v, rows, cols, sigma);
  B200_CUDA(cudaGetLastError());
  const long long n = (long long)rows * cols;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  sn_scale_kernel<<<blocks, 256, 0, st>>>(w_orig, sigma, n, w_out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

// Training-mode spectral norm (torch.nn.utils.spectral_norm with do_power_iteration, n_power_iterations = 1: what every
// critic forward of vocoder7/trainer.py:86-115 runs):  v <- normalize(W^T u),  u <- normalize(W v)  (x / max(||x||, eps)),
// sigma = u . (W v) = ||W v||^2 / max(||W v||, eps),  weight = W / sigma; u and v are updated in place.  This runs once per
// layer per forward, so unlike the load-time sigma kernel it is spread over the chip: two passes over W (the largest,
// 1024 x 10496, is 43 MB) plus two single-CTA vector normalisations.  fp32 products, fp64 accumulation.
__global__ void __launch_bounds__(256) sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u, int rows,
                                                      int cols, float* __restrict__ t) {
  // W^T u = sum_r u[r] W[r][:].  CTA = 32 columns x 8 row phases (a warp reads 128 contiguous bytes of a row); the
  // phases are combined through shared memory in a fixed order.  One column per thread over all rows (the first
  // version) left the 1024 x 10496 matrix of MSD's widest layer to 41 latency-bound CTAs: 240 us.
  __shared__ double part[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double acc = 0.0;
  if (c < cols)
    for (int r = ry; r < rows; r += 8) acc += (double)__ldg(u + r) * (double)__ldg(w + (long long)r * cols + c);
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][cx];
    t[c] = (float)s;
  }
}
__global__ void __launch_bounds__(256) sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ v, int rows,
                                                     int cols, float* __restrict__ s) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // one warp per row
  if (r >= rows) return;
  const float* wr = w + (long long)r * cols;
  double d = 0.0;
  for (int c = lane; c < cols; c += 32) d += (double)__ldg(wr + c) * (double)__ldg(v + c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) s[r] = (float)d;
}
// out = x / max(||x||, eps); sigma_out (optional) = ||x||^2 / max(||x||, eps) = out . x.  One CTA, fixed order.
__global__ void __launch_bounds__(1024) sn_normalize_kernel(const float* __restrict__ x, int n, float eps,
                                                            float* __restrict__ out, float* __restrict__ sigma_out) {
  __shared__ double part[32];
  __shared__ double total;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) { const double xv = (double)x[i]; acc += xv * xv; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 32; ++i) s += part[i];
    total = s;
  }
  __syncthreads();
  const float norm = (float)sqrt(total);
  const float den = fmaxf(norm, eps);
  for (int i = threadIdx.x; i < n; i += 1024) out[i] = x[i] / den;
  if (sigma_out && threadIdx.x == 0) *sigma_out = (float)(total / (double)den);
}
// The same five steps in ONE single-CTA launch for the small layers (rows * cols <= 32 k: most MPD / MBD layers, the
// narrow MSD ones): a critic forward in .train() mode runs this for every layer, and for the small ones the five launches
// cost more than their arithmetic (MPD: 25 layers).  fp32 products, fp64 accumulation, fixed order, as above.
constexpr int kSnSmallMax = 32 * 1024;
__device__ __forceinline__ double sn_block_sum(double v, double* sh) {   // 1024 threads; every thread gets the total
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                   // sh may still be read from the previous use
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < 32; ++i) s += sh[i];
  return s;
}
__global__ void __launch_bounds__(1024) sn_train_small_kernel(const float* __restrict__ w, float* __restrict__ u,
                                                              float* __restrict__ v, int rows, int cols, float eps,
                                                              float* __restrict__ w_out, float* __restrict__ sigma,
                                                              float* __restrict__ scratch) {
  __shared__ double sh[32];
  float* t = scratch;            // [cols]
  float* sv = scratch + cols;    // [rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // v <- normalize(W^T u)
  double sq = 0.0;
  for (int c = threadIdx.x; c < cols; c += 1024) {
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) acc += (double)u[r] * (double)__ldg(w + (long long)r * cols + c);
    const float tf = (float)acc;
    t[c] = tf;
    sq += (double)tf * (double)tf;
  }
  double tot = sn_block_sum(sq, sh);
  float den = fmaxf((float)sqrt(tot), eps);
  for (int c = threadIdx.x; c < cols; c += 1024) v[c] = t[c] / den;
  __syncthreads();
  // u <- normalize(W v)
  for (int r = warp; r < rows; r += 32) {
    const float* wr = w + (long long)r * cols;
    double d = 0.0;
    for (int c = lane; c < cols; c += 32) d += (double)__ldg(wr + c) * (double)v[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) sv[r] = (float)d;
  }
  __syncthreads();
  sq = 0.0;
  for (int r = threadIdx.x; r < rows; r += 1024) { const double x = (double)sv[r]; sq += x * x; }
  tot = sn_block_sum(sq, sh);
  den = fmaxf((float)sqrt(tot), eps);
  for (int r = threadIdx.x; r < rows; r += 1024) u[r] = sv[r] / den;
  const float sg = (float)(tot / (double)den);
  if (threadIdx.x == 0) *sigma = sg;
  // weight = W / sigma
  const int n = rows * cols;
  for (int i = threadIdx.x; i < n; i += 1024) w_out[i] = __ldg(w + i) / sg;
}
int spectral_norm_train_launch(const float* w_orig, float* u, float* v, int rows, int cols, float eps, float* w_out,
                               float* sigma, float* scratch, cudaStream_t st) {
  if ((long long)rows * cols <= kSnSmallMax) {
    sn_train_small_kernel<<<1, 1024, 0, st>>>(w_orig, u, v, rows, cols, eps, w_out, sigma, scratch);
    B200_CUDA(cudaGetLastError());
    return B200VOC_OK;
  }
  float* t = scratch;            // [cols]
  float* s = scratch + cols;     // [rows]
  sn_wt_u_kernel<<<(cols + 31) / 32, 256, 0, st>>>(w_orig, u, rows, cols, t);
  sn_normalize_kernel<<<1, 1024, 0, st>>>(t, cols, eps, v, nullptr);
  sn_w_v_kernel<<<(rows + 7) / 8, 256, 0, st>>>(w_orig, v, rows, cols, s);
  sn_normalize_kernel<<<1, 1024, 0, st>>>(s, rows, eps, u, sigma);
  B200_CUDA(cudaGetLastError());
  const long long n = (long long)rows * cols;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  sn_scale_kernel<<<blocks, 256, 0, st>>>(w_orig, sigma, n, w_out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

// F.avg_pool1d(x, kernel_size=4, stride=2, padding=1) (count_include_pad=True: the divisor is always 4).
__global__ void __launch_bounds__(256) avg_pool_k4s2p1_kernel(const float* __restrict__ x, long long rows, int Lin,
                                                              int Lout, float* __restrict__ y) {
  const long long total = rows * Lout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Lout;
    const int j = (int)(i - r * Lout);
    const float* xr = x + r * Lin;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = 2 * j - 1 + k;
      s += (t >= 0 && t < Lin) ? __ldg(xr + t) : 0.f;
    }
    y[i] = s * 0.25f;
  }
}
int avg_pool_launch(const float* x, long long rows, int Lin, float* y, cudaStream_t st) {
  const int Lout = (Lin + 2 - 4) / 2 + 1;
  const long long total = rows * Lout;
  const int blocks = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
  avg_pool_k4s2p1_kernel<<<blocks, 256, 0, st>>>(x, rows, Lin, Lout, y);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (include/b200voc.h)
extern "C" {

int b200voc_disc_conv_out_len(int Lin, int K, int stride, int pad) {
  if (Lin <= 0 || K <= 0 || stride <= 0 || pad < 0 || Lin + 2 * pad < K) return 0;
  return (Lin + 2 * pad - K) / stride + 1;
}

int b200voc_disc_conv(const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int Lin, int P,
                      int K, int stride, int pad, int64_t in_batch_stride, int64_t in_valid, float slope,
                      float* y_pre, float* y_act, void* stream) {
  B200_CHECK_ARG(x && w && bias && (y_pre || y_act), "disc_conv: null argument");
  B200_CHECK_ARG(B > 0 && Cin > 0 && Cout > 0 && P > 0 && K > 0 && stride > 0 && pad >= 0,
                 "disc_conv: bad shape (B=%d Cin=%d Cout=%d P=%d K=%d stride=%d pad=%d)", B, Cin, Cout, P, K, stride, pad);
  const int Lout = b200voc_disc_conv_out_len(Lin, K, stride, pad);
  B200_CHECK_ARG(Lout > 0, "disc_conv: input of %d rows is shorter than the kernel (K=%d, pad=%d)", Lin, K, pad);
  B200_CHECK_ARG(B <= 65535 && b200::ceil_div(Cout, 16) <= 65535, "disc_conv: batch / channel count exceeds the grid limits");
  b200::DiscConvParams p{};
  p.x = x; p.w = w; p.bias = bias; p.y_pre = y_pre; p.y_act = y_act;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Lin = Lin; p.Lout = Lout; p.P = P; p.K = K; p.stride = stride; p.pad = pad;
  p.in_batch_stride = in_batch_stride > 0 ? in_batch_stride : (long long)Cin * Lin * P;
  p.in_valid = in_valid > 0 ? in_valid : (long long)Lin * P;
  p.slope = slope;
  return b200::disc_conv_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_spectral_norm_weight(const float* w_orig, const float* u, const float* v, int rows, int cols,
                                 float* w_out, float* sigma_out, void* stream) {
  B200_CHECK_ARG(w_orig && u && v && w_out && sigma_out, "spectral_norm_weight: null argument");
  B200_CHECK_ARG(rows > 0 && cols > 0, "spectral_norm_weight: bad shape (%d x %d)", rows, cols);
  return b200::spectral_norm_launch(w_orig, u, v, rows, cols, w_out, sigma_out, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_spectral_norm_train(const float* w_orig, float* u, float* v, int rows, int cols, float eps, float* w_out,
                                float* sigma_out, float* scratch, void* stream) {
  B200_CHECK_ARG(w_orig && u && v && w_out && sigma_out && scratch, "spectral_norm_train: null argument");
  B200_CHECK_ARG(rows > 0 && cols > 0 && eps >= 0.f, "spectral_norm_train: bad shape (%d x %d)", rows, cols);
  return b200::spectral_norm_train_launch(w_orig, u, v, rows, cols, eps, w_out, sigma_out, scratch,
                                          reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_avg_pool1d_k4s2p1(const float* x, int64_t rows, int Lin, float* y, void* stream) {
  B200_CHECK_ARG(x && y, "avg_pool1d: null argument");
  B200_CHECK_ARG(rows > 0 && Lin >= 2, "avg_pool1d: bad shape (rows=%lld, L=%d)", (long long)rows, Lin);
  return b200::avg_pool_launch(x, rows, Lin, y, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
