// Critic backward kernels (the discriminator half of the training step,
// vocoder7/trainer.py:86-115: d_loss.backward() runs autograd through the three critics of
// vocoder7/discriminators.py:8-157).  For one spectral-normalised convolution layer
//     y = conv(x, W) + b,  W = W_orig / sigma,  sigma = u . (W_orig v),  a = LeakyReLU(y)
// with upstream gradients gy (every conv map is a returned feature), ga (every activation map too) and the dgrad of
// the next layer gn:
//     g   = gy + (ga + gn) * (y > 0 ? 1 : slope)                                disc_lrelu_bwd_kernel
//     db  = sum_{b, l} g                                                        disc_bias_grad_kernel
//     dx[b,ci,li] = sum_{co,k} W[co,ci,k] g[b,co,lo],  lo * stride - pad + k = li   (dgrad)
//     dW[co,ci,k] = sum_{b,lo} g[b,co,lo] x[b,ci,lo * stride - pad + k]             (wgrad)
//     dW_orig = (dW - <dW, W> u v^T) / sigma                                    sn_bwd_*_kernel
// (u, v are constants of the graph, as in torch.nn.utils.spectral_norm: the power iteration runs under no_grad).
//
// Where the contraction is GEMM-shaped (MSD's stride-1 64 -> 256 and 256 -> 1024 layers, 97 % of the FLOPs) it runs on
// the tensor cores:
//   * dgrad of a stride-1 layer IS a stride-1 convolution of g with the transposed, tap-flipped weights
//     (disc_flip_weight_kernel), so it goes through the forward implicit-GEMM kernel of disc_gemm.cu unchanged;
//   * wgrad is a plain K-major x K-major GEMM over the [B, C, L] maps: positions (contiguous in memory) are the K axis,
//     M = Cout, N = (ci, k) -- the B operand is the im2col of x, one time-shifted copy of a channel row per tap.  Both
//     operands are split bf16 (x = hi + lo, 16 significant bits at fp32 range: gradients are tiny, spectral-norm
//     weights large) laid out as [g_hi | g_lo | g_hi] x [x_hi | x_hi | x_lo] along K, so ONE pass of the split GEMM of
//     small_kernels.cu (splitgemm_kernel, tcgen05, fp32 accumulation in TMEM) yields dW in the weight's own layout.
// Everything else (strided / narrow layers, Conv2d (k,1) of MPD) is fp32 on the CUDA cores; reductions run in a fixed
// order (deterministic, no atomics).
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace b200 {

int disc_gemm_launch(const float* x, const void* w_split, const float* bias, int B, int Cin, int Cout, int L, int K,
                     int pad, float slope, float* y_pre, float* y_act, void* workspace, cudaStream_t st);
int disc_pack_w_launch(const float* w, int Cout, int Cin, int K, void* out, cudaStream_t st);
int splitgemm_f32_launch(const void* A, const void* W, const float* bias, int M, int N, long long Kdim, int fmt,
                         float* out, cudaStream_t st);

struct DiscBwdParams {
  const float* x;      // layer input (wgrad)
  const float* g;      // [B, Cout, Lout, P] gradient of the conv output
  const float* w;      // [Cout][Cin][K] spectral-normalised weight (dgrad)
  float* dx;           // dgrad output, laid out like the layer input
  float* dw;           // wgrad output [Z][Cout][Cin][K] partial sums (Z = gridDim.z)
  int B, Cin, Cout, Lin, Lout, P, K, stride, pad;
  long long in_batch_stride;   // elements between batch items of x / dx
  long long in_valid;          // elements of one (b, ci) row that exist (MPD's F.pad: the rest is padding)
  int accumulate;              // dgrad: add to dx instead of overwriting it
  long long r_per_z;           // wgrad: positions (b, lo, column) per grid z slice
};

constexpr int kDbThreads = 128;
constexpr int kDbCo = 8;   // output channels per shared-memory weight slab (dgrad) / per CTA (wgrad)

// ---------------------------------------------------------------------------------------------------- dgrad
// thread = one input position (li, column) x CI input channels; the taps that reach li are k = k0, k0 + stride, ...
// with k0 = (li + pad) mod stride, read from output position lo = (li + pad - k) / stride.
template <int CI>
__global__ void __launch_bounds__(kDbThreads) disc_dgrad_kernel(const DiscBwdParams p) {
  extern __shared__ __align__(16) float w_s[];   // [kDbCo * K][CI]
  const long long pos = (long long)blockIdx.x * kDbThreads + threadIdx.x;
  const int ci0 = blockIdx.y * CI, b = blockIdx.z;
  const bool active = pos < (long long)p.Lin * p.P;
  const int li = active ? (int)(pos / p.P) : 0, col = active ? (int)(pos - (long long)li * p.P) : 0;
  const int k0 = (li + p.pad) % p.stride, lo0 = (li + p.pad) / p.stride;
  float acc[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) acc[c] = 0.f;

  for (int co0 = 0; co0 < p.Cout; co0 += kDbCo) {
    const int nco = min(kDbCo, p.Cout - co0);
    __syncthreads();
    for (int i = threadIdx.x; i < nco * p.K * CI; i += kDbThreads) {
      const int c = i % CI, r = i / CI;   // r = cc * K + k
      const int cc = r / p.K, k = r - cc * p.K, ci = ci0 + c;
      w_s[i] = ci < p.Cin ? __ldg(p.w + ((long long)(co0 + cc) * p.Cin + ci) * p.K + k) : 0.f;
    }
    __syncthreads();
    if (active) {
      for (int cc = 0; cc < nco; ++cc) {
        const float* gc = p.g + ((long long)b * p.Cout + co0 + cc) * p.Lout * p.P + col;
        const float* wr = w_s + cc * p.K * CI;
        for (int k = k0, lo = lo0; k < p.K && lo >= 0; k += p.stride, --lo) {
          if (lo < p.Lout) {
            const float gv = __ldg(gc + (long long)lo * p.P);
#pragma unroll
            for (int c = 0; c < CI; ++c) acc[c] = fmaf(gv, wr[k * CI + c], acc[c]);
          }
        }
      }
    }
  }
  if (active && pos < p.in_valid) {
#pragma unroll
    for (int c = 0; c < CI; ++c) {
      const int ci = ci0 + c;
      if (ci < p.Cin) {
        float* o = p.dx + (long long)b * p.in_batch_stride + (long long)ci * p.Lin * p.P + pos;
        *o = p.accumulate ? *o + acc[c] : acc[c];
      }
    }
  }
}

template <int CI>
static int launch_dgrad(const DiscBwdParams& p, cudaStream_t st) {
  const long long npos = (long long)p.Lin * p.P;
  const size_t smem = (size_t)kDbCo * p.K * CI * sizeof(float);
  B200_CHECK_ARG(smem <= 48 * 1024, "disc_conv_dgrad: kernel size %d too large for the weight slab", p.K);
  dim3 grid((unsigned)((npos + kDbThreads - 1) / kDbThreads), (unsigned)ceil_div(p.Cin, CI), (unsigned)p.B);
  disc_dgrad_kernel<CI><<<grid, kDbThreads, smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}
int disc_dgrad_launch(const DiscBwdParams& p, cudaStream_t st) {
  // 16 input channels per thread reuse every g load 16 times, but a 16-channel map of a few thousand positions is then
  // under 100 CTAs: take fewer channels per thread until the grid fills the chip (g stays in L1 / L2)
  const long long blocks = (((long long)p.Lin * p.P + kDbThreads - 1) / kDbThreads) * p.B;
  if (p.Cin >= 16 && blocks * ceil_div(p.Cin, 16) >= 296) return launch_dgrad<16>(p, st);
  if (p.Cin >= 4 && blocks * ceil_div(p.Cin, 4) >= 296) return launch_dgrad<4>(p, st);
  if (p.Cin >= 16 && blocks * p.Cin < 296) return launch_dgrad<4>(p, st);   // tiny map: grid size does not matter
  return launch_dgrad<1>(p, st);
}

// ---------------------------------------------------------------------------------------------------- wgrad
// CTA = (input channel ci, 8 output channels, z slice of the positions); thread = (tap k, phase): walks the slice's
// positions phase, phase + nph, ... and keeps 8 partial sums; the phases are combined through shared memory in a
// fixed order.  Slices exist so that the first layers (a handful of weights, 1e5 positions) still fill the chip; they
// are summed by disc_sum_slices_kernel.
constexpr int kDwThreads = 256;
__global__ void __launch_bounds__(kDwThreads) disc_wgrad_kernel(const DiscBwdParams p) {
  __shared__ float part[kDwThreads * kDbCo];
  const int ci = blockIdx.x, co0 = blockIdx.y * kDbCo;
  const int nph = kDwThreads / p.K;
  const int k = threadIdx.x % p.K, phase = threadIdx.x / p.K;
  const long long R = (long long)p.B * p.Lout * p.P;
  const long long r_begin = (long long)blockIdx.z * p.r_per_z;
  const long long r_end = min(R, r_begin + p.r_per_z);
  const long long per_b = (long long)p.Lout * p.P;
  float acc[kDbCo];
#pragma unroll
  for (int c = 0; c < kDbCo; ++c) acc[c] = 0.f;
  if (phase < nph && r_begin < r_end) {
    // the slice [r_begin, r_end) of the flattened (b, lo, column) positions, batch item by batch item: 32-bit index
    // arithmetic inside an item (per_b < 2^31 is checked by the launcher), this thread takes every nph-th position
    const int b_first = (int)(r_begin / per_b), b_last = (int)((r_end - 1) / per_b);
    const int per_b32 = (int)per_b;
    long long r = r_begin + phase;                       // this thread's next position (global)
    for (int b = b_first; b <= b_last; ++b) {
      const long long base = (long long)b * per_b;
      const int q_end = (int)min((long long)per_b32, r_end - base);
      if (r >= base + q_end) continue;
      int q = (int)(r - base);
      const float* xb = p.x + (long long)b * p.in_batch_stride + (long long)ci * p.Lin * p.P;
      const float* gb = p.g + ((long long)b * p.Cout + co0) * per_b;
      const int valid = (int)min(p.in_valid, (long long)p.Lin * p.P);
      for (; q < q_end; q += nph) {
        int lo = q, col = 0;
        if (p.P != 1) { lo = q / p.P; col = q - lo * p.P; }
        const int li = lo * p.stride - p.pad + k;
        float xv = 0.f;
        if (li >= 0 && li < p.Lin) {
          const int idx = li * p.P + col;
          if (idx < valid) xv = __ldg(xb + idx);
        }
#pragma unroll
        for (int c = 0; c < kDbCo; ++c)
          if (co0 + c < p.Cout) acc[c] = fmaf(xv, __ldg(gb + (long long)c * per_b + q), acc[c]);
      }
      r = base + q;
    }
  }
#pragma unroll
  for (int c = 0; c < kDbCo; ++c) part[threadIdx.x * kDbCo + c] = acc[c];
  __syncthreads();
  // every (c, k) is summed over the phases in a fixed order
  for (int t = threadIdx.x; t < kDbCo * p.K; t += kDwThreads) {
    const int c = t / p.K, kk = t - c * p.K;
    if (co0 + c < p.Cout) {
      float s = 0.f;
      for (int ph = 0; ph < nph; ++ph) s += part[(ph * p.K + kk) * kDbCo + c];
      p.dw[(long long)blockIdx.z * p.Cout * p.Cin * p.K + ((long long)(co0 + c) * p.Cin + ci) * p.K + kk] = s;
    }
  }
}
// Score layers (Conv(C -> 1, k3), the last layer of every stack) and other layers with a handful of output channels and
// taps: the general kernel gives such a layer one useful lane in eight and two dependent loads per FMA (190 us for MSD's
// 1024-channel map).  Here CTA = one input channel, thread = positions tid, tid + 256, ... (coalesced along time), CO x KK
// accumulators per thread, then a fixed-order block reduction.
template <int CO, int KK>
__global__ void __launch_bounds__(kDwThreads) disc_wgrad_small_kernel(const DiscBwdParams p) {
  __shared__ float red[kDwThreads / 32][CO * KK];
  const int ci = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_b = p.Lout * p.P;
  const int valid = (int)min(p.in_valid, (long long)p.Lin * p.P);
  float acc[CO][KK];
#pragma unroll
  for (int c = 0; c < CO; ++c)
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[c][k] = 0.f;
  for (int b = 0; b < p.B; ++b) {
    const float* xb = p.x + (long long)b * p.in_batch_stride + (long long)ci * p.Lin * p.P;
    const float* gb = p.g + (long long)b * p.Cout * per_b;
    for (int q = threadIdx.x; q < per_b; q += kDwThreads) {
      int lo = q, col = 0;
      if (p.P != 1) { lo = q / p.P; col = q - lo * p.P; }
      float xv[KK];
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const int li = lo * p.stride - p.pad + k;
        const int idx = li * p.P + col;
        xv[k] = (k < p.K && li >= 0 && li < p.Lin && idx < valid) ? __ldg(xb + idx) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        if (c < p.Cout) {
          const float gv = __ldg(gb + (long long)c * per_b + q);
#pragma unroll
          for (int k = 0; k < KK; ++k) acc[c][k] = fmaf(gv, xv[k], acc[c][k]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CO; ++c)
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      float v = acc[c][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][c * KK + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < CO * KK) {
    const int c = threadIdx.x / KK, k = threadIdx.x - c * KK;
    if (c < p.Cout && k < p.K) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kDwThreads / 32; ++w) sum += red[w][threadIdx.x];
      p.dw[((long long)c * p.Cin + ci) * p.K + k] = sum;
    }
  }
}
__global__ void __launch_bounds__(256) disc_sum_slices_kernel(const float* __restrict__ part, long long n, int Z,
                                                              float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < Z; ++z) s += part[(long long)z * n + i];
    out[i] = s;
  }
}
static int wgrad_slices(int B, int Cin, int Cout, int Lout, int P) {
  const long long ctas = (long long)Cin * ceil_div(Cout, kDbCo);
  const long long R = (long long)B * Lout * P;
  long long z = ctas >= 592 ? 1 : (592 + ctas - 1) / ctas;         // ~4 CTAs per SM
  const long long zmax = (R + 255) / 256;                         // at least 256 positions per slice
  if (z > zmax) z = zmax;
  if (z < 1) z = 1;
  if (z > 4096) z = 4096;
  return (int)z;
}
int disc_wgrad_launch(DiscBwdParams p, float* dw, float* scratch, cudaStream_t st) {
  B200_CHECK_ARG(p.K <= kDwThreads, "disc_conv_wgrad: kernel size %d not supported (max %d)", p.K, kDwThreads);
  B200_CHECK_ARG((long long)p.Lin * p.P < (1ll << 31) && (long long)p.Lout * p.P < (1ll << 31), "disc_conv_wgrad: map too long");
  if (p.Cout <= 4 && p.K <= 5 && p.Cin >= 64) {          // few outputs per input channel, many input channels
    p.dw = dw;
    if (p.Cout == 1 && p.K <= 3) disc_wgrad_small_kernel<1, 3><<<p.Cin, kDwThreads, 0, st>>>(p);
    else disc_wgrad_small_kernel<4, 5><<<p.Cin, kDwThreads, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return B200VOC_OK;
  }
  const int Z = wgrad_slices(p.B, p.Cin, p.Cout, p.Lout, p.P);
  const long long R = (long long)p.B * p.Lout * p.P, n = (long long)p.Cout * p.Cin * p.K;
  p.r_per_z = (R + Z - 1) / Z;
  p.dw = Z == 1 ? dw : scratch;
  B200_CHECK_ARG(Z == 1 || scratch, "disc_conv_wgrad: scratch buffer missing");
  dim3 grid((unsigned)p.Cin, (unsigned)ceil_div(p.Cout, kDbCo), (unsigned)Z);
  disc_wgrad_kernel<<<grid, kDwThreads, 0, st>>>(p);
  B200_CUDA(cudaGetLastError());
  if (Z > 1) {
    const int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    disc_sum_slices_kernel<<<blocks, 256, 0, st>>>(scratch, n, Z, dw);
    B200_CUDA(cudaGetLastError());
  }
  return B200VOC_OK;
}

// ---------------------------------------------------------------------------------------------------- bias gradient
// One CTA per output channel; threads stride over (b, position), fixed-order tree in shared memory.
__global__ void __launch_bounds__(256) disc_bias_grad_kernel(const float* __restrict__ g, int B, int Cout, long long per_c,
                                                             float* __restrict__ db) {
  __shared__ float part[256];
  const int co = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* gp = g + ((long long)b * Cout + co) * per_c;
    for (long long i = threadIdx.x; i < per_c; i += 256) s += gp[i];
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) db[co] = part[0];
}

// ---------------------------------------------------------------------------------------------------- LeakyReLU
// g = gy + (ga + gn) * (y > 0 ? 1 : slope): the three gradient sources of a conv map (any of them may be absent).
__global__ void __launch_bounds__(256) disc_lrelu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy,
                                                             const float* __restrict__ ga, const float* __restrict__ gn,
                                                             float slope, long long n, float* __restrict__ g) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    if (ga) a += ga[i];
    if (gn) a += gn[i];
    float r = (ga || gn) ? a * (y[i] > 0.f ? 1.f : slope) : 0.f;
    if (gy) r += gy[i];
    g[i] = r;
  }
}

// ---------------------------------------------------------------------------------------------------- avg_pool1d(4, 2, 1)
// dx[t] = 0.25 * sum of gy[j] over the windows j that contain t (2j - 1 <= t <= 2j + 2).
__global__ void __launch_bounds__(256) avg_pool_bwd_kernel(const float* __restrict__ gy, long long rows, int Lin, int Lout,
                                                           float* __restrict__ dx) {
  const long long total = rows * Lin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Lin;
    const int t = (int)(i - r * Lin);
    const float* gr = gy + r * Lout;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int num = t + 1 - k;          // 2j
      if (num >= 0 && !(num & 1) && (num >> 1) < Lout) s += gr[num >> 1];
    }
    dx[i] = 0.25f * s;
  }
}

// ---------------------------------------------------------------------------------------------------- spectral norm
// dW_orig = (dW - <dW, W> u v^T) / sigma with W = W_orig / sigma.  The inner product runs over up to 10.7 M elements:
// up to 256 CTAs write fp64 partial sums, every CTA of the second kernel re-adds them in the same order.
constexpr int kSnParts = 256;
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ dw, const float* __restrict__ w,
                                                         long long n, double* __restrict__ parts) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    s += (double)dw[i] * (double)w[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) parts[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* __restrict__ dw, const float* __restrict__ u,
                                                           const float* __restrict__ v, const float* __restrict__ sigma,
                                                           const double* __restrict__ parts, int nparts, int rows, int cols,
                                                           float* __restrict__ out) {
  __shared__ float dot_s;
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int i = 0; i < nparts; ++i) d += parts[i];
    dot_s = (float)d;
  }
  __syncthreads();
  const float dot = dot_s, inv = 1.f / __ldg(sigma);
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    out[i] = (dw[i] - dot * __ldg(u + r) * __ldg(v + c)) * inv;
  }
}

// ---------------------------------------------------------------------------------------------------- tensor-core paths
// wt[ci][co][k] = w[co][ci][K - 1 - k]: the weight of the stride-1 convolution that IS the dgrad.
__global__ void __launch_bounds__(256) disc_flip_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int K,
                                                               float* __restrict__ wt) {
  const long long n = (long long)Cout * Cin * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long r = i / K;
    const int co = (int)(r % Cout), ci = (int)(r / Cout);
    wt[i] = w[((long long)co * Cin + ci) * K + (K - 1 - k)];
  }
}

__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
// A operand of the wgrad GEMM: a3[co][s * Pp + b * Lp + lo], segments s = (g_hi, g_lo, g_hi); lo >= Lout is zero.
__global__ void __launch_bounds__(256) disc_pack_g3_kernel(const float* __restrict__ g, int Bc, int Cout, int Lout, int Lp,
                                                           long long g_batch_stride, uint16_t* __restrict__ a3) {
  const long long Pp = (long long)Bc * Lp, Kdim = 3 * Pp, total = (long long)Cout * Pp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Pp);
    const long long col = i - (long long)co * Pp;
    const int b = (int)(col / Lp), lo = (int)(col - (long long)b * Lp);
    const float v = lo < Lout ? g[(long long)b * g_batch_stride + (long long)co * Lout + lo] : 0.f;
    uint16_t hi, l;
    split_bf16(v, hi, l);
    uint16_t* row = a3 + (long long)co * Kdim + col;
    row[0] = hi;
    row[Pp] = l;
    row[2 * Pp] = hi;
  }
}
// B operand: w3[ci * K + k][s * Pp + b * Lp + lo] = x[b, ci, lo - pad + k], segments (x_hi, x_hi, x_lo).  A thread packs 8
// consecutive positions of one row (Lp is a multiple of 64, so they share b) and writes three 16-byte vectors.
__global__ void __launch_bounds__(256) disc_pack_im2col3_kernel(const float* __restrict__ x, int Bc, int Cin, int Lin,
                                                                int Lout, int Lp, int K, int pad, long long x_batch_stride,
                                                                uint16_t* __restrict__ w3) {
  const long long Pp = (long long)Bc * Lp, Kdim = 3 * Pp;
  const int groups = (int)(Pp >> 3);                                  // 8-position groups per row
  const long long total = (long long)Cin * K * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int rowi = (int)(i / groups);                               // ci * K + k
    const int col = (int)(i - (long long)rowi * groups) << 3;
    const int ci = rowi / K, k = rowi - ci * K;
    const int b = col / Lp, lo0 = col - b * Lp;
    const float* xr = x + (long long)b * x_batch_stride + (long long)ci * Lin;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      uint16_t hh[2], ll[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int lo = lo0 + j + e, li = lo - pad + k;
        const float v = (lo < Lout && li >= 0 && li < Lin) ? __ldg(xr + li) : 0.f;
        split_bf16(v, hh[e], ll[e]);
      }
      h[j >> 1] = (uint32_t)hh[0] | ((uint32_t)hh[1] << 16);
      l[j >> 1] = (uint32_t)ll[0] | ((uint32_t)ll[1] << 16);
    }
    uint16_t* row = w3 + (long long)rowi * Kdim + col;
    const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row) = hv;
    *reinterpret_cast<uint4*>(row + Pp) = hv;
    *reinterpret_cast<uint4*>(row + 2 * Pp) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
__global__ void __launch_bounds__(256) disc_add_kernel(const float* __restrict__ a, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] += a[i];
}

static inline int ew_blocks(long long n, int cap = 8192) {
  const long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

// bytes of packed operands per batch chunk (B200VOC_WGRAD_TC_CAP_MB overrides it: tests force the chunked path with it)
static long long wgrad_tc_cap() {
  const char* e = getenv("B200VOC_WGRAD_TC_CAP_MB");
  const long long mb = e ? atoll(e) : 0;
  return mb > 0 ? mb << 20 : 1ll << 30;
}
static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }
struct WgTcPlan {
  int Lp, Bc;                  // padded positions per batch item, batch items per chunk
  long long a3_bytes, w3_bytes, bias_bytes, tmp_bytes, total;
};
static bool wgrad_tc_plan(int B, int Cin, int Cout, int Lout, int K, WgTcPlan* pl) {
  pl->Lp = (Lout + 63) & ~63;
  const long long per_b = ((long long)Cin * K + Cout) * 3 * pl->Lp * 2;
  const long long cap = wgrad_tc_cap();
  if (per_b > cap) return false;
  long long bc = cap / per_b;
  pl->Bc = (int)(bc < B ? bc : B);
  const long long Kdim = 3ll * pl->Bc * pl->Lp;
  pl->a3_bytes = align_up((long long)Cout * Kdim * 2, 1024);
  pl->w3_bytes = align_up((long long)Cin * K * Kdim * 2, 1024);
  pl->bias_bytes = align_up(((long long)Cin * K + 128) * 4, 1024);
  pl->tmp_bytes = pl->Bc < B ? align_up((long long)Cout * Cin * K * 4, 1024) : 0;
  pl->total = pl->a3_bytes + pl->w3_bytes + pl->bias_bytes + pl->tmp_bytes;
  return true;
}

int disc_wgrad_tc_launch(const float* x, const float* g, int B, int Cin, int Cout, int Lin, int K, int pad, float* dw,
                         void* workspace, cudaStream_t st) {
  const int Lout = Lin + 2 * pad - K + 1;
  WgTcPlan pl;
  B200_CHECK_ARG(wgrad_tc_plan(B, Cin, Cout, Lout, K, &pl), "disc_conv_wgrad_tc: one batch item exceeds the operand cap");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  uint16_t* a3 = reinterpret_cast<uint16_t*>(ws);
  uint16_t* w3 = reinterpret_cast<uint16_t*>(ws + pl.a3_bytes);
  float* zero_bias = reinterpret_cast<float*>(ws + pl.a3_bytes + pl.w3_bytes);
  float* tmp = reinterpret_cast<float*>(ws + pl.a3_bytes + pl.w3_bytes + pl.bias_bytes);
  B200_CUDA(cudaMemsetAsync(zero_bias, 0, (size_t)pl.bias_bytes, st));
  const long long n = (long long)Cout * Cin * K;
  for (int b0 = 0; b0 < B; b0 += pl.Bc) {
    const int bc = B - b0 < pl.Bc ? B - b0 : pl.Bc;
    const long long Pp = (long long)bc * pl.Lp;
    disc_pack_g3_kernel<<<ew_blocks((long long)Cout * Pp), 256, 0, st>>>(g + (long long)b0 * Cout * Lout, bc, Cout, Lout, pl.Lp,
                                                                          (long long)Cout * Lout, a3);
    B200_CUDA(cudaGetLastError());
    disc_pack_im2col3_kernel<<<ew_blocks((long long)Cin * K * (Pp >> 3), 65536), 256, 0, st>>>(
        x + (long long)b0 * Cin * Lin, bc, Cin, Lin, Lout, pl.Lp, K, pad, (long long)Cin * Lin, w3);
    B200_CUDA(cudaGetLastError());
    float* out = b0 == 0 ? dw : tmp;
    B200_TRY(splitgemm_f32_launch(a3, w3, zero_bias, Cout, Cin * K, 3 * Pp, /*bf16*/ 1, out, st));
    if (b0 != 0) {
      disc_add_kernel<<<ew_blocks(n), 256, 0, st>>>(tmp, n, dw);
      B200_CUDA(cudaGetLastError());
    }
  }
  return B200VOC_OK;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (include/b200voc.h)
extern "C" {

int b200voc_disc_conv_out_len(int Lin, int K, int stride, int pad);
int b200voc_disc_conv_tc_supported(int Cin, int Cout, int K, int stride, int P);
int64_t b200voc_disc_conv_tc_workspace_bytes(int B, int Cin, int L);

static int fill_bwd(b200::DiscBwdParams& p, int B, int Cin, int Cout, int Lin, int P, int K, int stride, int pad,
                    int64_t in_batch_stride, int64_t in_valid, const char* who) {
  B200_CHECK_ARG(B > 0 && Cin > 0 && Cout > 0 && P > 0 && K > 0 && stride > 0 && pad >= 0,
                 "%s: bad shape (B=%d Cin=%d Cout=%d P=%d K=%d stride=%d pad=%d)", who, B, Cin, Cout, P, K, stride, pad);
  const int Lout = b200voc_disc_conv_out_len(Lin, K, stride, pad);
  B200_CHECK_ARG(Lout > 0, "%s: input of %d rows is shorter than the kernel (K=%d, pad=%d)", who, Lin, K, pad);
  B200_CHECK_ARG(B <= 65535, "%s: batch exceeds the grid limits", who);
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Lin = Lin; p.Lout = Lout; p.P = P; p.K = K; p.stride = stride; p.pad = pad;
  p.in_batch_stride = in_batch_stride > 0 ? in_batch_stride : (long long)Cin * Lin * P;
  p.in_valid = in_valid > 0 ? in_valid : (long long)Lin * P;
  return B200VOC_OK;
}

int b200voc_disc_lrelu_bwd(const float* y_pre, const float* gy_pre, const float* gy_act, const float* g_next, float slope,
                           int64_t n, float* g_out, void* stream) {
  B200_CHECK_ARG(g_out && n > 0, "disc_lrelu_bwd: null argument");
  B200_CHECK_ARG(y_pre || !(gy_act || g_next), "disc_lrelu_bwd: activation gradients need the pre-activation map");
  b200::disc_lrelu_bwd_kernel<<<b200::ew_blocks(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y_pre, gy_pre, gy_act, g_next, slope, n, g_out);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int b200voc_disc_bias_grad(const float* g, int B, int Cout, int64_t per_channel, float* db, void* stream) {
  B200_CHECK_ARG(g && db, "disc_bias_grad: null argument");
  B200_CHECK_ARG(B > 0 && Cout > 0 && per_channel > 0, "disc_bias_grad: bad shape");
  b200::disc_bias_grad_kernel<<<Cout, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, B, Cout, per_channel, db);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int b200voc_disc_conv_dgrad(const float* g, const float* w, int B, int Cin, int Cout, int Lin, int P, int K, int stride,
                            int pad, int64_t in_batch_stride, int64_t in_valid, int accumulate, float* dx, void* stream) {
  B200_CHECK_ARG(g && w && dx, "disc_conv_dgrad: null argument");
  b200::DiscBwdParams p{};
  B200_TRY(fill_bwd(p, B, Cin, Cout, Lin, P, K, stride, pad, in_batch_stride, in_valid, "disc_conv_dgrad"));
  B200_CHECK_ARG(b200::ceil_div(Cin, 16) <= 65535, "disc_conv_dgrad: channel count exceeds the grid limits");
  p.g = g; p.w = w; p.dx = dx; p.accumulate = accumulate;
  return b200::disc_dgrad_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

int64_t b200voc_disc_conv_wgrad_scratch_bytes(int B, int Cin, int Cout, int Lin, int P, int K, int stride, int pad) {
  const int Lout = b200voc_disc_conv_out_len(Lin, K, stride, pad);
  if (B <= 0 || Cin <= 0 || Cout <= 0 || P <= 0 || Lout <= 0) return 0;
  const int Z = b200::wgrad_slices(B, Cin, Cout, Lout, P);
  return Z == 1 ? 0 : (int64_t)Z * Cout * Cin * K * 4;
}

int b200voc_disc_conv_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int Lin, int P, int K, int stride,
                            int pad, int64_t in_batch_stride, int64_t in_valid, float* dw, float* scratch, void* stream) {
  B200_CHECK_ARG(x && g && dw, "disc_conv_wgrad: null argument");
  b200::DiscBwdParams p{};
  B200_TRY(fill_bwd(p, B, Cin, Cout, Lin, P, K, stride, pad, in_batch_stride, in_valid, "disc_conv_wgrad"));
  B200_CHECK_ARG(Cin <= 65535 && b200::ceil_div(Cout, b200::kDbCo) <= 65535, "disc_conv_wgrad: channel count exceeds the grid limits");
  p.x = x; p.g = g;
  return b200::disc_wgrad_launch(p, dw, scratch, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_avg_pool1d_k4s2p1_bwd(const float* gy, int64_t rows, int Lin, float* dx, void* stream) {
  B200_CHECK_ARG(gy && dx, "avg_pool1d_bwd: null argument");
  B200_CHECK_ARG(rows > 0 && Lin >= 2, "avg_pool1d_bwd: bad shape (rows=%lld, L=%d)", (long long)rows, Lin);
  const int Lout = (Lin + 2 - 4) / 2 + 1;
  b200::avg_pool_bwd_kernel<<<b200::ew_blocks(rows * Lin), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gy, rows, Lin,
                                                                                                             Lout, dx);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int64_t b200voc_spectral_norm_bwd_scratch_bytes(void) { return (int64_t)b200::kSnParts * 8; }

int b200voc_spectral_norm_bwd(const float* dw, const float* w, const float* u, const float* v, const float* sigma, int rows,
                              int cols, float* dw_orig, void* scratch, void* stream) {
  B200_CHECK_ARG(dw && w && u && v && sigma && dw_orig && scratch, "spectral_norm_bwd: null argument");
  B200_CHECK_ARG(rows > 0 && cols > 0, "spectral_norm_bwd: bad shape (%d x %d)", rows, cols);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "spectral_norm_bwd: scratch must be 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n = (long long)rows * cols;
  const int parts = b200::ew_blocks(n, b200::kSnParts);
  double* pp = reinterpret_cast<double*>(scratch);
  b200::sn_bwd_dot_kernel<<<parts, 256, 0, st>>>(dw, w, n, pp);
  B200_CUDA(cudaGetLastError());
  b200::sn_bwd_apply_kernel<<<b200::ew_blocks(n, 4096), 256, 0, st>>>(dw, u, v, sigma, pp, parts, rows, cols, dw_orig);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

// ---- tensor-core paths ------------------------------------------------------------------------------------------
int b200voc_disc_conv_dgrad_tc_supported(int Cin, int Cout, int K, int stride, int P, int pad) {
  // the dgrad is Conv1d(Cout -> Cin, K, pad K - 1 - pad) with flipped weights
  return b200voc_disc_conv_tc_supported(Cout, Cin, K, stride, P) && pad >= 0 && pad < K;
}

int b200voc_disc_flip_weight(const float* w, int Cout, int Cin, int K, float* wt, void* stream) {
  B200_CHECK_ARG(w && wt, "disc_flip_weight: null argument");
  B200_CHECK_ARG(Cout > 0 && Cin > 0 && K > 0, "disc_flip_weight: bad shape (%d, %d, %d)", Cout, Cin, K);
  const long long n = (long long)Cout * Cin * K;
  b200::disc_flip_weight_kernel<<<b200::ew_blocks(n, 4096), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, Cout, Cin, K, wt);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

int b200voc_disc_conv_wgrad_tc_supported(int B, int Cin, int Cout, int Lin, int K, int stride, int P, int pad) {
  if (stride != 1 || P != 1 || B <= 0 || Cin < 64 || Cout < 64 || K < 1 || pad < 0 || (Cin * K) % 4 != 0) return 0;
  const int Lout = Lin + 2 * pad - K + 1;
  if (Lout <= 0) return 0;
  b200::WgTcPlan pl;
  return b200::wgrad_tc_plan(B, Cin, Cout, Lout, K, &pl) ? 1 : 0;
}

int64_t b200voc_disc_conv_wgrad_tc_workspace_bytes(int B, int Cin, int Cout, int Lin, int K, int pad) {
  const int Lout = Lin + 2 * pad - K + 1;
  b200::WgTcPlan pl;
  if (Lout <= 0 || !b200::wgrad_tc_plan(B, Cin, Cout, Lout, K, &pl)) return 0;
  return pl.total;
}

int b200voc_disc_conv_wgrad_tc(const float* x, const float* g, int B, int Cin, int Cout, int Lin, int K, int pad, float* dw,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  B200_CHECK_ARG(x && g && dw && workspace, "disc_conv_wgrad_tc: null argument");
  B200_CHECK_ARG(b200voc_disc_conv_wgrad_tc_supported(B, Cin, Cout, Lin, K, 1, 1, pad),
                 "disc_conv_wgrad_tc: unsupported shape (B=%d Cin=%d Cout=%d L=%d K=%d pad=%d)", B, Cin, Cout, Lin, K, pad);
  B200_CHECK_ARG(workspace_bytes >= b200voc_disc_conv_wgrad_tc_workspace_bytes(B, Cin, Cout, Lin, K, pad),
                 "disc_conv_wgrad_tc: workspace too small");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "disc_conv_wgrad_tc: workspace must be 1024-byte aligned");
  return b200::disc_wgrad_tc_launch(x, g, B, Cin, Cout, Lin, K, pad, dw, workspace, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
