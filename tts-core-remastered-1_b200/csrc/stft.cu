// K7/K8/K9: STFT family in fp32 (replaces vocoder7/stft.py:9-54 and the torchaudio
// MelSpectrogram call sites).  Shared-memory Stockham FFTs (fft_core.cuh): N/8 threads per frame,
// 8 frames per CTA, every input sample is read from HBM once per CTA (frames overlap 50-87 %),
// the complex spectrogram is never materialised when the consumer is |X|*gain, mel or log-mel.
//
//   stft:  frame = reflect-padded wav[f*hop - n/2 ...] * periodic Hann  ->  rFFT (packed complex
//          FFT of n/2 points + split)  ->  |X|*gain  |  re,im  |  |X|^2 -> sparse HTK mel -> log
//   istft: spectrum block -> inverse packed FFT -> * window, gather overlap-add, / sum w^2
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <cmath>

#include "common.cuh"
#include "fft_core.cuh"
#include "fft_warp.cuh"

namespace b200 {

using namespace fft;

constexpr int kFPB = 8;   // frames (thread groups) per CTA per round

enum StftMode { MODE_MAG = 0, MODE_COMPLEX = 1, MODE_MEL = 2, MODE_L1 = 3 };

struct MelTable {      // CSR by mel bin over frequency bins
  const int* lo;       // [n_mels] first bin with non-zero weight
  const int* cnt;      // [n_mels]
  const int* off;      // [n_mels] offset into w
  const float* w;
  int n_mels;
  int nnz;             // number of weights (length of w)
  // the same filters as 4-aligned, zero-padded runs (stft1024's mel stage reads bins and weights as float4):
  // filter m covers bins lo4[m] .. lo4[m] + 4*n4[m] - 1 with weights w4[off4[m] ...] (off4 a multiple of 4)
  const int* lo4;
  const int* n4;
  const int* off4;
  const float* w4;
  int nnz4;
};

struct FftTables {     // per (device, n_fft), built once on the host in double precision
  const float* win;    // [n_fft] periodic Hann
  const float2* tw;    // [<= n_fft/2] per-pass Stockham twiddle tables (fft_core.cuh)
  const float2* tw2;   // [n_fft/2+1] exp(-pi i k / (n_fft/2))
  const float2* tw1;   // [16][32] exp(-2 pi i lane k1 / 512): stage-1 twiddles of the warp-per-frame FFT (n_fft = 1024)
};

static int get_fft_tables(int n_fft, FftTables* out);

__device__ __forceinline__ void group_sync(int g, int threads) {
  if (threads == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(threads) : "memory");
}

struct StftParams {
  FftTables tab;
  const float* wav;    // [B, Nsamp]
  const float* wav2;   // MODE_L1: second waveform
  int B, Nsamp, hop, frames;
  const float* gain;   // [bins] or null
  float* out;
  double* out_sum;     // MODE_L1
  MelTable mel;
  int log_compress;
};

__device__ __forceinline__ int reflect(int s, int n) {
  if (s < 0) s = -s;
  if (s >= n) s = 2 * (n - 1) - s;
  return min(max(s, 0), n - 1);   // frames past the end of the signal (tail CTA) read clamped garbage, never stored
}

template <int NFFT, int MODE>
__global__ void __launch_bounds__(kFPB * (NFFT / 16))
stft_kernel(const StftParams p) {
  constexpr int N = NFFT / 2, T = N / 8, BINS = N + 1, NPAD = N + N / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int span_len = (kFPB - 1) * p.hop + NFFT;
  float* span = reinterpret_cast<float*>(smem_raw);
  float* win = span + ((span_len + 3) & ~3);
  float2* tw = reinterpret_cast<float2*>(win + NFFT);
  float2* tw2 = tw + N;
  float2* buf = tw2 + (N + 2);
  float* stage = reinterpret_cast<float*>(buf + kFPB * NPAD);   // [BINS][kFPB+1] (float2 for MODE_COMPLEX)

  const int g = threadIdx.x / T, t = threadIdx.x % T;
  const int b = blockIdx.y, f0 = blockIdx.x * kFPB;
  float2* mybuf = buf + g * NPAD;

  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) win[i] = __ldg(p.tab.win + i);
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = __ldg(p.tab.tw + i);
  for (int i = threadIdx.x; i <= N; i += blockDim.x) tw2[i] = __ldg(p.tab.tw2 + i);
  float l1_acc = 0.f;
  constexpr int NPASS = MODE == MODE_L1 ? 2 : 1;
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    const float* w = (pass == 0 ? p.wav : p.wav2) + (long long)b * p.Nsamp;
    __syncthreads();
    for (int i = threadIdx.x; i < span_len; i += blockDim.x)
      span[i] = __ldg(w + reflect(f0 * p.hop - N + i, p.Nsamp));
    __syncthreads();
    const float* fr = span + g * p.hop;
    auto load0 = [&](int m) { return make_float2(fr[2 * m] * win[2 * m], fr[2 * m + 1] * win[2 * m + 1]); };
    transform<N, false>(t, load0, mybuf, tw, [g] { group_sync(g, T); });
    // split -> bins k = t, t+T, ... and k = N
    for (int k = t; k <= N; k += T) {
      const float2 X = rfft_bin(mybuf, tw2, N, k);
      if (MODE == MODE_COMPLEX) {
        reinterpret_cast<float2*>(stage)[k * (kFPB + 1) + g] = X;
      } else if (MODE == MODE_MEL) {
        stage[k * (kFPB + 1) + g] = X.x * X.x + X.y * X.y;
      } else {
        float mag = sqrtf(X.x * X.x + X.y * X.y);
        if (p.gain) mag *= __ldg(p.gain + k);
        if (MODE == MODE_L1) {
          if (pass == 0) stage[k * (kFPB + 1) + g] = mag;
          else if (f0 + g < p.frames) l1_acc += fabsf(stage[k * (kFPB + 1) + g] - mag);
        } else {
          stage[k * (kFPB + 1) + g] = mag;
        }
      }
    }
  }
  __syncthreads();
  if (MODE == MODE_MAG) {
    float* o = p.out + (long long)b * BINS * p.frames;
    for (int i = threadIdx.x; i < BINS * kFPB; i += blockDim.x) {
      const int k = i / kFPB, f = i % kFPB;
      if (f0 + f < p.frames) o[(long long)k * p.frames + f0 + f] = stage[k * (kFPB + 1) + f];
    }
  } else if (MODE == MODE_COMPLEX) {
    float2* o = reinterpret_cast<float2*>(p.out) + (long long)b * BINS * p.frames;
    for (int i = threadIdx.x; i < BINS * kFPB; i += blockDim.x) {
      const int k = i / kFPB, f = i % kFPB;
      if (f0 + f < p.frames) o[(long long)k * p.frames + f0 + f] = reinterpret_cast<float2*>(stage)[k * (kFPB + 1) + f];
    }
  } else if (MODE == MODE_MEL) {
    float* o = p.out + (long long)b * p.mel.n_mels * p.frames;
    for (int i = threadIdx.x; i < p.mel.n_mels * kFPB; i += blockDim.x) {
      const int m = i / kFPB, f = i % kFPB;
      const int lo = __ldg(p.mel.lo + m), cnt = __ldg(p.mel.cnt + m);
      const float* wv = p.mel.w + __ldg(p.mel.off + m);
      float acc = 0.f;
      for (int k = 0; k < cnt; ++k) acc = fmaf(__ldg(wv + k), stage[(lo + k) * (kFPB + 1) + f], acc);
      if (p.log_compress) acc = logf(fmaxf(acc, 1e-5f));
      if (f0 + f < p.frames) o[(long long)m * p.frames + f0 + f] = acc;
    }
  } else {
    // block reduction of the L1 partial sum -> one double atomic per CTA
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) l1_acc += __shfl_xor_sync(0xffffffffu, l1_acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l1_acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) atomicAdd(p.out_sum, (double)v);
    }
  }
}

// ------------------------------------------------------------------ fast path, n_fft = 1024
// One warp per frame (fft_warp.cuh), 8 warps = 8 consecutive frames per round, kRounds rounds per
// CTA.  Window samples and the per-lane twiddles live in registers for the whole CTA.
constexpr int kFastWarps = 8;
// rounds (of 8 frames) per CTA: the fused log-mel gains from amortising the table set-up and the one cold span load over 8
// rounds (0.766 -> 0.744 ms), the store-heavy magnitude / complex modes are better with more, shorter CTAs (measured)
template <int MODE> struct FastRounds { static constexpr int value = MODE == 2 ? 8 : 4; };

// REGTAB (opt-in, B200VOC_STFT_REGTAB=1): the 16 window pairs and 15 stage-1 twiddles a lane needs are the same for
// every frame, so they are read from shared memory once per CTA and kept in registers -- 31 of the ~119 64-bit
// shared-memory accesses per frame (the LSU data pipe is this kernel's busiest unit) for ~60
// registers, i.e. two instead of three resident CTAs per SM.
template <int MODE, bool REGTAB>
__global__ void __launch_bounds__(kFastWarps * 32, REGTAB ? 2 : 3)
stft1024_kernel(const StftParams p) {
  constexpr int NFFT = 1024, N = 512, BINS = 513, SW = kFastWarps + 1;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int span_len = (kFastWarps - 1) * p.hop + NFFT;
  // The staged outputs [BINS][SW] share their storage with the round's samples (dead once every warp holds its frame
  // in registers; one more __syncthreads per round): 79 -> 68 KB per CTA, i.e. three resident CTAs per SM instead of
  // two.  MODE_L1 reads a second span while the first pass's outputs are staged, so it keeps both.
  constexpr bool ALIAS = MODE != MODE_L1;
  constexpr int STAGE_BYTES = BINS * SW * (MODE == MODE_COMPLEX ? 8 : 4);
  const int span_bytes = ((span_len + 3) & ~3) * 4;
  const int head_bytes = ALIAS ? (((span_bytes > STAGE_BYTES ? span_bytes : STAGE_BYTES) + 15) & ~15) : span_bytes;
  float* span = reinterpret_cast<float*>(smem_raw);                              // one round's samples
  float2* tw2 = reinterpret_cast<float2*>(smem_raw + head_bytes);                // [N + 1] (padded to N + 2)
  float2* win2 = tw2 + (N + 2);                                                  // [N] window pairs (w[2n], w[2n+1])
  float2* tw1 = win2 + N;                                                        // [16][32] W_512^(lane*k1)
  float2* wbuf = tw1 + 512;                                                      // [warps][576]
  float* stage = ALIAS ? reinterpret_cast<float*>(smem_raw) : reinterpret_cast<float*>(wbuf + kFastWarps * 576);   // [BINS][SW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  float2* T = wbuf + warp * 576;       // 16 x 33 transpose scratch, later Z[512] linear (padded)

  for (int i = threadIdx.x; i <= N; i += blockDim.x) tw2[i] = __ldg(p.tab.tw2 + i);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    win2[i] = __ldg(reinterpret_cast<const float2*>(p.tab.win) + i);
    tw1[i] = __ldg(p.tab.tw1 + i);
  }
  // MODE_MEL: the sparse filterbank's weights live in shared memory (4 KB for 80 HTK mels): read through __ldg they
  // were 130 of the ~550 load/store-pipe wavefronts per frame of a kernel whose LSU pipe is 76 % busy
  // The mel stage reads FOUR bins and four weights per load (float4): the staged power spectrum is frame-major there
  // ([8 frames][516]: 516 = 4 mod 32, so the 8 frames of a quarter warp cover all 32 banks) and the filters are
  // 4-aligned zero-padded runs -- with one 4-byte load per bin and per weight the stage was 30 % of the kernel's
  // shared-memory wavefronts (2.2 per bin load: the four filters a warp evaluates conflicted).
  constexpr int kMelSmemMax = 2560, MSW = 516;
  float* sMelW = reinterpret_cast<float*>(wbuf + kFastWarps * 576);
  const bool mel_smem = MODE == MODE_MEL && p.mel.nnz4 <= kMelSmemMax;
  if (mel_smem)
    for (int i = threadIdx.x; i < p.mel.nnz4; i += blockDim.x) sMelW[i] = __ldg(p.mel.w4 + i);
  float l1_acc = 0.f;
  constexpr int NPASS = MODE == MODE_L1 ? 2 : 1;
  float2 wreg[REGTAB ? 16 : 1], treg[REGTAB ? 16 : 1];
  if constexpr (REGTAB) {
    __syncthreads();                                     // the tables above are complete
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      wreg[n1] = win2[n1 * 32 + lane];
      treg[n1] = tw1[n1 * 32 + lane];
    }
  }
  // The NEXT round's samples are requested into registers (up to PF float4 per thread: hop <= 292) as soon as this
  // round's frames sit in registers, and stored to shared memory at the top of the next round: the span fill was a
  // phase in which every warp of the CTA waited for HBM (long_scoreboard 18 % of the stall samples).
  constexpr int PF = 3;
  float4 pf[PF];
  bool have_pf = false;
  auto interior = [&](int f0r) {
    const int s0r = f0r * p.hop - N;
    return s0r >= 0 && s0r + span_len <= p.Nsamp && ((s0r | p.Nsamp | span_len) & 3) == 0;   // (hop % 4 == 2: span_len % 4 == 2)
  };
  constexpr int kFastRounds = FastRounds<MODE>::value;
#pragma unroll 1
  for (int round = 0; round < kFastRounds; ++round) {
    const int f0 = (blockIdx.x * kFastRounds + round) * kFastWarps;
    if (f0 >= p.frames) break;
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
      const float* w = (pass == 0 ? p.wav : p.wav2) + (long long)b * p.Nsamp;
      __syncthreads();                                   // previous users of span / stage are done
      const int s0 = f0 * p.hop - N;                     // first sample of the span (may be < 0: reflect)
      if (have_pf) {
#pragma unroll
        for (int t = 0; t < PF; ++t) {
          const int i = threadIdx.x + t * (kFastWarps * 32);
          if (i < (span_len >> 2)) reinterpret_cast<float4*>(span)[i] = pf[t];
        }
        have_pf = false;
      } else if (interior(f0)) {
        const float4* src = reinterpret_cast<const float4*>(w + s0);     // interior: plain vector loads
        for (int i = threadIdx.x; i < (span_len >> 2); i += blockDim.x) reinterpret_cast<float4*>(span)[i] = __ldg(src + i);
      } else {
        for (int i = threadIdx.x; i < span_len; i += blockDim.x) span[i] = __ldg(w + reflect(s0 + i, p.Nsamp));
      }
      __syncthreads();
      const float2* fr = reinterpret_cast<const float2*>(span + warp * p.hop);   // hop is even (checked on the host)
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const float2 x = fr[n1 * 32 + lane];
        float2 wn;
        if constexpr (REGTAB) wn = wreg[n1];
        else wn = win2[n1 * 32 + lane];
        v[n1] = make_float2(x.x * wn.x, x.y * wn.y);
      }
      if constexpr (ALIAS) __syncthreads();              // every warp has its frame: the span's storage becomes the stage
      if (NPASS == 1 && round + 1 < kFastRounds && (span_len >> 2) <= PF * kFastWarps * 32) {
        const int f0n = f0 + kFastWarps;
        if (f0n < p.frames && interior(f0n)) {
          const float4* src = reinterpret_cast<const float4*>(w + f0n * p.hop - N);
#pragma unroll
          for (int t = 0; t < PF; ++t) {
            const int i = threadIdx.x + t * (kFastWarps * 32);
            if (i < (span_len >> 2)) pf[t] = __ldg(src + i);
          }
          have_pf = true;
        }
      }
      if constexpr (REGTAB) warp_fft512_regtw<false>(v, T, treg, lane);
      else warp_fft512<false>(v, T, tw1, lane);
      // Z -> linear per-warp buffer (reusing the transpose scratch), then the real-FFT split
      {
        // linear (unpadded) Z: a half-warp stores / loads 16 consecutive float2 = all 32 banks once; the k + k/8 padding
        // of the Stockham path made lane 15 collide with lane 0 (2 x the wavefronts of 48 accesses per frame, ncu)
        float2* zp = T + (lane & 15) + 256 * (lane >> 4);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) zp[16 * k2] = v[k2];
      }
      __syncwarp();
      // real-FFT split, two bins per pair of loads: with a = (Z[k] + conj Z[N-k])/2, d = (Z[k] - conj Z[N-k])/2,
      // t = i * exp(-i pi k / N) * d:   X[k] = a - t,   X[N-k] = conj(a + t)
      auto emit = [&](int k, float2 X) {
        if (MODE == MODE_COMPLEX) {
          reinterpret_cast<float2*>(stage)[k * SW + warp] = X;
        } else if (MODE == MODE_MEL) {
          stage[warp * MSW + k] = X.x * X.x + X.y * X.y;       // frame-major (see the mel stage)
        } else {
          float mag = sqrtf(X.x * X.x + X.y * X.y);
          if (p.gain) mag *= __ldg(p.gain + k);
          if (MODE == MODE_L1) {
            if (pass == 0) stage[k * SW + warp] = mag;
            else if (f0 + warp < p.frames) l1_acc += fabsf(stage[k * SW + warp] - mag);
          } else {
            stage[k * SW + warp] = mag;
          }
        }
      };
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;                       // 0 .. 255, partner bin N - k
        const float2 zk = T[k], zn = T[(N - k) & (N - 1)], wk = tw2[k];
        const float2 a = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));
        const float2 wd = cmul(wk, d);
        const float2 t = make_float2(-wd.y, wd.x);        // i * wd
        emit(k, make_float2(a.x - t.x, a.y - t.y));
        emit(N - k, make_float2(a.x + t.x, -(a.y + t.y)));
      }
      if (lane == 0) {
        const float2 zm = T[N / 2];
        emit(N / 2, make_float2(zm.x, -zm.y));
      }
      if (MODE == MODE_MEL && lane >= 1 && lane <= 3) stage[warp * MSW + BINS - 1 + lane] = 0.f;   // zero-weight padding bins 513..515
      __syncwarp();
    }
    __syncthreads();
    // cooperative write-out: thread = (frame f = tid & 7, bin k = tid >> 3 + 32*i): 8 frames of one bin are
    // one 32-byte sector of out[b, k, f0 .. f0+7]
    const int f = threadIdx.x & (kFastWarps - 1), kk = threadIdx.x >> 3;
    const bool fok = f0 + f < p.frames;
    if (MODE == MODE_MAG) {
      float* o = p.out + ((long long)b * BINS + kk) * p.frames + f0 + f;
      const float* sp = stage + kk * SW + f;
      for (int k = kk; k < BINS; k += 32, o += 32ll * p.frames, sp += 32 * SW)
        if (fok) *o = *sp;
    } else if (MODE == MODE_COMPLEX) {
      float2* o = reinterpret_cast<float2*>(p.out) + ((long long)b * BINS + kk) * p.frames + f0 + f;
      const float2* sp = reinterpret_cast<const float2*>(stage) + kk * SW + f;
      for (int k = kk; k < BINS; k += 32, o += 32ll * p.frames, sp += 32 * SW)
        if (fok) *o = *sp;
    } else if (MODE == MODE_MEL) {
      float* o = p.out + (long long)b * p.mel.n_mels * p.frames + f0 + f;
      for (int m = kk; m < p.mel.n_mels; m += 32) {
        const int lo4 = __ldg(p.mel.lo4 + m), n4 = __ldg(p.mel.n4 + m), woff = __ldg(p.mel.off4 + m);
        const float4* sp = reinterpret_cast<const float4*>(stage + f * MSW + lo4);
        const float4* wv = reinterpret_cast<const float4*>((mel_smem ? sMelW : p.mel.w4) + woff);
        float acc = 0.f;
        for (int i = 0; i < n4; ++i) {
          const float4 w4 = mel_smem ? wv[i] : __ldg(wv + i);
          const float4 p4 = sp[i];
          acc = fmaf(w4.x, p4.x, acc);                    // bins in ascending order, as the 4-byte loop summed them
          acc = fmaf(w4.y, p4.y, acc);
          acc = fmaf(w4.z, p4.z, acc);
          acc = fmaf(w4.w, p4.w, acc);
        }
        if (p.log_compress) acc = logf(fmaxf(acc, 1e-5f));
        if (fok) o[(long long)m * p.frames] = acc;
      }
    }
  }
  if (MODE == MODE_L1) {
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) l1_acc += __shfl_xor_sync(0xffffffffu, l1_acc, o);
    if (lane == 0) red[warp] = l1_acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float vv = threadIdx.x < kFastWarps ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, o);
      if (threadIdx.x == 0) atomicAdd(p.out_sum, (double)vv);
    }
  }
}

template <int MODE>
static int launch_stft1024(const StftParams& p, cudaStream_t st) {
  const int span_len = (kFastWarps - 1) * p.hop + 1024;
  const size_t span_bytes = (size_t)((span_len + 3) & ~3) * 4, stage_bytes = (size_t)513 * (kFastWarps + 1) * (MODE == MODE_COMPLEX ? 8 : 4);
  const size_t head = MODE != MODE_L1 ? (((span_bytes > stage_bytes ? span_bytes : stage_bytes) + 15) & ~(size_t)15) : span_bytes + stage_bytes;
  const size_t mel_bytes = (MODE == MODE_MEL && p.mel.nnz4 <= 2560) ? (size_t)p.mel.nnz4 * 4 : 0;
  const size_t smem = head + 514 * 8 + 512 * 8 + 512 * 8 + (size_t)kFastWarps * 576 * 8 + mel_bytes + 64;
  B200_CHECK_ARG(smem <= 227 * 1024, "stft: hop %d needs %zu bytes of shared memory", p.hop, smem);
  dim3 grid(ceil_div(p.frames, kFastWarps * FastRounds<MODE>::value), p.B);
  static const bool regtab = [] { const char* e = getenv("B200VOC_STFT_REGTAB"); return e && e[0] == '1'; }();
  if (regtab) {   // opt-in A/B variant (written after the round's GPU budget was spent: not yet measured)
    B200_CUDA(cudaFuncSetAttribute(stft1024_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft1024_kernel<MODE, true><<<grid, kFastWarps * 32, smem, st>>>(p);
  } else {
    B200_CUDA(cudaFuncSetAttribute(stft1024_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft1024_kernel<MODE, false><<<grid, kFastWarps * 32, smem, st>>>(p);
  }
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int NFFT>
static size_t stft_smem(int hop, int mode) {
  constexpr int N = NFFT / 2, NPAD = N + N / 8, BINS = N + 1;
  const int span_len = (kFPB - 1) * hop + NFFT;
  size_t s = (size_t)((span_len + 3) & ~3) * 4 + NFFT * 4 + (size_t)N * 8 + (size_t)(N + 2) * 8 + (size_t)kFPB * NPAD * 8;
  s += (size_t)BINS * (kFPB + 1) * (mode == MODE_COMPLEX ? 8 : 4);
  return s + 64;
}

template <int NFFT, int MODE>
static int launch_stft(const StftParams& p, cudaStream_t st) {
  const size_t smem = stft_smem<NFFT>(p.hop, MODE);
  B200_CHECK_ARG(smem <= 227 * 1024, "stft: hop %d needs %zu bytes of shared memory", p.hop, smem);
  B200_CUDA(cudaFuncSetAttribute(stft_kernel<NFFT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(p.frames, kFPB), p.B);
  stft_kernel<NFFT, MODE><<<grid, kFPB * (NFFT / 16), smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int MODE>
static int dispatch_stft(int n_fft, StftParams p, cudaStream_t st) {
  if (n_fft != 512 && n_fft != 1024 && n_fft != 2048) {
    set_error("stft: n_fft=%d unsupported (512/1024/2048, vocoder7/config.py:39 stft_sizes)", n_fft);
    return B200VOC_ERR_UNSUPPORTED;
  }
  B200_TRY(get_fft_tables(n_fft, &p.tab));
  switch (n_fft) {
    case 512: return launch_stft<512, MODE>(p, st);
    case 1024:
      if (p.hop % 2 == 0 && !getenv("B200VOC_STFT_V1")) return launch_stft1024<MODE>(p, st);
      return launch_stft<1024, MODE>(p, st);
    case 2048: return launch_stft<2048, MODE>(p, st);
  }
  set_error("stft: n_fft=%d unsupported (512/1024/2048, vocoder7/config.py:39 stft_sizes)", n_fft);
  return B200VOC_ERR_UNSUPPORTED;
}

static int check_stft_args(const float* wav, int B, int N, int n_fft, int hop) {
  B200_CHECK_ARG(wav != nullptr, "stft: null waveform");
  B200_CHECK_ARG(B > 0 && N > 0, "stft: empty input (B=%d, N=%d)", B, N);
  B200_CHECK_ARG(hop > 0 && hop <= n_fft, "stft: hop %d out of range", hop);
  B200_CHECK_ARG(N > n_fft / 2, "stft: reflect padding needs N=%d > n_fft/2=%d (same as torch.stft)", N, n_fft / 2);
  return B200VOC_OK;
}

// ------------------------------------------------------------------ FFT tables (host, cached)
static std::mutex g_tab_mu;
static std::map<std::pair<int, int>, FftTables> g_tab_cache;

static int get_fft_tables(int n_fft, FftTables* out) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_pair(dev, n_fft);
  auto it = g_tab_cache.find(key);
  if (it == g_tab_cache.end()) {
    const int N = n_fft / 2;
    std::vector<float> win(n_fft);
    std::vector<float2> tw(N), tw2(N + 1);
    for (int i = 0; i < n_fft; ++i) win[i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / n_fft));
    if (N == 256) fill_pass_twiddles<256>(tw.data());
    else if (N == 512) fill_pass_twiddles<512>(tw.data());
    else fill_pass_twiddles<1024>(tw.data());   // per-pass tables, conflict-free layout (fft_core.cuh)
    for (int i = 0; i <= N; ++i) tw2[i] = make_float2((float)std::cos(-M_PI * i / N), (float)std::sin(-M_PI * i / N));
    float* dwin;
    float2 *dtw, *dtw2;
    B200_CUDA(cudaMalloc(&dwin, n_fft * sizeof(float)));
    B200_CUDA(cudaMalloc(&dtw, N * sizeof(float2)));
    B200_CUDA(cudaMalloc(&dtw2, (N + 1) * sizeof(float2)));
    B200_CUDA(cudaMemcpy(dwin, win.data(), n_fft * sizeof(float), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(dtw, tw.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(dtw2, tw2.data(), (N + 1) * sizeof(float2), cudaMemcpyHostToDevice));
    // stage-1 twiddles of the warp FFT (fft_warp.cuh): the kernels used to evaluate 512 sincospif per CTA for them
    std::vector<float2> tw1(512);
    for (int i = 0; i < 512; ++i) {
      const int k1 = i >> 5, tl = i & 31;
      const double a = -2.0 * M_PI * (double)((tl * k1) & 511) / 512.0;
      tw1[i] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    float2* dtw1;
    B200_CUDA(cudaMalloc(&dtw1, 512 * sizeof(float2)));
    B200_CUDA(cudaMemcpy(dtw1, tw1.data(), 512 * sizeof(float2), cudaMemcpyHostToDevice));
    FftTables t{};
    t.win = dwin; t.tw = dtw; t.tw2 = dtw2; t.tw1 = dtw1;
    it = g_tab_cache.emplace(key, t).first;
  }
  *out = it->second;
  return B200VOC_OK;
}

// ------------------------------------------------------------------ mel filterbank (host, cached)
// HTK mel scale, triangular, no area normalisation: the matrix torchaudio's MelSpectrogram builds
// (reference_encoder/utils.py:31-36 call site).  Stored sparse by mel bin.
struct MelDev {
  int *lo, *cnt, *off;
  float* w;
  int nnz;
  int *lo4, *n4, *off4;
  float* w4;
  int nnz4;
};
static std::mutex g_mel_mu;
static std::map<std::tuple<int, int, int, int>, MelDev> g_mel_cache;

static int get_mel_table(int n_fft, int n_mels, int sr, MelTable* out) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mel_mu);
  auto key = std::make_tuple(dev, n_fft, n_mels, sr);
  auto it = g_mel_cache.find(key);
  if (it == g_mel_cache.end()) {
    const int bins = n_fft / 2 + 1;
    const double f_max = (double)(sr / 2);
    auto hz2mel = [](double f) { return 2595.0 * std::log10(1.0 + f / 700.0); };
    auto mel2hz = [](double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); };
    std::vector<double> fpts(n_mels + 2);
    const double m_lo = hz2mel(0.0), m_hi = hz2mel(f_max);
    for (int i = 0; i < n_mels + 2; ++i) fpts[i] = mel2hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<int> lo(n_mels), cnt(n_mels), off(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
      int first = -1, last = -1;
      std::vector<float> col(bins);
      for (int k = 0; k < bins; ++k) {
        const double f = (double)(sr / 2) * k / (bins - 1);
        const double down = (f - fpts[m]) / (fpts[m + 1] - fpts[m]);
        const double up = (fpts[m + 2] - f) / (fpts[m + 2] - fpts[m + 1]);
        const double v = std::fmax(0.0, std::fmin(down, up));
        col[k] = (float)v;
        if (v > 0.0) {
          if (first < 0) first = k;
          last = k;
        }
      }
      lo[m] = first < 0 ? 0 : first;
      cnt[m] = first < 0 ? 0 : last - first + 1;
      off[m] = (int)w.size();
      for (int k = 0; k < cnt[m]; ++k) w.push_back(col[lo[m] + k]);
    }
    if (w.empty()) w.push_back(0.f);
    MelDev d{};
    B200_CUDA(cudaMalloc(&d.lo, n_mels * sizeof(int)));
    B200_CUDA(cudaMalloc(&d.cnt, n_mels * sizeof(int)));
    B200_CUDA(cudaMalloc(&d.off, n_mels * sizeof(int)));
    B200_CUDA(cudaMalloc(&d.w, w.size() * sizeof(float)));
    B200_CUDA(cudaMemcpy(d.lo, lo.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(d.cnt, cnt.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(d.off, off.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(d.w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    d.nnz = (int)w.size();
    {
      std::vector<int> lo4(n_mels), n4(n_mels), off4(n_mels);
      std::vector<float> w4;
      for (int m = 0; m < n_mels; ++m) {
        lo4[m] = lo[m] & ~3;
        const int end = lo[m] + cnt[m];
        n4[m] = cnt[m] > 0 ? (end - lo4[m] + 3) / 4 : 0;
        off4[m] = (int)w4.size();
        for (int i = 0; i < 4 * n4[m]; ++i) {
          const int k = lo4[m] + i;
          w4.push_back(k >= lo[m] && k < end ? w[off[m] + (k - lo[m])] : 0.f);
        }
      }
      if (w4.empty()) w4.resize(4, 0.f);
      B200_CUDA(cudaMalloc(&d.lo4, n_mels * sizeof(int)));
      B200_CUDA(cudaMalloc(&d.n4, n_mels * sizeof(int)));
      B200_CUDA(cudaMalloc(&d.off4, n_mels * sizeof(int)));
      B200_CUDA(cudaMalloc(&d.w4, w4.size() * sizeof(float)));
      B200_CUDA(cudaMemcpy(d.lo4, lo4.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
      B200_CUDA(cudaMemcpy(d.n4, n4.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
      B200_CUDA(cudaMemcpy(d.off4, off4.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
      B200_CUDA(cudaMemcpy(d.w4, w4.data(), w4.size() * sizeof(float), cudaMemcpyHostToDevice));
      d.nnz4 = (int)w4.size();
    }
    it = g_mel_cache.emplace(key, d).first;
  }
  out->lo = it->second.lo;
  out->cnt = it->second.cnt;
  out->off = it->second.off;
  out->w = it->second.w;
  out->n_mels = n_mels;
  out->nnz = it->second.nnz;
  out->lo4 = it->second.lo4;
  out->n4 = it->second.n4;
  out->off4 = it->second.off4;
  out->w4 = it->second.w4;
  out->nnz4 = it->second.nnz4;
  return B200VOC_OK;
}

// ------------------------------------------------------------------ K9: iSTFT
// One CTA produces OPB = (2*kFPB - n/hop + 1) * hop output samples from the 2*kFPB frames that
// touch them: the spectrum block is loaded coalesced (frames are the contiguous axis), each frame
// group runs the inverse packed FFT, and every output sample gathers its n/hop windowed
// contributions and the window-envelope (sum w^2) -- deterministic, no atomics.
struct IstftParams {
  FftTables tab;
  const float2* spec;   // [B, bins, frames]
  float* wav;           // [B, Nout]
  int B, frames, hop, Nout;
  int joff;             // padded-signal coordinate of output sample 0: n_fft/2 (iSTFT: centre trimmed) or 0 (adjoint)
  int normalize;        // 1: divide by the window envelope (torch.istft); 0: plain overlap-add (adjoint of the STFT)
};

template <int NFFT>
__global__ void __launch_bounds__(kFPB * (NFFT / 16))
istft_kernel(const IstftParams p) {
  constexpr int N = NFFT / 2, T = N / 8, NPAD = N + N / 8, SLOTS = 2 * kFPB;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* win = reinterpret_cast<float*>(smem_raw);
  float2* tw = reinterpret_cast<float2*>(win + NFFT);
  float2* tw2 = tw + N;
  float2* xN = tw2 + (N + 2);               // [SLOTS] Nyquist bins
  float2* buf = xN + SLOTS;                 // [SLOTS][NPAD]
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  const int b = blockIdx.y;
  const int ratio = NFFT / p.hop;
  const int opb = (SLOTS - ratio + 1) * p.hop;
  const int n0 = blockIdx.x * opb;                       // first output sample of this CTA
  const int f_lo = (n0 + p.joff - NFFT) / p.hop + 1;     // may be negative (the numerator is a multiple of hop)
  const int bins = N + 1;

  for (int i = threadIdx.x; i < NFFT; i += blockDim.x) win[i] = __ldg(p.tab.win + i);
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = __ldg(p.tab.tw + i);
  for (int i = threadIdx.x; i <= N; i += blockDim.x) tw2[i] = __ldg(p.tab.tw2 + i);
  // coalesced load of the spectrum block: consecutive threads -> consecutive frames of one bin
  const float2* sp = p.spec + (long long)b * bins * p.frames;
  for (int i = threadIdx.x; i < bins * SLOTS; i += blockDim.x) {
    const int k = i / SLOTS, s = i % SLOTS;
    const int f = f_lo + s;
    float2 v = make_float2(0.f, 0.f);
    if (f >= 0 && f < p.frames) v = __ldg(sp + (long long)k * p.frames + f);
    if (k == 0 || k == N) v.y = 0.f;                     // c2r ignores the imaginary part of DC / Nyquist
    if (k < N) buf[s * NPAD + pad(k)] = v;
    else xN[s] = v;
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int slot = r * kFPB + g;
    float2* mybuf = buf + slot * NPAD;
    const float2 xn = xN[slot];
    auto load0 = [&](int k) {
      const float2 xk = mybuf[pad(k)];
      const float2 xnk = k == 0 ? xn : mybuf[pad(N - k)];
      return irfft_pack(xk, cconj(xnk), tw2[k]);
    };
    transform<N, true>(t, load0, mybuf, tw, [g] { group_sync(g, T); });
  }
  __syncthreads();
  // gather overlap-add
  const float inv_n = 1.0f / N;
  float* o = p.wav + (long long)b * p.Nout;
  for (int i = threadIdx.x; i < opb; i += blockDim.x) {
    const int n = n0 + i;
    if (n >= p.Nout) break;
    const int j = n + p.joff;                            // padded-signal coordinate
    const int f_hi = j / p.hop;
    float acc = 0.f, env = 0.f;
    for (int q = 0; q < ratio; ++q) {
      const int f = f_hi - q;
      const int idx = j - f * p.hop;                     // position inside frame f
      if (f < 0 || f >= p.frames || idx >= NFFT) continue;
      const float w = win[idx];
      const float2 z = buf[(f - f_lo) * NPAD + pad(idx >> 1)];
      acc = fmaf((idx & 1) ? z.y : z.x, w * inv_n, acc);
      env = fmaf(w, w, env);
    }
    o[n] = !p.normalize ? acc : env > 1e-11f ? acc / env : 0.f;
  }
}

// ------------------------------------------------------------------ fast iSTFT, n_fft = 1024
// 16 frame slots per CTA (2 rounds x 8 warps), one warp per frame (fft_warp.cuh, inverse).  The
// spectrum block is loaded coalesced (frames are the contiguous axis), every warp packs its frame's
// half spectrum into the 512-point complex input on the fly, runs the inverse FFT in registers and
// writes the windowed, scaled samples back over its slot; then every thread gathers n_fft/hop
// contributions per output sample and divides by the window envelope.
constexpr int kISlotsLog = 5, kISlots = 1 << kISlotsLog, kISlotStride = 577;     // float2 per slot (odd stride: conflict-free block load)

__global__ void __launch_bounds__(kISlots * 32, kISlots == 16 ? 2 : 1)
istft1024_kernel(const IstftParams p) {
  constexpr int NFFT = 1024, N = 512;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* tw2 = reinterpret_cast<float2*>(smem_raw);      // [N + 2]
  float2* win2 = tw2 + (N + 2);                            // [N]
  float2* tw1 = win2 + N;                                  // [16][32]
  float2* xN = tw1 + 512;                                  // [kISlots] Nyquist bins
  float2* slots = xN + kISlots;                            // [kISlots][kISlotStride]: X[k] (padded) -> transpose scratch -> y[m]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ratio = NFFT / p.hop;
  const int opb = (kISlots - ratio + 1) * p.hop;
  const int n0 = blockIdx.x * opb;
  const int f_lo = (n0 + p.joff - NFFT) / p.hop + 1;       // may be negative (the numerator is a multiple of hop)
  const int bins = N + 1;

  for (int i = threadIdx.x; i <= N; i += blockDim.x) tw2[i] = __ldg(p.tab.tw2 + i);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    win2[i] = __ldg(reinterpret_cast<const float2*>(p.tab.win) + i);
    tw1[i] = __ldg(p.tab.tw1 + i);
  }
  {
    // thread = (frame slot s = tid & 15, bin k = tid >> 4 + 32*j): 16 slots of one bin are one 128-byte
    // line of spec[b, k, f_lo ..]; loads are issued 9 at a time so their latencies overlap
    const int sidx = threadIdx.x & (kISlots - 1), k0 = threadIdx.x >> kISlotsLog;
    const int f = f_lo + sidx;
    const bool fok = f >= 0 && f < p.frames;
    const float2* sp = p.spec + (long long)b * bins * p.frames + (fok ? f : 0);
    float2* dst = slots + sidx * kISlotStride;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      float2 v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int k = k0 + 32 * (jj * 9 + j);
        v[j] = (fok && k < bins) ? __ldg(sp + (long long)k * p.frames) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int k = k0 + 32 * (jj * 9 + j);
        if (k == 0 || k == N) v[j].y = 0.f;                // c2r ignores the imaginary part of DC / Nyquist
        if (k < N) dst[k] = v[j];            // linear slots: the Stockham k + k/8 padding 2-way conflicts here
        else if (k == N) xN[sidx] = v[j];
      }
    }
  }
  __syncthreads();
  const float inv_n = 1.0f / N;
  {
    const int slot = warp;                                 // one warp per frame slot
    float2* X = slots + slot * kISlotStride;
    const float2 xn = xN[slot];
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int k = 32 * n1 + lane;
      const float2 xk = X[k];
      const float2 xnk = k == 0 ? xn : X[N - k];
      v[n1] = irfft_pack(xk, cconj(xnk), tw2[k]);
    }
    __syncwarp();                                          // all reads of X done before it is overwritten
    warp_fft512<true>(v, X, tw1, lane);                  // the slot doubles as the transpose scratch
    // lane (k1, p) holds z[m], m = k1 + 16*k2 + 256*p  ->  samples (2m, 2m+1), windowed and scaled
    const int mb = (lane & 15) + 256 * (lane >> 4);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const int m = mb + 16 * k2;
      const float2 w = win2[m];
      X[m] = make_float2(v[k2].x * w.x * inv_n, v[k2].y * w.y * inv_n);
    }
  }
  __syncthreads();
  float* o = p.wav + (long long)b * p.Nout;
  const float* wv = reinterpret_cast<const float*>(win2);
  // gather overlap-add.  Fast path (hop a power of two, every frame of this CTA's block inside the signal): thread =
  // sample offset within a hop, no divisions, no per-contribution bounds tests; the window envelope sum_q w^2 depends
  // only on the offset, so it is computed once per thread.  (The generic loop below spent ~150 instructions per
  // sample on index arithmetic: the ALU pipe was this kernel's busiest unit, profiles/r01_final_stft_kernels_ncu.txt.)
  const int hop = p.hop;
  const bool pow2 = (hop & (hop - 1)) == 0 && hop <= (int)blockDim.x && (p.joff % hop) == 0;
  const bool interior = f_lo >= 0 && f_lo + kISlots <= p.frames;
  if (pow2 && interior) {
    const int nhop = opb / hop;                          // output hops of this CTA (kISlots - ratio + 1)
    const int lanes_per = blockDim.x / hop;              // threads that share one offset (they split the hops)
    const int r = threadIdx.x % hop, part = threadIdx.x / hop;
    float env = 0.f;
    for (int q = 0; q < ratio; ++q) { const float w = wv[r + q * hop]; env = fmaf(w, w, env); }
    const float inv_env = p.normalize ? (env > 1e-11f ? 1.0f / env : 0.f) : 1.0f;
    // output sample n = n0 + h * hop + r: padded coordinate j = n + joff, newest frame f_hi = j / hop, contribution q
    // comes from slot (f_hi - q - f_lo) at position r + q * hop
    const int s_hi0 = (n0 + p.joff) / hop - f_lo;        // slot of f_hi for h = 0 (= ratio - 1)
    const float* flat = reinterpret_cast<const float*>(slots);
    constexpr int S2 = 2 * kISlotStride;                 // floats per slot
    if (ratio == 4) {
      // hop = n_fft / 4 (the reference's setting): the four contributions of a sample sit at fixed offsets from one
      // running pointer -- 4 loads, 3 adds, 1 multiply, 1 store per sample (the generic loop below spent ~50
      // instructions per sample on index arithmetic: 17 % of the kernel's instructions)
      const float* pq = flat + (s_hi0 + part) * S2 + r;
      float* on = o + n0 + part * hop + r;
      const int step = lanes_per * S2, ostep = lanes_per * hop;
      int n = n0 + part * hop + r;
#pragma unroll 2
      for (int h = part; h < nhop && n < p.Nout; h += lanes_per, pq += step, on += ostep, n += ostep) {
        const float acc = ((pq[0] + pq[hop - S2]) + pq[2 * hop - 2 * S2]) + pq[3 * hop - 3 * S2];
        *on = p.normalize ? acc * inv_env : acc;
      }
      return;
    }
    for (int h = part; h < nhop; h += lanes_per) {
      const int n = n0 + h * hop + r;
      if (n >= p.Nout) break;
      float acc = 0.f;
#pragma unroll 4
      for (int q = 0; q < ratio; ++q) {
        const int idx = r + q * hop;
        acc += flat[(s_hi0 + h - q) * S2 + idx];          // sample idx of the frame: one 4-byte load
      }
      o[n] = p.normalize ? acc * inv_env : acc;
    }
    return;
  }
  for (int i = threadIdx.x; i < opb; i += blockDim.x) {
    const int n = n0 + i;
    if (n >= p.Nout) break;
    const int j = n + p.joff;                              // padded-signal coordinate
    const int f_hi = j / p.hop;
    float acc = 0.f, env = 0.f;
    for (int q = 0; q < ratio; ++q) {
      const int f = f_hi - q;
      const int idx = j - f * p.hop;
      if (f < 0 || f >= p.frames || idx >= NFFT) continue;
      const float w = wv[idx];
      acc += reinterpret_cast<const float*>(slots + (f - f_lo) * kISlotStride)[idx];
      env = fmaf(w, w, env);
    }
    o[n] = !p.normalize ? acc : env > 1e-11f ? acc / env : 0.f;
  }
}

static int launch_istft1024(const IstftParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(514 + 512 + 512 + kISlots + kISlots * kISlotStride) * 8 + 64;
  B200_CUDA(cudaFuncSetAttribute(istft1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ratio = 1024 / p.hop;
  const int opb = (kISlots - ratio + 1) * p.hop;
  dim3 grid(ceil_div(p.Nout, opb), p.B);
  istft1024_kernel<<<grid, kISlots * 32, smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

template <int NFFT>
static int launch_istft(const IstftParams& p, cudaStream_t st) {
  constexpr int N = NFFT / 2, NPAD = N + N / 8, SLOTS = 2 * kFPB;
  const size_t smem = (size_t)NFFT * 4 + (size_t)N * 8 + (size_t)(N + 2) * 8 + SLOTS * 8 + (size_t)SLOTS * NPAD * 8 + 64;
  B200_CUDA(cudaFuncSetAttribute(istft_kernel<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ratio = NFFT / p.hop;
  const int opb = (SLOTS - ratio + 1) * p.hop;
  dim3 grid(ceil_div(p.Nout, opb), p.B);
  istft_kernel<NFFT><<<grid, kFPB * (NFFT / 16), smem, st>>>(p);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}


// ------------------------------------------------------------------ STFTLoss backward (stft.py:48-54)
// One resolution of  loss = scale * mean_{b,k,m} | |X_f| g_k - |X_r| g_k |  (scale = lambda * upstream grad):
//   dL/dX_f[k,m]   = scale/numel * |g_k| * sign(|X_f| - |X_r|) * X_f / |X_f|
//   dL/dwav_f      = STFT^T (dL/dX_f): per frame  w[n] * sum_{k=0..N/2} Re(G_k e^{+2 pi i k n / N}),  overlap-added
//                    on the padded signal, then folded back through the reflect padding.
// The per-frame sum is evaluated by the inverse real FFT of this file with Y_0 = N G_0, Y_{N/2} = N G_{N/2}
// (real parts only: Im X_0 = Im X_{N/2} = 0 identically), Y_k = (N/2) G_k  (irfft weights undone).
__global__ void __launch_bounds__(256) stft_l1_grad_spec_kernel(float2* __restrict__ xf, const float* __restrict__ magr,
                                                                 const float* __restrict__ gain, int bins, int frames,
                                                                 long long total, float w_elem, int n_fft,
                                                                 float* __restrict__ grad_gain) {
  // block = 256 consecutive elements of [B, bins, frames]; grad_gain accumulates per-bin |d| sums
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  int k = -1;
  if (i < total) {
    k = (int)((i / frames) % bins);
    const float2 x = xf[i];
    const float a = sqrtf(x.x * x.x + x.y * x.y);
    const float d = a - magr[i];
    const float g = __ldg(gain + k);
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const bool edge = k == 0 || k == bins - 1;
    const float coef = (edge ? (float)n_fft : 0.5f * n_fft) * w_elem * fabsf(g) * sgn;
    const float inv = a > 0.f ? coef / a : 0.f;
    xf[i] = make_float2(x.x * inv, edge ? 0.f : x.y * inv);
    contrib = w_elem * (g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f)) * fabsf(d);
  }
  if (grad_gain != nullptr) {
    // consecutive threads mostly share k (frames >> 32): reduce runs of equal k inside the warp
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k0 = __shfl_sync(full, k, 0);
    if (__all_sync(full, k == k0)) {
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(full, contrib, o);
      if (lane == 0 && k0 >= 0) atomicAdd(grad_gain + k0, contrib);
    } else if (k >= 0) {
      atomicAdd(grad_gain + k, contrib);
    }
  }
}

// General form for LearnableSTFT.forward (stft.py:22-34), out = |X| * gain with an arbitrary upstream gradient G:
//   dL/dX = G * gain * X / |X|  (same irfft weights as above),  dL/dgain[k] = sum_{b,t} G * |X|.
__global__ void __launch_bounds__(256) stft_mag_grad_spec_kernel(float2* __restrict__ xf, const float* __restrict__ gout,
                                                                  const float* __restrict__ gain, int bins, int frames,
                                                                  long long total, int n_fft, float* __restrict__ grad_gain) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  int k = -1;
  if (i < total) {
    k = (int)((i / frames) % bins);
    const float2 x = xf[i];
    const float a = sqrtf(x.x * x.x + x.y * x.y);
    const float go = gout[i];
    const float g = gain ? __ldg(gain + k) : 1.f;
    const bool edge = k == 0 || k == bins - 1;
    const float coef = (edge ? (float)n_fft : 0.5f * n_fft) * g * go;
    const float inv = a > 0.f ? coef / a : 0.f;
    xf[i] = make_float2(x.x * inv, edge ? 0.f : x.y * inv);
    contrib = go * a;
  }
  if (grad_gain != nullptr) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k0 = __shfl_sync(full, k, 0);
    if (__all_sync(full, k == k0)) {
      for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(full, contrib, o);
      if (lane == 0 && k0 >= 0) atomicAdd(grad_gain + k0, contrib);
    } else if (k >= 0) {
      atomicAdd(grad_gain + k, contrib);
    }
  }
}

// grad_wav[b, n] += P[b, pad + n] + reflected contributions of the two padded ends (P is the overlap-add on
// the padded signal of length N + 2 pad):  x_pad[j] = x[pad - j] (j < pad),  x_pad[pad + N + i] = x[N - 2 - i].
__global__ void __launch_bounds__(256) stft_fold_reflect_kernel(const float* __restrict__ P, int B, int N, int pad,
                                                                 float* __restrict__ grad_wav) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N), n = (int)(i % N);
  const float* p = P + (long long)b * (N + 2 * pad);
  float g = p[pad + n];
  if (n >= 1 && n <= pad) g += p[pad - n];
  if (n >= N - 1 - pad && n <= N - 2) g += p[pad + 2 * N - 2 - n];
  grad_wav[i] += g;
}

}  // namespace b200

using namespace b200;

// Adjoint of the framed, windowed forward transform applied to a prepared spectral gradient (irfft weights undone by the
// caller): plain overlap-add on the padded signal (the iSTFT kernels without the envelope division), then the reflect
// padding folded back; ACCUMULATES into grad_wav.
static int stft_adjoint_fold(float2* xf, float* P, int B, int N, int n_fft, int hop, float* grad_wav, cudaStream_t st) {
  const int frames = 1 + N / hop, pad = n_fft / 2;
  IstftParams p{};
  p.spec = xf; p.wav = P; p.B = B; p.frames = frames; p.hop = hop; p.Nout = N + n_fft; p.joff = 0; p.normalize = 0;
  B200_TRY(get_fft_tables(n_fft, &p.tab));
  int rc;
  switch (n_fft) {
    case 512: rc = launch_istft<512>(p, st); break;
    case 1024: rc = (n_fft / hop <= kISlots / 2) ? launch_istft1024(p, st) : launch_istft<1024>(p, st); break;
    case 2048: rc = launch_istft<2048>(p, st); break;
    default: set_error("stft backward: n_fft=%d unsupported (512/1024/2048)", n_fft); return B200VOC_ERR_UNSUPPORTED;
  }
  B200_TRY(rc);
  stft_fold_reflect_kernel<<<(unsigned)(((long long)B * N + 255) / 256), 256, 0, st>>>(P, B, N, pad, grad_wav);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

extern "C" {

/* Builds (and caches, per device) the tables the STFT family needs -- window, twiddles, and for n_mels > 0 the sparse
 * HTK mel filterbank -- so that the transform calls themselves never allocate or copy synchronously: call it once per
 * (n_fft, n_mels, sample_rate) before capturing the transforms in a CUDA graph or calling them from a latency-critical
 * loop.  Without it the first call of each configuration builds the tables lazily (cudaMalloc + synchronous copies). */
int b200voc_stft_prepare(int n_fft, int n_mels, int sample_rate) {
  if (n_fft != 512 && n_fft != 1024 && n_fft != 2048) {
    set_error("stft_prepare: n_fft=%d unsupported (512/1024/2048, vocoder7/config.py:39 stft_sizes)", n_fft);
    return B200VOC_ERR_UNSUPPORTED;
  }
  FftTables t;
  B200_TRY(get_fft_tables(n_fft, &t));
  if (n_mels > 0) {
    B200_CHECK_ARG(n_mels <= 512 && sample_rate > 0, "stft_prepare: bad mel arguments");
    MelTable m;
    B200_TRY(get_mel_table(n_fft, n_mels, sample_rate, &m));
  }
  return B200VOC_OK;
}

int b200voc_stft_mag(const float* wav, int B, int N, int n_fft, int hop, const float* gain, float* out, void* stream) {
  B200_TRY(check_stft_args(wav, B, N, n_fft, hop));
  B200_CHECK_ARG(out != nullptr, "stft_mag: null output");
  StftParams p{};
  p.wav = wav; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = 1 + N / hop; p.gain = gain; p.out = out;
  return dispatch_stft<MODE_MAG>(n_fft, p, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_stft_complex(const float* wav, int B, int N, int n_fft, int hop, float* out_ri, void* stream) {
  B200_TRY(check_stft_args(wav, B, N, n_fft, hop));
  B200_CHECK_ARG(out_ri != nullptr, "stft_complex: null output");
  StftParams p{};
  p.wav = wav; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = 1 + N / hop; p.out = out_ri;
  return dispatch_stft<MODE_COMPLEX>(n_fft, p, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_stft_logmel(const float* wav, int B, int N, int n_fft, int hop, int n_mels, int sample_rate,
                        int log_compress, float* out, void* stream) {
  B200_TRY(check_stft_args(wav, B, N, n_fft, hop));
  B200_CHECK_ARG(out != nullptr && n_mels > 0 && n_mels <= 512 && sample_rate > 0, "stft_logmel: bad arguments");
  StftParams p{};
  p.wav = wav; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = 1 + N / hop; p.out = out; p.log_compress = log_compress;
  B200_TRY(get_mel_table(n_fft, n_mels, sample_rate, &p.mel));
  return dispatch_stft<MODE_MEL>(n_fft, p, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_stft_l1(const float* wav_fake, const float* wav_real, int B, int N, int n_fft, int hop,
                    const float* gain, double* out_sum, void* stream) {
  B200_TRY(check_stft_args(wav_fake, B, N, n_fft, hop));
  B200_CHECK_ARG(wav_real != nullptr && out_sum != nullptr, "stft_l1: null argument");
  StftParams p{};
  p.wav = wav_fake; p.wav2 = wav_real; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = 1 + N / hop; p.gain = gain;
  p.out_sum = out_sum;
  return dispatch_stft<MODE_L1>(n_fft, p, reinterpret_cast<cudaStream_t>(stream));
}

int b200voc_istft(const float* spec_ri, int B, int frames, int n_fft, int hop, int N, float* wav, void* stream) {
  B200_CHECK_ARG(spec_ri && wav, "istft: null argument");
  B200_CHECK_ARG(B > 0 && frames > 0 && N > 0, "istft: empty input");
  B200_CHECK_ARG(hop > 0 && (n_fft / 2) % hop == 0, "istft: hop %d must divide n_fft/2 = %d", hop, n_fft / 2);
  B200_CHECK_ARG(n_fft / hop <= 2 * kFPB, "istft: n_fft/hop = %d too large", n_fft / hop);
  B200_CHECK_ARG((long long)N <= (long long)hop * (frames - 1) + n_fft / 2 + n_fft / 2,
                 "istft: length %d exceeds the signal the %d frames cover", N, frames);
  IstftParams p{};
  p.spec = reinterpret_cast<const float2*>(spec_ri); p.wav = wav; p.B = B; p.frames = frames; p.hop = hop; p.Nout = N;
  p.joff = n_fft / 2; p.normalize = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_fft == 512 || n_fft == 1024 || n_fft == 2048) B200_TRY(get_fft_tables(n_fft, &p.tab));
  switch (n_fft) {
    case 512: return launch_istft<512>(p, st);
    case 1024:
      if (n_fft / hop <= kISlots / 2 && !getenv("B200VOC_STFT_V1")) return launch_istft1024(p, st);
      return launch_istft<1024>(p, st);
    case 2048: return launch_istft<2048>(p, st);
  }
  set_error("istft: n_fft=%d unsupported (512/1024/2048)", n_fft);
  return B200VOC_ERR_UNSUPPORTED;
}


/* STFTLoss backward, one resolution (see the kernels above).  grad_wav and grad_gain are ACCUMULATED. */
int64_t b200voc_stft_l1_backward_workspace_bytes(int B, int N, int n_fft, int hop) {
  if (B <= 0 || N <= 0 || n_fft <= 0 || hop <= 0) return 0;
  const long long bins = n_fft / 2 + 1, frames = 1 + N / hop;
  const long long e = (long long)B * bins * frames;
  return e * 8 + e * 4 + (long long)B * (N + n_fft) * 4 + 1024;
}
int b200voc_stft_l1_backward(const float* wav_fake, const float* wav_real, int B, int N, int n_fft, int hop,
                             const float* gain, float scale, float* grad_wav, float* grad_gain, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  B200_TRY(check_stft_args(wav_fake, B, N, n_fft, hop));
  B200_CHECK_ARG(wav_real && gain && grad_wav && workspace, "stft_l1_backward: null argument");
  B200_CHECK_ARG(workspace_bytes >= b200voc_stft_l1_backward_workspace_bytes(B, N, n_fft, hop),
                 "stft_l1_backward: workspace too small");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "stft_l1_backward: workspace must be 16B aligned");
  B200_CHECK_ARG((n_fft / 2) % hop == 0 && n_fft / hop <= 2 * kFPB, "stft_l1_backward: hop %d must divide n_fft/2", hop);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bins = n_fft / 2 + 1, frames = 1 + N / hop;
  const long long e = (long long)B * bins * frames;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float2* xf = reinterpret_cast<float2*>(ws);
  float* magr = reinterpret_cast<float*>(ws + e * 8);
  float* P = reinterpret_cast<float*>(ws + ((e * 12 + 15) & ~15ll));
  {  // X_f (complex) and |X_r|
    StftParams p{};
    p.wav = wav_fake; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = frames; p.out = reinterpret_cast<float*>(xf);
    B200_TRY(dispatch_stft<MODE_COMPLEX>(n_fft, p, st));
    StftParams q{};
    q.wav = wav_real; q.B = B; q.Nsamp = N; q.hop = hop; q.frames = frames; q.gain = nullptr; q.out = magr;
    B200_TRY(dispatch_stft<MODE_MAG>(n_fft, q, st));
  }
  const float w_elem = scale / (float)e;
  stft_l1_grad_spec_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(xf, magr, gain, bins, frames, e, w_elem, n_fft,
                                                                        grad_gain);
  B200_CUDA(cudaGetLastError());
  return stft_adjoint_fold(xf, P, B, N, n_fft, hop, grad_wav, st);
}

int64_t b200voc_stft_mag_backward_workspace_bytes(int B, int N, int n_fft, int hop) {
  if (B <= 0 || N <= 0 || n_fft <= 0 || hop <= 0) return 0;
  const long long bins = n_fft / 2 + 1, frames = 1 + N / hop;
  const long long e = (long long)B * bins * frames;
  return e * 8 + (long long)B * (N + n_fft) * 4 + 1024;
}
int b200voc_stft_mag_backward(const float* wav, int B, int N, int n_fft, int hop, const float* gain, const float* grad_out,
                              float* grad_wav, float* grad_gain, void* workspace, int64_t workspace_bytes, void* stream) {
  B200_TRY(check_stft_args(wav, B, N, n_fft, hop));
  B200_CHECK_ARG(grad_out && grad_wav && workspace, "stft_mag_backward: null argument");
  B200_CHECK_ARG(workspace_bytes >= b200voc_stft_mag_backward_workspace_bytes(B, N, n_fft, hop),
                 "stft_mag_backward: workspace too small");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "stft_mag_backward: workspace must be 16B aligned");
  B200_CHECK_ARG((n_fft / 2) % hop == 0 && n_fft / hop <= 2 * kFPB, "stft_mag_backward: hop %d must divide n_fft/2", hop);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bins = n_fft / 2 + 1, frames = 1 + N / hop;
  const long long e = (long long)B * bins * frames;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float2* xf = reinterpret_cast<float2*>(ws);
  float* P = reinterpret_cast<float*>(ws + ((e * 8 + 15) & ~15ll));
  StftParams p{};
  p.wav = wav; p.B = B; p.Nsamp = N; p.hop = hop; p.frames = frames; p.out = reinterpret_cast<float*>(xf);
  B200_TRY(dispatch_stft<MODE_COMPLEX>(n_fft, p, st));
  stft_mag_grad_spec_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(xf, grad_out, gain, bins, frames, e, n_fft, grad_gain);
  B200_CUDA(cudaGetLastError());
  return stft_adjoint_fold(xf, P, B, N, n_fft, hop, grad_wav, st);
}

}  // extern "C"
