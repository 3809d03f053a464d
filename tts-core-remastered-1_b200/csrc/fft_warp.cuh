// Warp-per-frame 512-point complex FFT (the packed real FFT of an n_fft = 1024 frame) with both
// radix-16 stages in registers:
//   N = 16 x 32.  Lane t holds z[32*n1 + t], n1 = 0..15 -> 16-point FFT over n1 in registers,
//   twiddle W_512^(t*k1) (15 per-lane constants kept in registers across frames), ONE shared-memory
//   transpose, then each of the 16 size-32 column FFTs is done by a lane pair: lane (k1, p) runs a
//   16-point FFT over the even (p=0) or odd (p=1) column entries in registers, odd results get
//   W_32^k2, and a single shfl_xor(16) exchange finishes the radix-2 combine.
// Result: lane (k1 = t & 15, p = t >> 4) holds Z[k1 + 16*k2 + 256*p] in v[k2].
// Compared with the shared-memory Stockham path this is 1 exchange instead of 5 and ~2.6x fewer
// instructions per frame.
#pragma once
#include "fft_core.cuh"

namespace b200 {
namespace fft {

// exp(-2*pi*i*K/32), K = 0..15, as compile-time constants
template <int K>
__device__ __forceinline__ float2 w32() {
  // cos(pi*K/16), sin(pi*K/16) for K = 0..8
  constexpr float c[9] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                          0.19509032201612826785f, 0.0f};
  constexpr float cr = K <= 8 ? c[K] : -c[16 - K];
  constexpr float si = K <= 8 ? c[8 - K] : c[K - 8];
  return make_float2(cr, -si);
}

template <bool INV, int K>
struct OddTwiddle {
  __device__ __forceinline__ static void run(float2* v, bool odd) {
    if constexpr (K < 16) {
      float2 w = w32<K>();
      if (INV) w = cconj(w);
      const float2 m = cmul(v[K], w);
      v[K] = odd ? m : v[K];
      OddTwiddle<INV, K + 1>::run(v, odd);
    }
  }
};

// T: per-warp scratch of 16 x 33 float2.  tw1[k1 * 32 + lane] = exp(-2*pi*i*lane*k1/512) (shared memory,
// conflict-free: consecutive lanes read consecutive words).
template <bool INV>
__device__ __forceinline__ void warp_fft512(float2 (&v)[16], float2* T, const float2* tw1, int lane) {
  fft_reg<16, INV>(v);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) {
    const float2 w = tw1[k1 * 32 + lane];
    v[k1] = cmul(v[k1], INV ? cconj(w) : w);
  }
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) T[k1 * 33 + lane] = v[k1];
  __syncwarp();
  const int k1p = lane & 15, p = lane >> 4;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = T[k1p * 33 + 2 * q + p];
  __syncwarp();
  fft_reg<16, INV>(v);
  OddTwiddle<INV, 1>::run(v, p != 0);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float ox = __shfl_xor_sync(0xffffffffu, v[k2].x, 16), oy = __shfl_xor_sync(0xffffffffu, v[k2].y, 16);
    v[k2] = p ? make_float2(ox - v[k2].x, oy - v[k2].y) : make_float2(v[k2].x + ox, v[k2].y + oy);
  }
}

// Same transform with the stage-1 twiddles W_512^(lane*k1), k1 = 1..15, held in registers by the caller
// (tw[k1], tw[0] unused): 15 fewer 64-bit shared-memory loads per frame at the price of 30 registers.
template <bool INV>
__device__ __forceinline__ void warp_fft512_regtw(float2 (&v)[16], float2* T, const float2 (&tw)[16], int lane) {
  fft_reg<16, INV>(v);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], INV ? cconj(tw[k1]) : tw[k1]);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) T[k1 * 33 + lane] = v[k1];
  __syncwarp();
  const int k1p = lane & 15, p = lane >> 4;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = T[k1p * 33 + 2 * q + p];
  __syncwarp();
  fft_reg<16, INV>(v);
  OddTwiddle<INV, 1>::run(v, p != 0);
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float ox = __shfl_xor_sync(0xffffffffu, v[k2].x, 16), oy = __shfl_xor_sync(0xffffffffu, v[k2].y, 16);
    v[k2] = p ? make_float2(ox - v[k2].x, oy - v[k2].y) : make_float2(v[k2].x + ox, v[k2].y + oy);
  }
}

}  // namespace fft
}  // namespace b200
