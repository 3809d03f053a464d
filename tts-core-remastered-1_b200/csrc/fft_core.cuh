// Shared-memory Stockham FFT building blocks for the STFT family (K7/K9).
//
// A length-n real frame is transformed as a length-N = n/2 complex FFT plus a split
// (real-FFT) post-pass.  N/8 threads own one frame; every pass keeps 8 complex values per thread
// in registers (radix-8 butterflies; one radix-4 or radix-16 pass fixes up N = 256 / 1024), and the
// passes exchange data through a padded shared-memory buffer (index i -> i + i/8 keeps the
// 8-strided writes bank-conflict free).  Everything is __host__ __device__ so the index math and
// twiddles are unit-tested on the CPU (tests/host/fft_core_host.cu) without a GPU.
#pragma once
#include <cuda_runtime.h>

namespace b200 {
namespace fft {

#define B200_HD __host__ __device__ __forceinline__

B200_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B200_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B200_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
B200_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
B200_HD int pad(int i) { return i + (i >> 3); }

// exp(-2*pi*i*k/16), k = 0..7
template <int K16>
B200_HD float2 w16() {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  if constexpr (K16 == 0) return make_float2(1.f, 0.f);
  else if constexpr (K16 == 1) return make_float2(c1, -s1);
  else if constexpr (K16 == 2) return make_float2(h, -h);
  else if constexpr (K16 == 3) return make_float2(s1, -c1);
  else if constexpr (K16 == 4) return make_float2(0.f, -1.f);
  else if constexpr (K16 == 5) return make_float2(-s1, -c1);
  else if constexpr (K16 == 6) return make_float2(-h, -h);
  else return make_float2(-c1, -s1);
}

// x * exp(-+2*pi*i*K16/16) with the trivial cases (1, -+i, (1 -+ i)/sqrt2) strength-reduced: without
// fast-math the compiler may not fold multiplications by 0 and 1.
template <int K16, bool INV>
B200_HD float2 mul_w16(float2 x) {
  constexpr float h = 0.70710678118654752440f;
  if constexpr (K16 == 0) return x;
  else if constexpr (K16 == 4) return INV ? make_float2(-x.y, x.x) : make_float2(x.y, -x.x);
  else if constexpr (K16 == 2) return INV ? make_float2(h * (x.x - x.y), h * (x.x + x.y)) : make_float2(h * (x.x + x.y), h * (x.y - x.x));
  else if constexpr (K16 == 6) return INV ? make_float2(-h * (x.x + x.y), h * (x.x - x.y)) : make_float2(h * (x.y - x.x), -h * (x.x + x.y));
  else {
    float2 w = w16<K16>();
    if (INV) w = cconj(w);
    return cmul(x, w);
  }
}

template <int R, bool INV, int K>
struct Combine {
  B200_HD static void run(const float2* e, const float2* o, float2* v) {
    if constexpr (K < R / 2) {
      const float2 t = mul_w16<K * (16 / R), INV>(o[K]);
      v[K] = cadd(e[K], t);
      v[K + R / 2] = csub(e[K], t);
      Combine<R, INV, K + 1>::run(e, o, v);
    }
  }
};

// in-register DFT of R points (natural order in and out); INV uses the +i sign, unnormalised.
template <int R, bool INV>
B200_HD void fft_reg(float2* v) {
  if constexpr (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R > 2) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      e[k] = v[2 * k];
      o[k] = v[2 * k + 1];
    }
    fft_reg<R / 2, INV>(e);
    fft_reg<R / 2, INV>(o);
    Combine<R, INV, 0>::run(e, o, v);
  }
}

// One Stockham pass of radix R over a length-N transform whose previous radices multiply to NS.
// T = N/8 threads per transform; thread t owns butterflies t, t+T, ...  The read half (load,
// twiddle, butterfly into `v`) and the write half are separate so the caller can put the barrier
// between them (in-place exchange).  `tw[m] = exp(-2*pi*i*m/N)`.
template <int N, int R, int NS, bool INV>
struct Pass {
  static constexpr int T = N / 8;
  static constexpr int NB = N / R;
  static constexpr int ITER = (NB + T - 1) / T;
  static constexpr int VALS = ITER * R;

  template <typename Loader>
  B200_HD static void read(int t, Loader load, const float2* tw, float2* v) {
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int j = t + it * T;
      if (j < NB) {
        const int k = j % NS;
        float2* vv = v + it * R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float2 x = load(j + r * NB);
          if (NS > 1 && r > 0) {
            float2 w = tw[(r - 1) * NS + k];          // pass table: conflict-free in k
            if (INV) w = cconj(w);
            x = cmul(x, w);
          }
          vv[r] = x;
        }
        fft_reg<R, INV>(vv);
      }
    }
  }
  B200_HD static void write(int t, float2* buf, const float2* v) {
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int j = t + it * T;
      if (j < NB) {
        const int k = j % NS;
        const int j0 = (j / NS) * NS * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[pad(j0 + r * NS)] = v[it * R + r];
      }
    }
  }
};

// radix plans: N = 256 -> 4,8,8 ; 512 -> 8,8,8 ; 1024 -> 8,8,16
template <int N> struct Plan;
template <> struct Plan<256> { static constexpr int R0 = 4, R1 = 8, R2 = 8; };
template <> struct Plan<512> { static constexpr int R0 = 8, R1 = 8, R2 = 8; };
template <> struct Plan<1024> { static constexpr int R0 = 8, R1 = 8, R2 = 16; };

// Twiddle storage: per pass p (previous radices multiply to NS, radix R) a table
//   T_p[(r-1)*NS + k] = exp(-2*pi*i * r*k / (NS*R)),  r = 1..R-1, k = 0..NS-1
// laid out so that the threads of a pass (k = j % NS consecutive) read consecutive words.
// Pass 0 needs none; pass 1 starts at offset 0, pass 2 after it.
template <int N> B200_HD constexpr int tw_offset1() { return 0; }
template <int N> B200_HD constexpr int tw_offset2() { return (Plan<N>::R1 - 1) * Plan<N>::R0; }
template <int N> B200_HD constexpr int tw_total() {
  return (Plan<N>::R1 - 1) * Plan<N>::R0 + (Plan<N>::R2 - 1) * Plan<N>::R0 * Plan<N>::R1;
}
// host helper: fill the pass tables
template <int N>
inline void fill_pass_twiddles(float2* out) {
  using P = Plan<N>;
  int o = 0;
  for (int pass = 1; pass <= 2; ++pass) {
    const int NS = pass == 1 ? P::R0 : P::R0 * P::R1, R = pass == 1 ? P::R1 : P::R2;
    for (int r = 1; r < R; ++r)
      for (int k = 0; k < NS; ++k) {
        const double a = -2.0 * 3.14159265358979323846 * (double)r * k / ((double)NS * R);
        out[o++] = make_float2((float)cos(a), (float)sin(a));
      }
  }
}

// Full N-point transform by the N/8 threads of one frame group.  `sync` is the barrier between
// the phases (a no-op lambda in the sequential host emulation, which instead loops over t per
// phase).  After the call `buf[pad(k)]` holds bin k.
template <int N, bool INV, typename Loader, typename Sync>
B200_HD void transform(int t, Loader load0, float2* buf, const float2* tw, Sync sync) {
  using P = Plan<N>;
  using P0 = Pass<N, P::R0, 1, INV>;
  using P1 = Pass<N, P::R1, P::R0, INV>;
  using P2 = Pass<N, P::R2, P::R0 * P::R1, INV>;
  auto from_buf = [buf](int i) { return buf[pad(i)]; };
  float2 v[16];
  P0::read(t, load0, tw, v);
  sync();
  P0::write(t, buf, v);
  sync();
  P1::read(t, from_buf, tw + tw_offset1<N>(), v);
  sync();
  P1::write(t, buf, v);
  sync();
  P2::read(t, from_buf, tw + tw_offset2<N>(), v);
  sync();
  P2::write(t, buf, v);
  sync();
}

// Real-FFT split: Z = FFT_N(z), z[m] = x[2m] + i x[2m+1]  ->  X[k], k = 0..N (n = 2N real points)
//   X[k] = (Z[k] + conj(Z[N-k]))/2 - i/2 * exp(-i*pi*k/N) * (Z[k] - conj(Z[N-k]))
// `tw2[k] = exp(-i*pi*k/N)`, k = 0..N  (half-step twiddles).
B200_HD float2 rfft_bin(const float2* buf, const float2* tw2, int N, int k) {
  const float2 zk = buf[pad(k % N)];
  const float2 zn = cconj(buf[pad((N - k) % N)]);
  const float2 a = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
  const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y - zn.y));
  const float2 wd = cmul(tw2[k], d);                 // exp(-i pi k/N) * d
  return make_float2(a.x + wd.y, a.y - wd.x);        // a - i*wd
}

// inverse split: given X[k] (k = 0..N) build Z[k] (k = 0..N-1) such that IFFT_N(Z)[m] =
// x[2m] + i x[2m+1] (unnormalised: result is N * ..., caller scales by 1/N... see istft kernel):
//   Z[k] = (X[k] + conj(X[N-k])) + i * exp(+i*pi*k/N) * (X[k] - conj(X[N-k]))   (times 1/2)
B200_HD float2 irfft_pack(float2 xk, float2 xnk_conj, float2 tw2k) {
  const float2 a = make_float2(0.5f * (xk.x + xnk_conj.x), 0.5f * (xk.y + xnk_conj.y));
  const float2 d = make_float2(0.5f * (xk.x - xnk_conj.x), 0.5f * (xk.y - xnk_conj.y));
  const float2 wd = cmul(cconj(tw2k), d);            // exp(+i pi k/N) * d
  return make_float2(a.x - wd.y, a.y + wd.x);        // a + i*wd
}

}  // namespace fft
}  // namespace b200
