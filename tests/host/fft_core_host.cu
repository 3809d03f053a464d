// Host-side unit test of csrc/fft_core.cuh (test infrastructure): emulates the N/8 threads of one
// frame group sequentially (phase by phase, which is what the barriers enforce on the GPU) and
// compares forward/inverse transforms and the real-FFT split against a naive double-precision DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../tts-core-remastered-1_b200/csrc/fft_core.cuh"

using namespace b200::fft;

template <int N, int R, int NS, bool INV, typename Loader>
void run_pass(Loader load, float2* buf, const float2* tw) {
  using P = Pass<N, R, NS, INV>;
  std::vector<float2> regs(P::T * 16);
  for (int t = 0; t < P::T; ++t) P::read(t, load, tw, &regs[t * 16]);
  for (int t = 0; t < P::T; ++t) P::write(t, buf, &regs[t * 16]);
}

template <int N, bool INV>
double test_complex() {
  std::vector<float2> in(N), tw(N), buf(N + N / 8 + 8);
  for (int i = 0; i < N; ++i) in[i] = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
  fill_pass_twiddles<N>(tw.data());
  using PL = Plan<N>;
  auto load0 = [&](int i) { return in[i]; };
  auto from_buf = [&](int i) { return buf[pad(i)]; };
  run_pass<N, PL::R0, 1, INV>(load0, buf.data(), tw.data());
  run_pass<N, PL::R1, PL::R0, INV>(from_buf, buf.data(), tw.data() + tw_offset1<N>());
  run_pass<N, PL::R2, PL::R0 * PL::R1, INV>(from_buf, buf.data(), tw.data() + tw_offset2<N>());
  double maxerr = 0;
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const double ang = (INV ? 2.0 : -2.0) * M_PI * (double)k * n / N;
      re += in[n].x * cos(ang) - in[n].y * sin(ang);
      im += in[n].x * sin(ang) + in[n].y * cos(ang);
    }
    maxerr = fmax(maxerr, fmax(fabs(re - buf[pad(k)].x), fabs(im - buf[pad(k)].y)));
  }
  return maxerr;
}

// real FFT of n = 2N points through the packed complex transform + split, and the inverse packing
template <int N>
double test_real(double* inv_err) {
  const int n = 2 * N;
  std::vector<float> x(n);
  std::vector<float2> tw(N), tw2(N + 1), buf(N + N / 8 + 8);
  for (int i = 0; i < n; ++i) x[i] = (float)rand() / RAND_MAX - 0.5f;
  fill_pass_twiddles<N>(tw.data());
  for (int i = 0; i <= N; ++i) tw2[i] = make_float2((float)cos(-M_PI * i / N), (float)sin(-M_PI * i / N));
  using PL = Plan<N>;
  auto load0 = [&](int m) { return make_float2(x[2 * m], x[2 * m + 1]); };
  auto from_buf = [&](int i) { return buf[pad(i)]; };
  run_pass<N, PL::R0, 1, false>(load0, buf.data(), tw.data());
  run_pass<N, PL::R1, PL::R0, false>(from_buf, buf.data(), tw.data() + tw_offset1<N>());
  run_pass<N, PL::R2, PL::R0 * PL::R1, false>(from_buf, buf.data(), tw.data() + tw_offset2<N>());
  std::vector<float2> X(N + 1);
  double maxerr = 0;
  for (int k = 0; k <= N; ++k) {
    X[k] = rfft_bin(buf.data(), tw2.data(), N, k);
    double re = 0, im = 0;
    for (int j = 0; j < n; ++j) {
      re += x[j] * cos(-2.0 * M_PI * (double)k * j / n);
      im += x[j] * sin(-2.0 * M_PI * (double)k * j / n);
    }
    maxerr = fmax(maxerr, fmax(fabs(re - X[k].x), fabs(im - X[k].y)));
  }
  // inverse: pack -> inverse complex FFT -> x[2m] + i x[2m+1] (scaled by N)
  std::vector<float2> Z(N);
  for (int k = 0; k < N; ++k) Z[k] = irfft_pack(X[k], cconj(X[N - k]), tw2[k]);
  auto loadz = [&](int i) { return Z[i]; };
  run_pass<N, PL::R0, 1, true>(loadz, buf.data(), tw.data());
  run_pass<N, PL::R1, PL::R0, true>(from_buf, buf.data(), tw.data() + tw_offset1<N>());
  run_pass<N, PL::R2, PL::R0 * PL::R1, true>(from_buf, buf.data(), tw.data() + tw_offset2<N>());
  double ie = 0;
  for (int m = 0; m < N; ++m) {
    ie = fmax(ie, fabs(buf[pad(m)].x / N - x[2 * m]));
    ie = fmax(ie, fabs(buf[pad(m)].y / N - x[2 * m + 1]));
  }
  *inv_err = ie;
  return maxerr;
}

int main() {
  srand(7);
  int bad = 0;
  auto chk = [&](const char* name, double e, double tol) {
    printf("%-28s max err %.3e (tol %.1e) %s\n", name, e, tol, e <= tol ? "ok" : "FAIL");
    if (!(e <= tol)) ++bad;
  };
  chk("fft256 fwd", test_complex<256, false>(), 2e-5);
  chk("fft512 fwd", test_complex<512, false>(), 3e-5);
  chk("fft1024 fwd", test_complex<1024, false>(), 5e-5);
  chk("fft512 inv", test_complex<512, true>(), 3e-5);
  double ie;
  chk("rfft512 (N=256)", test_real<256>(&ie), 3e-5); chk("  irfft roundtrip", ie, 2e-6);
  chk("rfft1024 (N=512)", test_real<512>(&ie), 5e-5); chk("  irfft roundtrip", ie, 2e-6);
  chk("rfft2048 (N=1024)", test_real<1024>(&ie), 8e-5); chk("  irfft roundtrip", ie, 2e-6);
  return bad;
}
