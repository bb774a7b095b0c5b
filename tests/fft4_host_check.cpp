// Host-side check of the four-step FFT's per-thread code (psa_b200/csrc/fft4.cuh): the index maps, twiddles and
// butterflies of both stages are run thread by thread on the CPU and compared with a direct float64 transform.
// Built and run by tests/test_host.py (g++, no GPU).  Prints "OK <max relative error>" per length.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../psa_b200/csrc/fft4.cuh"

using namespace psa::fft4;
typedef std::complex<double> cplx;

static void fft_ref(std::vector<cplx>& a) {                 // iterative radix-2, float64
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double ang = -2.0 * M_PI * (double)k / (double)len;
        const cplx w(std::cos(ang), std::sin(ang));
        const cplx u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

template <int N1, int T>
static double run() {
  typedef Geo<N1, T> G;
  const int kThreads = T, kN1Tile = G::w;
  const int n = G::n, cols = kColsPerGroup;
  std::vector<c2> w128(128), tw(n), twb(G::q * 17), t1(16 * N1), t2(N1);
  for (int e = 0; e < 128; ++e) w128[e] = mk(std::cos(-2 * M_PI * e / 128.0), std::sin(-2 * M_PI * e / 128.0));
  for (int e = 0; e < n; ++e) tw[e] = mk(std::cos(-2 * M_PI * e / (double)n), std::sin(-2 * M_PI * e / (double)n));
  for (int j = 0; j < G::q; ++j)
    for (int s = 0; s < 16; ++s) twb[j * 17 + s] = tw[kN2 * j * s];
  for (int n1 = 0; n1 < N1; ++n1) {
    t2[n1] = tw[16 * n1];
    for (int s = 0; s < 16; ++s) t1[s * N1 + n1] = tw[n1 * s];
  }
  std::vector<std::vector<cplx>> x(cols, std::vector<cplx>(n));
  srand(N1);
  for (int c = 0; c < cols; ++c)
    for (int t = 0; t < n; ++t)
      x[c][t] = cplx((float)(rand() / (double)RAND_MAX - 0.5) + (c == 3 ? 30.0 * std::cos(2 * M_PI * 37 * t / n) : 0.0),
                     (float)(rand() / (double)RAND_MAX - 0.5));
  std::vector<c2> y((size_t)cols * n), exch(G::exchange_elems);
  // stage A: every column, every n1 tile
  for (int c = 0; c < cols; ++c)
    for (int tile = 0; tile < G::a_tiles_per_column; ++tile) {
      const int n1_0 = tile * kN1Tile;
      auto load = [&](int t) { return mk(x[c][t].real(), x[c][t].imag()); };
      for (int tid = 0; tid < kThreads; ++tid) stage_a_pass1<N1, T>(tid, n1_0, load, w128.data(), exch.data());
      for (int tid = 0; tid < kThreads; ++tid) stage_a_pass2<N1, T>(tid, n1_0, exch.data(), t1.data(), t2.data(), y.data() + (size_t)c * n);
    }
  // stage B: every k2 tile of the group
  std::vector<std::vector<cplx>> got(cols, std::vector<cplx>(n));
  std::vector<int> hits((size_t)cols * n, 0);
  for (int tile = 0; tile < G::tiles_per_group; ++tile) {
    const int k2_0 = tile * G::k2_per_tile;
    auto load_y = [&](int tau, int n1) { return y[((size_t)(tau & 15) * kN2 + k2_0 + (tau >> 4)) * N1 + n1]; };
    for (int tid = 0; tid < kThreads; ++tid) stage_b_pass1<N1, T>(tid, load_y, twb.data(), exch.data());
    auto sink = [&](int c, int k2l, int k1, c2 v) {
      const int f = k2_0 + k2l + kN2 * k1;
      got[c][f] = cplx(v.x, v.y);
      ++hits[(size_t)c * n + f];
    };
    for (int tid = 0; tid < kThreads; ++tid) stage_b_pass2<N1, T>(tid, exch.data(), sink);
  }
  double worst = 0.0;
  for (int c = 0; c < cols; ++c) {
    std::vector<cplx> ref = x[c];
    fft_ref(ref);
    double scale = 0.0;
    for (int f = 0; f < n; ++f) scale = std::max(scale, std::abs(ref[f]));
    for (int f = 0; f < n; ++f) {
      if (hits[(size_t)c * n + f] != 1) { std::printf("FAIL N1=%d: output (%d, %d) written %d times\n", N1, c, f, hits[(size_t)c * n + f]); std::exit(1); }
      worst = std::max(worst, std::abs(got[c][f] - ref[f]) / scale);
    }
  }
  return worst;
}

int main() {
  const double e[5] = {run<64, 256>(), run<128, 256>(), run<256, 256>(), run<64, 128>(), run<128, 128>()};
  std::printf("OK %.3e %.3e %.3e %.3e %.3e\n", e[0], e[1], e[2], e[3], e[4]);
  for (double v : e)
    if (!(v < 1e-13)) return 1;
  return 0;
}
