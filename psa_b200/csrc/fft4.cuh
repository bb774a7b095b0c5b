// Four-step time FFT for long power-of-two columns (n_t = 8192, 16384, 32768): the per-thread arithmetic.
//
// n = N1 * 128.  With t = n1 + N1 n2 (n1 < N1, n2 < 128) and f = k2 + 128 k1 (k2 < 128, k1 < N1):
//     X[k2 + 128 k1] = sum_n1 w_N1^(n1 k1) * [ w_n^(n1 k2) * sum_n2 x[n1 + N1 n2] w_128^(n2 k2) ]
// Stage A (the bracket): 128-point transforms over n2 for 32 adjacent n1 of one column -> Y[k2][n1] (float64).
// Stage B: N1-point transforms over n1 for 16 adjacent columns and 256 / N1 values of k2 -> the result, stored as
//          128-byte runs of the reference's (n_f, n_k, 3) layout (16 adjacent (k, pol) columns of one frequency).
// Both stages are two register passes (radix 16, then radix 8 / 4 / 8 / 16) around ONE shared-memory exchange;
// every global access is a full-width contiguous run.  Everything is float64 (see fft.cu for why): one float32
// rounding, at the final store.
//
// The functions take the thread index explicitly and touch memory only through plain pointers, so the same source
// runs on the host for the index / arithmetic checks in tests/fft4_host_check.cu (no GPU needed).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define F4_HD __host__ __device__ __forceinline__
#else
#define F4_HD inline
#endif

namespace psa {
namespace fft4 {

struct c2 {
  double x, y;
};
F4_HD c2 mk(double x, double y) { c2 r; r.x = x; r.y = y; return r; }
F4_HD c2 operator+(c2 a, c2 b) { return mk(a.x + b.x, a.y + b.y); }
F4_HD c2 operator-(c2 a, c2 b) { return mk(a.x - b.x, a.y - b.y); }
F4_HD c2 cmul(c2 a, c2 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
F4_HD c2 mul_neg_i(c2 a) { return mk(a.y, -a.x); }

constexpr int kN2 = 128;            // stage A transform length (over n2)
constexpr int kColsPerGroup = 16;   // adjacent (k, pol) columns stored together (128 bytes per frequency)
// A tile of either stage is 16 points per thread: T = 256 threads -> 4096 points (64 KiB exchange, 2 CTAs per SM),
// T = 128 -> 2048 points (32 KiB, 4 CTAs per SM: the same warps per SM, but four independent tile phases instead of
// two, and barriers that only span four warps).

// forward DFTs in registers, natural output order
F4_HD void bfly4(c2& a0, c2& a1, c2& a2, c2& a3) {
  const c2 t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_neg_i(a1 - a3);
  a0 = t0 + t2; a1 = t1 + t3; a2 = t0 - t2; a3 = t1 - t3;
}
F4_HD void dft8(c2 (&a)[8]) {
  const double h = 0.70710678118654752440;
  c2 u0 = a[0] + a[4], u1 = a[1] + a[5], u2 = a[2] + a[6], u3 = a[3] + a[7];
  c2 v0 = a[0] - a[4], d1 = a[1] - a[5], d2 = a[2] - a[6], d3 = a[3] - a[7];
  c2 v1 = mk((d1.x + d1.y) * h, (d1.y - d1.x) * h);        // * w8
  c2 v2 = mul_neg_i(d2);                                     // * w8^2
  c2 v3 = mk((d3.y - d3.x) * h, -(d3.x + d3.y) * h);        // * w8^3
  bfly4(u0, u1, u2, u3);
  bfly4(v0, v1, v2, v3);
  a[0] = u0; a[2] = u1; a[4] = u2; a[6] = u3;
  a[1] = v0; a[3] = v1; a[5] = v2; a[7] = v3;
}
// 16-point DFT; on return register 4 r + m holds frequency r + 4 m (see out16())
F4_HD void dft16(c2 (&x)[16]) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
#pragma unroll
  for (int j = 0; j < 4; ++j) bfly4(x[j], x[j + 4], x[j + 8], x[j + 12]);
  // x[j + 4 r] *= w16^(j r)
  x[5] = cmul(x[5], mk(c1, -s1));            // w^1
  x[9] = mk((x[9].x + x[9].y) * h, (x[9].y - x[9].x) * h);      // w^2
  x[13] = cmul(x[13], mk(s1, -c1));          // w^3
  x[6] = mk((x[6].x + x[6].y) * h, (x[6].y - x[6].x) * h);      // w^2
  x[10] = mul_neg_i(x[10]);                  // w^4
  x[14] = mk((x[14].y - x[14].x) * h, -(x[14].x + x[14].y) * h);   // w^6
  x[7] = cmul(x[7], mk(s1, -c1));            // w^3
  x[11] = mk((x[11].y - x[11].x) * h, -(x[11].x + x[11].y) * h);   // w^6
  x[15] = cmul(x[15], mk(-c1, s1));          // w^9
#pragma unroll
  for (int r = 0; r < 4; ++r) bfly4(x[4 * r], x[4 * r + 1], x[4 * r + 2], x[4 * r + 3]);
}
F4_HD constexpr int out16(int s) { return 4 * (s & 3) + (s >> 2); }   // register of dft16() that holds frequency s

// q-point DFT (q = 4, 8, 16) of x[0..q), natural order in y[0..q)
template <int Q>
F4_HD void dftq(c2 (&x)[Q]) {
  if constexpr (Q == 4) bfly4(x[0], x[1], x[2], x[3]);
  if constexpr (Q == 8) dft8(x);
  if constexpr (Q == 16) {
    dft16(x);
    c2 t[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) t[s] = x[out16(s)];
#pragma unroll
    for (int s = 0; s < 16; ++s) x[s] = t[s];
  }
}

// ------------------------------------------------------------------------------------------------ geometry
template <int N1, int T>
struct Geo {
  static constexpr int n = N1 * kN2;
  static constexpr int tile_points = 16 * T;
  static constexpr int w = T / 8;                            // n1 values per stage-A tile (a 64- or 128-byte run per n2)
  static constexpr int q = N1 / 16;                          // second-pass length of stage B (4, 8, 16)
  static constexpr int k2_per_tile = tile_points / (kColsPerGroup * N1);
  static constexpr int transforms = kColsPerGroup * k2_per_tile;          // stage-B transforms per tile
  static constexpr int S = N1 + 1;                           // exchange stride of one transform (odd: conflict-free reads)
  static constexpr int a_tiles_per_column = N1 / w;
  static constexpr int tiles_per_group = kColsPerGroup * a_tiles_per_column;   // of either stage, per 16-column group
  static constexpr int exchange_elems = (transforms * S > tile_points) ? transforms * S : tile_points;
  static_assert(k2_per_tile >= 1 && kN2 % k2_per_tile == 0 && kN2 / k2_per_tile == tiles_per_group, "tile geometry");
  static_assert(transforms * q == T, "stage B: one thread per (transform, j)");
};

// ------------------------------------------------------------------------------------------------ stage A
// Tile: one column, n1 in [n1_0, n1_0 + w), w = T / 8.  Thread: l = n1 - n1_0 = tid % w, j = tid / w (0..7).
// pass 1: x[n1 + N1 (j + 8 i)], i < 16  -> radix 16 over i -> y_s[j] * w_128^(j s)  -> exchange[s][j][l]
// pass 2: exchange[s][0..8)[l] for s = j, j + 8 -> radix 8 over j -> k2 = 16 m + s
//         -> * w_n^(n1 k2) -> Y[k2][n1]
// w128: the 128 values w_128^e.  t1[s * N1 + n1] = w_n^(n1 s) (s < 16) and t2[n1] = w_n^(16 n1): the four-step twiddle
// w_n^(n1 k2), k2 = 16 m + s, is t1 * t2^m; both tables are laid out so that the lanes of a warp (consecutive n1) read
// consecutive entries.
template <int N1, int T, class LoadF>
F4_HD void stage_a_pass1(int tid, int n1_0, const LoadF& load, const c2* w128, c2* exch) {
  constexpr int W = Geo<N1, T>::w;
  const int lane = tid % W, j = tid / W;
  c2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = load(n1_0 + lane + N1 * (j + 8 * i));
  dft16(x);
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    c2 v = x[out16(s)];
    if (s != 0) v = cmul(v, w128[j * s]);
    exch[(s * 8 + j) * W + lane] = v;
  }
}

template <int N1, int T>
F4_HD void stage_a_pass2(int tid, int n1_0, const c2* exch, const c2* t1, const c2* t2, c2* y_col) {
  constexpr int W = Geo<N1, T>::w;
  const int lane = tid % W, warp = tid / W;
  const int n1 = n1_0 + lane;
  const c2 step = t2[n1];                                    // w_n^(16 n1): from k2 to k2 + 16
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int s = warp + 8 * half;
    c2 x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = exch[(s * 8 + j) * W + lane];
    dft8(x);
    c2 w = t1[s * N1 + n1];                                  // w_n^(n1 s)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      y_col[(int64_t)(16 * m + s) * N1 + n1] = cmul(x[m], w);
      if (m != 7) w = cmul(w, step);
    }
  }
}

// ------------------------------------------------------------------------------------------------ stage B
// Tile: 16 adjacent columns x k2 in [k2_0, k2_0 + k2_per_tile); transform tau = k2l * 16 + c.
// pass 1: thread (tau, j), j < q:  Y[c][k2][j + q i], i < 16 -> radix 16 over i -> * w_N1^(j s) -> exch[tau S + s q + j]
//         lanes: q consecutive j of one transform are adjacent (q x 16 contiguous bytes per load); for q = 4 a
//         quarter-warp pairs transforms tau and tau + 4 so that its 8 shared-memory stores hit 8 distinct bank groups.
// pass 2: thread (c = tid & 15, k2l, u): exch[tau S + (u + q v) q + j], j < q -> radix q -> k1 = 16 m + u + q v.
//         lanes run over the 16 columns: every store instruction writes 128 contiguous bytes per half-warp.
template <int N1, int T>
F4_HD void stage_b_thread_pass1(int tid, int& tau, int& j) {
  constexpr int q = Geo<N1, T>::q;
  if constexpr (q == 4) {                                    // warp = 8 transforms: lane = j + 4 h + 8 k', tau = 8 warp + k' + 4 h
    const int lane = tid & 31, warp = tid >> 5;
    j = lane & 3;
    tau = 8 * warp + ((lane >> 3) & 3) + 4 * ((lane >> 2) & 1);
  } else {
    j = tid & (q - 1);
    tau = tid / q;
  }
}

template <int N1, int T, class LoadY>
F4_HD void stage_b_pass1(int tid, const LoadY& load_y, const c2* twb, c2* exch) {
  constexpr int q = Geo<N1, T>::q, S = Geo<N1, T>::S;
  int tau, j;
  stage_b_thread_pass1<N1, T>(tid, tau, j);
  c2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = load_y(tau, j + q * i);
  dft16(x);
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    c2 v = x[out16(s)];
    if (s != 0) v = cmul(v, twb[j * 17 + s]);                // w_N1^(j s); rows padded to 17 entries
    exch[tau * S + s * q + j] = v;
  }
}

// sink(c, k2l, k1, value): called for the 16 outputs of this thread
template <int N1, int T, class Sink>
F4_HD void stage_b_pass2(int tid, const c2* exch, Sink& sink) {
  constexpr int q = Geo<N1, T>::q, S = Geo<N1, T>::S;
  const int c = tid & 15, g = tid >> 4;
  const int k2l = g / q, u = g % q;
  const int tau = k2l * 16 + c;
#pragma unroll
  for (int v = 0; v < 16 / q; ++v) {
    const int s = u + q * v;
    c2 x[q];
#pragma unroll
    for (int j = 0; j < q; ++j) x[j] = exch[tau * S + s * q + j];
    dftq<q>(x);
#pragma unroll
    for (int m = 0; m < q; ++m) sink(c, k2l, 16 * m + s, x[m]);
  }
}

}  // namespace fft4
}  // namespace psa
