// Time-axis FFT of the projected columns, fused with the 1/n_t scale and the SED assembly.
//
// One CTA transforms one column (k, pol [, group]) entirely in shared memory and writes the spectrum
// straight into the result layout:
//   coherent   : complex64 out[f][k][pol]                   (reference: sed_calculator.py:296-311)
//   incoherent : float32  out[f][k] = sum_g sum_pol |S|^2    (reference: sed_calculator.py:313-327)
//
// Precision.  NumPy's complex64 FFT (the reference, sed_calculator.py:83) is accurate to a single
// float32 rounding (measured rms error 2e-8), far better than a transform whose every butterfly
// rounds to float32 (1e-7, scripts/fft_accuracy.py).  To stay at the reference's level the whole
// transform is carried in float64: float64 shared-memory storage, float64 butterflies and twiddles,
// and ONE rounding to float32 when the spectrum is stored (scripts/fft_accuracy.py: the rms error
// equals NumPy's at every length).  B200 runs FP64 at half the FP32 rate, so this costs little; the
// float32 <-> float64 conversions (16-lane XU pipe) are confined to the load and the final store.
//
// Where the time goes (ncu, 3000 columns of 16384 points, 0.76 ms): ~190 thread instructions per
// point, issue slots 34 % busy, FP64 pipe 20 %, XU 16 %, L1 data pipe 59 %; half of the stall
// samples sit in the load-time split (waiting on L2), a fifth in the scattered result store.  The
// kernel is latency-bound at 24 warps per SM (80 registers per thread for float64 butterflies).
// Tried and measured slower or equal in round 1 (scripts/fft_tune.py), and therefore not kept:
//   * float32 storage with float64 butterflies (same speed, 2x the rounding error);
//   * clusters of 2-8 ADJACENT columns exchanging spectra through distributed shared memory with an extra
//     pull-and-transpose pass in shared memory (0.80-0.86 ms vs 0.76 ms); the lean variant that reads the
//     peers' staged spectra directly while storing (fft_sed_cluster_kernel) does pay, 4 %, and is the default;
//   * the R CTAs of ONE column as a cluster, each loading 1/R of the samples and gathering the rest
//     from its peers' shared memory instead of re-reading them from L2 (0.90 ms vs 0.78 ms);
//   * twiddles factored into two 64-entry shared-memory tables instead of L2-resident per-pass tables
//     (0.775 ms vs 0.765 ms).
//
// Mixed-radix core (forward, decimation in frequency, in place, m = blk * 2^a 3^b 5^c points with
// blk = 16, 8 or 4, m <= 4096): covers the frame counts molecular-dynamics runs actually produce
// (10 000, 20 000, 50 000 ...) as well as the powers of two.
//   * shared-memory passes of radix 8 (one leading radix-2 or radix-4 pass), then 5, then 3, down to
//     contiguous blocks of blk points.  Consecutive threads take consecutive butterflies, so every
//     quarter-warp (the unit of a 128-bit shared access) touches consecutive elements: conflict-free.
//   * final stage: each thread pulls one contiguous block into registers, finishes it (radix 4 x 4,
//     8 or 4) and hands the results (digit-reversed frequency index) to a sink.  The array is
//     padded by one element per 16 (index p lives at p + p/16), which makes these per-thread
//     contiguous reads conflict-free as well.
//   * a CTA transforms at most 4096 points (68 KiB of float64 storage, three CTAs per SM).  Longer
//     columns are split by a radix-R decimation-in-frequency step applied while loading, giving R
//     independent sub-transforms (one CTA each) that produce the frequencies f = R f' + r; the R
//     CTAs of a column run side by side and share its samples through L2.
//   * twiddles: float64 tables, per pass and contiguous in the butterfly index (a warp reads short
//     contiguous runs), plus w_n^j for the load-time split.
//
// Frame counts the core cannot factor (odd, or with a prime factor above 5; the reference accepts any
// n_t through pocketfft) use Bluestein's chirp-z identity on top of it: with b_t = exp(i pi t^2 / n),
//   X_f = conj(b_f) * sum_t (x_t conj(b_t)) b_{f-t},
// i.e. one forward transform of length M >= 2n-1 (M = 2^s), a point-wise product with the
// precomputed spectrum of b, and one inverse transform; the chirped spectrum of each column makes
// one round trip through a scratch buffer.
#include <stdlib.h>

#include "common.cuh"

namespace psa {

#ifndef PSA_FFT_THREADS
#define PSA_FFT_THREADS 256
#endif
#ifndef PSA_FFT_UNROLL
#define PSA_FFT_UNROLL 2
#endif
#ifndef PSA_FFT_MIN_CTAS
#define PSA_FFT_MIN_CTAS 3
#endif
constexpr int kFftThreads = PSA_FFT_THREADS;
constexpr int kFftUnroll = PSA_FFT_UNROLL;
constexpr int kFftMinCtas = PSA_FFT_MIN_CTAS;
constexpr int64_t kMaxSmemPoints = 8192;       // complex128 points that fit one CTA (128 KiB + padding)
constexpr int64_t kDefaultSmemPoints = 4096;   // 68 KiB: three CTAs per SM
constexpr int64_t kMaxTransform = (int64_t)1 << 20;
constexpr int kBlk = 16;                    // points finished in registers per thread
constexpr int kMaxPasses = 6;

// ---------------------------------------------------------------------------------------------
// complex helpers (float64 arithmetic; sc = the shared-memory element)
// ---------------------------------------------------------------------------------------------
struct cd {
  double x, y;
};
__device__ __forceinline__ cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd widen(float2 a) { return mk((double)a.x, (double)a.y); }
__device__ __forceinline__ float2 narrow(cd a) { return make_float2((float)a.x, (float)a.y); }
using sc = double2;
__device__ __forceinline__ cd to_cd(sc a) { return mk(a.x, a.y); }
__device__ __forceinline__ sc to_sc(cd a) { return make_double2(a.x, a.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ cd cmul(cd a, double2 b) { return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ cd cmul_conj(cd a, cd b) { return mk(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y)); }   // a conj(b)
__device__ __forceinline__ cd cadd(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd mul_neg_i(cd a) { return mk(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ int phys(int p) { return p + (p >> 4); }

// thread-block cluster helpers
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float2 ld_peer(uint32_t smem_addr, uint32_t rank) {
  uint32_t remote;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote) : "memory");
  return v;
}

// forward DFTs in natural output order (no external twiddles)
__device__ __forceinline__ void bfly2(cd& a0, cd& a1) {
  cd t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}
__device__ __forceinline__ void bfly4(cd& a0, cd& a1, cd& a2, cd& a3) {
  cd t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
__device__ __forceinline__ void bfly8(cd (&a)[8]) {
  const double h = 0.70710678118654752440;
  cd u0 = cadd(a[0], a[4]), u1 = cadd(a[1], a[5]), u2 = cadd(a[2], a[6]), u3 = cadd(a[3], a[7]);
  cd v0 = csub(a[0], a[4]), d1 = csub(a[1], a[5]), d2 = csub(a[2], a[6]), d3 = csub(a[3], a[7]);
  cd v1 = mk((d1.x + d1.y) * h, (d1.y - d1.x) * h);        // * w8   = (1 - i)/sqrt2
  cd v2 = mul_neg_i(d2);                                     // * w8^2 = -i
  cd v3 = mk((d3.y - d3.x) * h, -(d3.x + d3.y) * h);        // * w8^3 = (-1 - i)/sqrt2
  bfly4(u0, u1, u2, u3);                                     // even outputs 0,2,4,6
  bfly4(v0, v1, v2, v3);                                     // odd outputs 1,3,5,7
  a[0] = u0; a[2] = u1; a[4] = u2; a[6] = u3;
  a[1] = v0; a[3] = v1; a[5] = v2; a[7] = v3;
}

__device__ __forceinline__ void bfly3(cd& a0, cd& a1, cd& a2) {           // w3 = -1/2 - i sin(2 pi/3)
  const double sn = 0.86602540378443864676;
  const cd t = cadd(a1, a2), d = csub(a1, a2);
  const cd m = mk(a0.x - 0.5 * t.x, a0.y - 0.5 * t.y), js = mk(sn * d.y, -sn * d.x);   // js = -i sn d
  a0 = cadd(a0, t);
  a1 = cadd(m, js);
  a2 = csub(m, js);
}
__device__ __forceinline__ void bfly5(cd (&a)[5]) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2 pi/5), cos(4 pi/5)
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2 pi/5), sin(4 pi/5)
  const cd t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), d1 = csub(a[1], a[4]), d2 = csub(a[2], a[3]);
  const cd m1 = mk(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
  const cd m2 = mk(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
  const cd n1 = mk(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
  const cd n2 = mk(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
  a[0] = cadd(a[0], cadd(t1, t2));
  a[1] = mk(m1.x + n1.y, m1.y - n1.x);                   // m1 - i n1
  a[4] = mk(m1.x - n1.y, m1.y + n1.x);                   // m1 + i n1
  a[2] = mk(m2.x + n2.y, m2.y - n2.x);
  a[3] = mk(m2.x - n2.y, m2.y + n2.x);
}

// ---------------------------------------------------------------------------------------------
// the mixed-radix core (lengths blk * 2^a 3^b 5^c, blk = 16, 8 or 4)
// ---------------------------------------------------------------------------------------------
// Pass schedule for an m-point sub-transform and the layout of its per-pass twiddle tables: the passes
// (decimation in frequency, span L[p], radix r[p] in {2, 3, 4, 5, 8}) take the span from m down to
// contiguous blocks of `blk` points (16, else 8, else 4 - the largest that divides m), which one thread
// finishes in registers.  Pass p has q = L/r butterflies per block and q table entries w_L^j (j < q)
// starting at tw_off[p]; the higher powers w_L^{2j} .. w_L^{(r-1)j} are formed in float64 registers
// instead of being loaded.  n_pass < 0: m has a prime factor other than 2, 3, 5 (or is not a multiple of 4).
struct PassPlan {
  int n_pass, total, blk;
  int L[kMaxPasses], r[kMaxPasses], tw_off[kMaxPasses];
  int sh[kMaxPasses];       // log2 of (L / r) / blk, the block-index weight of pass p's digit, or -1 if not a power of two
};
__host__ __device__ inline void plan_add(PassPlan& pp, int& L, int r) {
  if (pp.n_pass < 0 || pp.n_pass >= kMaxPasses) { pp.n_pass = -1; return; }
  pp.L[pp.n_pass] = L;
  pp.r[pp.n_pass] = r;
  pp.tw_off[pp.n_pass] = pp.total;
  pp.total += L / r;
  L /= r;
  const int stride = L / pp.blk;
  int sh = -1;
  if ((stride & (stride - 1)) == 0)
    for (sh = 0; (1 << sh) < stride; ++sh) {}
  pp.sh[pp.n_pass] = sh;
  ++pp.n_pass;
}
__host__ __device__ inline PassPlan make_passes(int m) {
  PassPlan pp;
  pp.n_pass = 0;
  pp.total = 0;
  pp.blk = m % 16 == 0 ? 16 : (m % 8 == 0 ? 8 : (m % 4 == 0 ? 4 : 0));
  if (m <= 0 || pp.blk == 0) { pp.n_pass = -1; return pp; }
  int x = m / pp.blk, L = m, twos = 0;
  while (x % 2 == 0) { x /= 2; ++twos; }
  if (twos % 3) plan_add(pp, L, 1 << (twos % 3));        // one leading radix-2 or radix-4 pass, then radix 8
  for (int i = 0; i < twos / 3; ++i) plan_add(pp, L, 8);
  while (x % 5 == 0) { x /= 5; plan_add(pp, L, 5); }
  while (x % 3 == 0) { x /= 3; plan_add(pp, L, 3); }
  if (x != 1) pp.n_pass = -1;
  return pp;
}

// Geometry of one launch: transform length n_fft = m * R.
struct FftGeom {
  int m, R, n_fft;
  const double2* tw;        // w_n^j, j < n_fft (load-time split of long transforms)
  const double2* pass_tw;   // per-pass tables
  PassPlan pp;
};

// kPow2: q is a power of two and a multiple of 16 (every pass of a power-of-two length with 16-point blocks):
// mask instead of division, and a constant leg stride in the padded storage.
template <int RADIX, bool kPow2>
__device__ __forceinline__ void smem_butterfly(sc* __restrict__ s, int b, int q, const double2* __restrict__ tab) {
  const int j = kPow2 ? (b & (q - 1)) : b % q;
  const int p0 = (b - j) * RADIX + j;                         // (b / q) * L + j; legs are q apart
  const double2 t = __ldg(tab + j);
  const int pp0 = phys(p0), qs = q + (q >> 4);
  cd a[RADIX];
#pragma unroll
  for (int i = 0; i < RADIX; ++i) a[i] = to_cd(s[kPow2 ? pp0 + i * qs : phys(p0 + i * q)]);
  const cd w1 = mk(t.x, t.y);
  if (RADIX == 2) {
    bfly2(a[0], a[1]);
    a[1] = cmul(a[1], w1);
  }
  if (RADIX == 3) {
    bfly3(a[0], a[1], a[2]);
    a[1] = cmul(a[1], w1);
    a[2] = cmul(a[2], cmul(w1, w1));
  }
  if (RADIX == 4) {
    bfly4(a[0], a[1], a[2], a[3]);
    const cd w2 = cmul(w1, w1);
    a[1] = cmul(a[1], w1);
    a[2] = cmul(a[2], w2);
    a[3] = cmul(a[3], cmul(w2, w1));
  }
  if (RADIX == 5) {
    bfly5(reinterpret_cast<cd(&)[5]>(a));
    const cd w2 = cmul(w1, w1);
    a[1] = cmul(a[1], w1);
    a[2] = cmul(a[2], w2);
    a[3] = cmul(a[3], cmul(w2, w1));
    a[4] = cmul(a[4], cmul(w2, w2));
  }
  if (RADIX == 8) {
    bfly8(reinterpret_cast<cd(&)[8]>(a));
    const cd w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
    a[1] = cmul(a[1], w1);
    a[2] = cmul(a[2], w2);
    a[3] = cmul(a[3], w3);
    a[4] = cmul(a[4], w4);
    a[5] = cmul(a[5], cmul(w4, w1));
    a[6] = cmul(a[6], cmul(w3, w3));
    a[7] = cmul(a[7], cmul(w4, w3));
  }
#pragma unroll
  for (int i = 0; i < RADIX; ++i) s[kPow2 ? pp0 + i * qs : phys(p0 + i * q)] = to_sc(a[i]);
}

// One body per (radix, index mode) in the binary: the pass schedule is unrolled over six slots and would otherwise
// inline sixty butterfly loops (the fused kernel grew to 480 KB of SASS and stalled on instruction fetch).
template <int RADIX, bool kPow2>
__device__ __noinline__ void smem_pass_impl(sc* __restrict__ s, int m, int q, const double2* __restrict__ tab) {
  const int n_bfly = m / RADIX, step = blockDim.x;
  if (n_bfly % step == 0) {                            // uniform trip count: no exit test between iterations
    const int per_thread = n_bfly / step;
#pragma unroll kFftUnroll
    for (int i = 0; i < per_thread; ++i) smem_butterfly<RADIX, kPow2>(s, threadIdx.x + i * step, q, tab);
  } else {
    for (int b = threadIdx.x; b < n_bfly; b += step) smem_butterfly<RADIX, kPow2>(s, b, q, tab);
  }
  __syncthreads();
}
template <int RADIX>
__device__ __forceinline__ void smem_pass(sc* __restrict__ s, int m, int L, const double2* __restrict__ tab) {
  const int q = L / RADIX;
  if (RADIX != 3 && RADIX != 5 && (q & (q - 1)) == 0 && (q & 15) == 0) smem_pass_impl<RADIX, true>(s, m, q, tab);
  else smem_pass_impl<RADIX, false>(s, m, q, tab);
}

__device__ void fft_smem_passes(sc* __restrict__ s, const FftGeom& g) {
#pragma unroll                                           // static indices keep the plan in the constant bank
  for (int p = 0; p < kMaxPasses; ++p) {
    if (p < g.pp.n_pass) {
      const double2* tab = g.pass_tw + g.pp.tw_off[p];
      if (g.pp.r[p] == 8) smem_pass<8>(s, g.m, g.pp.L[p], tab);
      else if (g.pp.r[p] == 4) smem_pass<4>(s, g.m, g.pp.L[p], tab);
      else if (g.pp.r[p] == 2) smem_pass<2>(s, g.m, g.pp.L[p], tab);
      else if (g.pp.r[p] == 5) smem_pass<5>(s, g.m, g.pp.L[p], tab);
      else smem_pass<3>(s, g.m, g.pp.L[p], tab);
    }
  }
}

// w_16^k = exp(-2 pi i k / 16), k = 0..9
__constant__ double2 kW16[10] = {
    {1.0, 0.0},
    {0.9238795325112867, -0.3826834323650898},
    {0.7071067811865476, -0.7071067811865476},
    {0.3826834323650898, -0.9238795325112867},
    {0.0, -1.0},
    {-0.3826834323650898, -0.9238795325112867},
    {-0.7071067811865476, -0.7071067811865476},
    {-0.9238795325112867, -0.3826834323650898},
    {-1.0, 0.0},
    {-0.9238795325112867, 0.3826834323650898}};

// Finish one contiguous 16-point block held in registers: radix 4 (L=16), radix 4 (L=4).
// Register e then holds the block-local frequency (e >> 2) + 4 * (e & 3).
__device__ __forceinline__ void fft16_registers(cd (&x)[kBlk]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bfly4(x[j], x[j + 4], x[j + 8], x[j + 12]);
    if (j != 0) {
      x[j + 4] = cmul(x[j + 4], kW16[j]);
      x[j + 8] = cmul(x[j + 8], kW16[2 * j]);
      x[j + 12] = cmul(x[j + 12], kW16[3 * j]);
    }
  }
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) bfly4(x[4 * blk], x[4 * blk + 1], x[4 * blk + 2], x[4 * blk + 3]);
}

// Frequency (within the m-point sub-transform) of element 0 of block b; register e of the finished block
// adds (m / blk) * rev(e).  The block index is a mixed-radix number, most significant digit = first pass;
// that digit is the LEAST significant digit of the frequency.
__device__ __forceinline__ int block_base_frequency(int b, const FftGeom& g) {
  int stride = g.m / g.pp.blk, f = 0, weight = 1;
#pragma unroll
  for (int p = 0; p < kMaxPasses; ++p) {
    if (p < g.pp.n_pass) {
      stride /= g.pp.r[p];                             // uniform, from the constant bank
      const int d = g.pp.sh[p] >= 0 ? (b >> g.pp.sh[p]) : b / stride;
      b -= d * stride;
      f += d * weight;
      weight *= g.pp.r[p];
    }
  }
  return f;
}

// Finish one contiguous block held in registers; rev(e) = block-local frequency index of register e.
template <int BLK>
__device__ __forceinline__ void finish_block(cd (&x)[BLK]) {
  if constexpr (BLK == 16) fft16_registers(x);
  if constexpr (BLK == 8) bfly8(x);
  if constexpr (BLK == 4) bfly4(x[0], x[1], x[2], x[3]);
}
template <int BLK>
__device__ __forceinline__ constexpr int block_rev(int e) { return BLK == 16 ? (e >> 2) + 4 * (e & 3) : e; }

// output r of a 4-point forward DFT: sum_j x_j (-i)^{jr}
__device__ __forceinline__ cd dif4(cd x0, cd x1, cd x2, cd x3, int r) {
  if (r == 0) return cadd(cadd(x0, x2), cadd(x1, x3));
  if (r == 2) return csub(cadd(x0, x2), cadd(x1, x3));
  const cd d = mul_neg_i(csub(x1, x3));               // -i (x1 - x3)
  return r == 1 ? cadd(csub(x0, x2), d) : csub(csub(x0, x2), d);
}

// Radix-RR split for small odd-ish RR (frame counts like 20 000 = 5 x 4000): the RR twiddles w_RR^{jr} live in
// registers and two outputs (2 RR loads) are in flight per thread.
template <int RR, class Fetch>
__device__ void load_split_small(sc* __restrict__ s, const Fetch& fetch, const FftGeom& g, int r) {
  double2 w[RR];
#pragma unroll
  for (int j = 0; j < RR; ++j) w[j] = __ldg(g.tw + ((j * r) % RR) * g.m);
  auto residue = [&](const cd (&x)[RR], int t) -> sc {
    cd acc = x[0];
#pragma unroll
    for (int j = 1; j < RR; ++j) acc = cadd(acc, r ? cmul(x[j], w[j]) : x[j]);
    return to_sc(r ? cmul(acc, __ldg(g.tw + (int64_t)t * r)) : acc);
  };
  const int step = blockDim.x;
  int t = threadIdx.x;
  for (; t + step < g.m; t += 2 * step) {
    cd x0[RR], x1[RR];
#pragma unroll
    for (int j = 0; j < RR; ++j) {
      x0[j] = fetch(t + j * g.m);
      x1[j] = fetch(t + step + j * g.m);
    }
    s[phys(t)] = residue(x0, t);
    s[phys(t + step)] = residue(x1, t + step);
  }
  for (; t < g.m; t += step) {
    cd x0[RR];
#pragma unroll
    for (int j = 0; j < RR; ++j) x0[j] = fetch(t + j * g.m);
    s[phys(t)] = residue(x0, t);
  }
}

// Fill shared memory with sub-sequence r of the radix-R split of fetch(0..n_fft) (R == 1: plain copy).
template <class Fetch>
__device__ void load_column(sc* __restrict__ s, const Fetch& fetch, const FftGeom& g, int r) {
  const int step = blockDim.x;
  if (g.R == 1) {
    // Latency-bound: every thread first issues the loads of 16 elements (32 loads in flight), then parks
    // them in shared memory.  A plain strided loop keeps an exit test between unrolled iterations, which
    // serialises load -> store (ncu: 1/3 of all stall samples sat on that store).
    int t = threadIdx.x;
    for (; t + 15 * step < g.m; t += 16 * step) {
      sc v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = to_sc(fetch(t + u * step));
#pragma unroll
      for (int u = 0; u < 16; ++u) s[phys(t + u * step)] = v[u];
    }
    for (; t < g.m; t += step) s[phys(t)] = to_sc(fetch(t));
    return;
  }
  if (g.R == 2) {                                     // x[t] +- x[t + m], batched like above
    int t = threadIdx.x;
    for (; t + 7 * step < g.m; t += 8 * step) {
      cd a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[u] = fetch(t + u * step);
        b[u] = fetch(t + u * step + g.m);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        s[phys(t + u * step)] = to_sc(r ? cmul(csub(a[u], b[u]), __ldg(g.tw + t + u * step)) : cadd(a[u], b[u]));
    }
    for (; t < g.m; t += step) {
      cd a = fetch(t), b = fetch(t + g.m);
      s[phys(t)] = to_sc(r ? cmul(csub(a, b), __ldg(g.tw + t)) : cadd(a, b));
    }
    return;
  }
  if (g.R == 4) {                                     // sum_j x[t + j m] (-i)^{jr}, times w_n^{tr}
    int t = threadIdx.x;
    for (; t + 3 * step < g.m; t += 4 * step) {
      cd x[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[u][j] = fetch(t + u * step + j * g.m);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const cd y = dif4(x[u][0], x[u][1], x[u][2], x[u][3], r);
        s[phys(t + u * step)] = to_sc(r ? cmul(y, __ldg(g.tw + (t + u * step) * r)) : y);
      }
    }
    for (; t < g.m; t += step) {
      const cd y = dif4(fetch(t), fetch(t + g.m), fetch(t + 2 * g.m), fetch(t + 3 * g.m), r);
      s[phys(t)] = to_sc(r ? cmul(y, __ldg(g.tw + t * r)) : y);
    }
    return;
  }
  if (g.R == 8) {                                     // even and odd j are radix-4 sums: E + w_8^r O
    const double2 w8r = __ldg(g.tw + r * (g.n_fft >> 3));
    int t = threadIdx.x;
    for (; t + step < g.m; t += 2 * step) {
      cd x[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = fetch(t + u * step + j * g.m);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const cd y = cadd(dif4(x[u][0], x[u][2], x[u][4], x[u][6], r & 3),
                          cmul(dif4(x[u][1], x[u][3], x[u][5], x[u][7], r & 3), w8r));
        s[phys(t + u * step)] = to_sc(r ? cmul(y, __ldg(g.tw + (t + u * step) * r)) : y);
      }
    }
    for (; t < g.m; t += step) {
      cd x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fetch(t + j * g.m);
      const cd y = cadd(dif4(x[0], x[2], x[4], x[6], r & 3), cmul(dif4(x[1], x[3], x[5], x[7], r & 3), w8r));
      s[phys(t)] = to_sc(r ? cmul(y, __ldg(g.tw + t * r)) : y);
    }
    return;
  }
  if (g.R == 3) return load_split_small<3>(s, fetch, g, r);
  if (g.R == 5) return load_split_small<5>(s, fetch, g, r);
  if (g.R == 6) return load_split_small<6>(s, fetch, g, r);
  if (g.R == 10) return load_split_small<10>(s, fetch, g, r);
  for (int t = threadIdx.x; t < g.m; t += step) {     // general R: sum_j x[t + j m] w_R^{jr}, times w_n^{tr}
    cd acc = mk(0.0, 0.0);
    for (int j = 0; j < g.R; ++j) {
      cd x = fetch(t + j * g.m);
      const int wi = ((j * r) % g.R) * g.m;                            // w_R^{jr} = w_n^{(jr mod R) m}
      acc = cadd(acc, wi ? cmul(x, __ldg(g.tw + wi)) : x);
    }
    s[phys(t)] = to_sc(r ? cmul(acc, __ldg(g.tw + (int64_t)t * r)) : acc);   // w_n^{tr}, t r < n
  }
}

// Transform the column in shared memory and feed every (slot, frequency, value) to the sink.
// slot = padded in-place position, owned by the same thread on every call with the same geometry.
template <int BLK, class Sink>
__device__ __forceinline__ void finish_and_emit(sc* __restrict__ s, const FftGeom& g, int r, Sink& sink) {
  const int n_blocks = g.m / BLK;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
    cd x[BLK];
#pragma unroll
    for (int e = 0; e < BLK; ++e) x[e] = to_cd(s[phys(b * BLK + e)]);
    finish_block<BLK>(x);
    const int f0 = block_base_frequency(b, g);
#pragma unroll
    for (int e = 0; e < BLK; ++e) sink(phys(b * BLK + e), (f0 + n_blocks * block_rev<BLK>(e)) * g.R + r, x[e]);
  }
}
template <class Sink>
__device__ void transform_and_emit(sc* __restrict__ s, const FftGeom& g, int r, Sink& sink) {
  fft_smem_passes(s, g);
  if (g.pp.blk == 16) finish_and_emit<16>(s, g, r, sink);
  else if (g.pp.blk == 8) finish_and_emit<8>(s, g, r, sink);
  else finish_and_emit<4>(s, g, r, sink);
}

// ---------------------------------------------------------------------------------------------
// fetchers (return the element widened to float64)
// ---------------------------------------------------------------------------------------------
struct FetchPlanar {            // rows of P: Re and Im as separate float rows; optional taper (exact in float64)
  const float* re;
  const float* im;
  const float* win;
  __device__ cd operator()(int t) const {
    cd v = mk((double)__ldg(re + t), (double)__ldg(im + t));
    if (win != nullptr) {
      const double w = (double)__ldg(win + t);
      v.x *= w;
      v.y *= w;
    }
    return v;
  }
};
struct FetchChirped {           // x_t * conj(b_t) for t < n, zero padding up to the transform length
  const float* re;
  const float* im;
  const float* win;
  const double2* chirp;
  int n;
  __device__ cd operator()(int t) const {
    if (t >= n) return mk(0.0, 0.0);
    const double2 b = __ldg(chirp + t);
    cd v = mk((double)__ldg(re + t), (double)__ldg(im + t));
    if (win != nullptr) {
      const double w = (double)__ldg(win + t);
      v.x *= w;
      v.y *= w;
    }
    return cmul_conj(v, mk(b.x, b.y));
  }
};
struct FetchConj {              // conj of an interleaved complex128 column (inverse transform by conjugation)
  const double2* src;
  __device__ cd operator()(int t) const {
    double2 v = __ldg(src + t);
    return mk(v.x, -v.y);
  }
};
struct FetchComplexD {
  const double2* src;
  __device__ cd operator()(int t) const {
    double2 v = __ldg(src + t);
    return mk(v.x, v.y);
  }
};

// ---------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------
// value -> S: scale (power-of-two path) or the Bluestein unchirp, in float64, rounded once to float32.
struct Unchirp {                // Bluestein: X_f = conj(b_f) * conj(y_f) / M, then / n   (f < n only)
  const double2* chirp;
  int n;
  double inv_m, n_d;
  __device__ bool operator()(int f, cd y, float2& out) const {
    if (f >= n) return false;
    const double2 b = __ldg(chirp + f);
    cd X = cmul_conj(mk(y.x * inv_m, -y.y * inv_m), mk(b.x, b.y));
    out = make_float2((float)(X.x / n_d), (float)(X.y / n_d));          // divide by n_t like the reference
    return true;
  }
};
struct ScaleOnly {              // power of two: multiply by the exact reciprocal 1 / n_t
  double inv_n;
  __device__ bool operator()(int, cd y, float2& out) const {
    out = make_float2((float)(y.x * inv_n), (float)(y.y * inv_n));
    return true;
  }
};

template <class Post>
struct SinkCoherent {
  float2* o;                    // already offset to (k, pol)
  int64_t fstride;
  Post post;
  __device__ void operator()(int, int f, cd v) const {
    float2 S;
    if (post(f, v, S)) o[(int64_t)f * fstride] = S;
  }
};
template <class Post>
struct SinkAccumulate {         // s_acc[slot] += |S|^2 of the float32 S; the slot is private to the calling thread
  float* s_acc;
  Post post;
  __device__ void operator()(int slot, int f, cd v) const {
    float2 S;
    if (post(f, v, S)) s_acc[slot] += S.x * S.x + S.y * S.y;
  }
};
struct SinkTimesSpectrum {      // natural-order float64 store of value * bhat[f] (Bluestein forward leg; the chirped
  double2* dst;                 // spectrum stays in float64 between the two legs - one float32 rounding in total)
  const double2* bhat;
  __device__ void operator()(int, int f, cd v) const {
    const double2 b = __ldg(bhat + f);
    const cd y = cmul(v, mk(b.x, b.y));
    dst[f] = make_double2(y.x, y.y);
  }
};
struct SinkStoreD {             // natural-order float64 store (plan construction)
  double2* dst;
  __device__ void operator()(int, int f, cd v) const { dst[f] = make_double2(v.x, v.y); }
};

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
struct SedArgs {
  const float* P;
  int n_groups;
  int64_t group_stride, ldp;
  int n_t;
  void* out;
  int64_t n_k_total, k_offset;
  double inv_n_t;           // 1 / n_t from the host: a float64 division per CTA showed up as 5 % of the stall samples
  const float* window;      // optional taper w[t] (NULL = rectangular, the reference)
};

__device__ __forceinline__ void column_rows(const SedArgs& a, int g, int k, int pol, const float*& re, const float*& im) {
  const float* base = a.P + (int64_t)g * a.group_stride;
  re = base + ((int64_t)(2 * k) * 3 + pol) * a.ldp;
  im = base + ((int64_t)(2 * k + 1) * 3 + pol) * a.ldp;
}

// write the intensities accumulated per slot (same thread -> slot mapping as transform_and_emit)
template <int BLK>
__device__ __forceinline__ void flush_blocks(const float* __restrict__ s_acc, float* __restrict__ o, int64_t fstride,
                                             int n_valid, const FftGeom& g, int r) {
  const int n_blocks = g.m / BLK;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
    const int f0 = block_base_frequency(b, g);
#pragma unroll
    for (int e = 0; e < BLK; ++e) {
      const int f = (f0 + n_blocks * block_rev<BLK>(e)) * g.R + r;
      if (f < n_valid) o[(int64_t)f * fstride] = s_acc[phys(b * BLK + e)];
    }
  }
}
__device__ void flush_accumulator(const float* __restrict__ s_acc, float* __restrict__ o, int64_t fstride, int n_valid,
                                  const FftGeom& g, int r) {
  if (g.pp.blk == 16) flush_blocks<16>(s_acc, o, fstride, n_valid, g, r);
  else if (g.pp.blk == 8) flush_blocks<8>(s_acc, o, fstride, n_valid, g, r);
  else flush_blocks<4>(s_acc, o, fstride, n_valid, g, r);
}

// Direct lengths (n_t = R * m with a mixed-radix m): P -> result in one kernel.
template <int kMode>
__global__ void __launch_bounds__(kFftThreads, kFftMinCtas) fft_sed_kernel(SedArgs a, FftGeom g) {
  extern __shared__ sc s_data[];
  const ScaleOnly post{a.inv_n_t};
  const int r = blockIdx.x % g.R;
  if (kMode == PSA_MODE_COHERENT) {          // block -> (k, pol, r)
    const int pol = (blockIdx.x / g.R) % 3, k = blockIdx.x / (3 * g.R);
    FetchPlanar fetch;
    fetch.win = a.window;
    column_rows(a, 0, k, pol, fetch.re, fetch.im);
    load_column(s_data, fetch, g, r);
    __syncthreads();
    SinkCoherent<ScaleOnly> sink{reinterpret_cast<float2*>(a.out) + (a.k_offset + k) * 3 + pol, a.n_k_total * 3, post};
    transform_and_emit(s_data, g, r, sink);
  } else {                                   // block -> (k, r); loop over groups and polarisations
    const int k = blockIdx.x / g.R;
    const int padded = g.m + (g.m >> 4);
    float* s_acc = reinterpret_cast<float*>(s_data + padded);
    for (int i = threadIdx.x; i < padded; i += blockDim.x) s_acc[i] = 0.f;
    SinkAccumulate<ScaleOnly> acc{s_acc, post};
    for (int grp = 0; grp < a.n_groups; ++grp)
      for (int pol = 0; pol < 3; ++pol) {
        FetchPlanar fetch;
        fetch.win = a.window;
        column_rows(a, grp, k, pol, fetch.re, fetch.im);
        __syncthreads();
        load_column(s_data, fetch, g, r);
        __syncthreads();
        transform_and_emit(s_data, g, r, acc);
      }
    // every slot was accumulated and is flushed by the same thread: no barrier needed
    flush_accumulator(s_acc, reinterpret_cast<float*>(a.out) + a.k_offset + k, a.n_k_total, g.n_fft, g, r);
  }
}

// Coherent result with wider stores (default C = 4; PSA_FFT_CLUSTER=1|2|4): the CTAs of C adjacent columns (same
// residue r) form a cluster.  Each leaves its finished float32 spectrum in its own shared memory in NATURAL
// order (aliasing the column storage: needs one finished block per thread); after a cluster barrier CTA c
// writes frequencies [c m/C, (c+1) m/C) of all C columns - a warp reads 128-byte runs from each peer through
// distributed shared memory and stores C x 8-byte runs instead of single 8-byte pieces.
template <int BLK>
__device__ __forceinline__ void finish_and_stage(sc* __restrict__ s, const FftGeom& g, int r, const ScaleOnly& post) {
  const int n_blocks = g.m / BLK, b = threadIdx.x;
  const bool have = b < n_blocks;
  cd x[BLK];
  if (have) {
#pragma unroll
    for (int e = 0; e < BLK; ++e) x[e] = to_cd(s[phys(b * BLK + e)]);
    finish_block<BLK>(x);
  }
  __syncthreads();                                     // every block is in registers: the storage can be reused
  if (have) {
    float2* stage = reinterpret_cast<float2*>(s);
    const int f0 = block_base_frequency(b, g);
#pragma unroll
    for (int e = 0; e < BLK; ++e) {
      const int fp = f0 + n_blocks * block_rev<BLK>(e);
      float2 S = make_float2(0.f, 0.f);
      post(fp * g.R + r, x[e], S);
      stage[fp] = S;
    }
  }
}

__global__ void __launch_bounds__(kFftThreads, kFftMinCtas) fft_sed_cluster_kernel(SedArgs a, FftGeom g, int n_cols) {
  extern __shared__ sc s_data[];
  const ScaleOnly post{a.inv_n_t};
  const int C = (int)cluster_size(), rank = (int)cluster_rank();
  const int r = (blockIdx.x / C) % g.R;
  const int col0 = (blockIdx.x / (C * g.R)) * C, col = col0 + rank;
  if (col < n_cols) {
    FetchPlanar fetch;
    fetch.win = a.window;
    column_rows(a, 0, col / 3, col % 3, fetch.re, fetch.im);
    load_column(s_data, fetch, g, r);
    __syncthreads();
    fft_smem_passes(s_data, g);
    if (g.pp.blk == 16) finish_and_stage<16>(s_data, g, r, post);
    else if (g.pp.blk == 8) finish_and_stage<8>(s_data, g, r, post);
    else finish_and_stage<4>(s_data, g, r, post);
  }
  cluster_barrier();                                   // every column of the cluster is staged
  const int per = g.m / C, shift = 31 - __clz(C), live = min(C, n_cols - col0);
  const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(s_data);
  float2* o = reinterpret_cast<float2*>(a.out) + a.k_offset * 3 + col0;
  const int64_t fstride = a.n_k_total * 3;
  constexpr int kBatch = 8;
  for (int i0 = threadIdx.x; i0 < per * C; i0 += kBatch * blockDim.x) {
    float2 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * blockDim.x, c = i & (C - 1), fp = rank * per + (i >> shift);
      if (i < per * C && c < live) v[u] = ld_peer(stage_addr + (uint32_t)fp * 8u, (uint32_t)c);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * blockDim.x, c = i & (C - 1), fp = rank * per + (i >> shift);
      if (i < per * C && c < live) o[(int64_t)(fp * g.R + r) * fstride + c] = v[u];
    }
  }
  cluster_barrier();                                   // nobody leaves while a peer may still read its stage
}

// Bluestein leg 1: column -> chirp, zero-pad, forward transform, times the chirp spectrum -> scratch.
// block -> (column, r), column = (group, k, pol) flattened.
__global__ void __launch_bounds__(kFftThreads, kFftMinCtas) bluestein_forward_kernel(SedArgs a, FftGeom g, int n_k,
                                                                        const double2* __restrict__ chirp,
                                                                        const double2* __restrict__ bhat,
                                                                        double2* __restrict__ scratch) {
  extern __shared__ sc s_data[];
  const int r = blockIdx.x % g.R;
  const int col = blockIdx.x / g.R;
  const int pol = col % 3, k = (col / 3) % n_k, grp = col / (3 * n_k);
  FetchChirped fetch;
  fetch.win = a.window;
  column_rows(a, grp, k, pol, fetch.re, fetch.im);
  fetch.chirp = chirp;
  fetch.n = a.n_t;
  load_column(s_data, fetch, g, r);
  __syncthreads();
  SinkTimesSpectrum sink{scratch + (int64_t)col * g.n_fft, bhat};
  transform_and_emit(s_data, g, r, sink);
}

// Bluestein leg 2: scratch -> inverse transform (by conjugation), unchirp, / n_t, SED assembly.
template <int kMode>
__global__ void __launch_bounds__(kFftThreads, kFftMinCtas) bluestein_inverse_kernel(SedArgs a, FftGeom g, int n_k,
                                                                        const double2* __restrict__ chirp,
                                                                        const double2* __restrict__ scratch) {
  extern __shared__ sc s_data[];
  const Unchirp post{chirp, a.n_t, 1.0 / (double)g.n_fft, (double)a.n_t};
  const int r = blockIdx.x % g.R;
  if (kMode == PSA_MODE_COHERENT) {
    const int col = blockIdx.x / g.R;          // (k, pol), single group
    const int pol = col % 3, k = col / 3;
    FetchConj fetch{scratch + (int64_t)col * g.n_fft};
    load_column(s_data, fetch, g, r);
    __syncthreads();
    SinkCoherent<Unchirp> sink{reinterpret_cast<float2*>(a.out) + (a.k_offset + k) * 3 + pol, a.n_k_total * 3, post};
    transform_and_emit(s_data, g, r, sink);
  } else {
    const int k = blockIdx.x / g.R;
    const int padded = g.m + (g.m >> 4);
    float* s_acc = reinterpret_cast<float*>(s_data + padded);
    for (int i = threadIdx.x; i < padded; i += blockDim.x) s_acc[i] = 0.f;
    SinkAccumulate<Unchirp> acc{s_acc, post};
    for (int grp = 0; grp < a.n_groups; ++grp)
      for (int pol = 0; pol < 3; ++pol) {
        const int col = (grp * n_k + k) * 3 + pol;
        FetchConj fetch{scratch + (int64_t)col * g.n_fft};
        __syncthreads();
        load_column(s_data, fetch, g, r);
        __syncthreads();
        transform_and_emit(s_data, g, r, acc);
      }
    flush_accumulator(s_acc, reinterpret_cast<float*>(a.out) + a.k_offset + k, a.n_k_total, a.n_t, g, r);
  }
}

// float64-in, float64-out forward transform of one column, natural order (the chirp spectrum of a plan)
__global__ void __launch_bounds__(kFftThreads, kFftMinCtas) fft_c2c_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                              FftGeom g) {
  extern __shared__ sc s_data[];
  const int r = blockIdx.x % g.R;
  FetchComplexD fetch{src};
  load_column(s_data, fetch, g, r);
  __syncthreads();
  SinkStoreD sink{dst};
  transform_and_emit(s_data, g, r, sink);
}

// ---------------------------------------------------------------------------------------------
// plan: float64 tables in one caller-owned buffer (double2 entries)
//   power of two : [ tw (n_t) | pass tables ]
//   otherwise    : [ tw (M)   | pass tables | chirp (n_t) | bhat (M) | work (M) ]
// ---------------------------------------------------------------------------------------------
__global__ void twiddle_kernel(int64_t n, double2* __restrict__ tw) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s, c;
  sincospi(2.0 * (double)j / (double)n, &s, &c);
  tw[j] = make_double2(c, -s);
}

__global__ void pass_table_kernel(int L, int radix, double2* __restrict__ table) {   // table already offset to the pass
  const int q = L / radix;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= q) return;
  double s, c;
  sincospi(2.0 * (double)j / (double)L, &s, &c);
  table[j] = make_double2(c, -s);
}

// chirp[t] = exp(+i pi t^2 / n) with t^2 reduced mod 2n in integers; padded[] = the circular kernel of length M
__global__ void chirp_kernel(int64_t n, int64_t M, double2* __restrict__ chirp, double2* __restrict__ padded) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M) return;
  double2 val = make_double2(0.0, 0.0);
  int64_t src = t < n ? t : (M - t < n ? M - t : -1);        // b_{-t} = b_t wraps to M - t
  if (src >= 0) {
    int64_t q = (src * src) % (2 * n);
    double s, c;
    sincospi((double)q / (double)n, &s, &c);
    val = make_double2(c, s);
    if (t < n) chirp[t] = val;
  }
  padded[t] = val;
}

static int64_t bluestein_length(int64_t n) {
  int64_t m = 32;
  while (m < 2 * n - 1) m <<= 1;
  return m;
}

// n_fft = R * m: the largest m <= max_points that the mixed-radix core can transform (smallest R).
// Returns 0 when there is none (a large prime factor, or n not a multiple of 4): Bluestein takes over.
static int64_t sub_length(int64_t n_fft, int* R_out) {
  static const int64_t max_points = []() -> int64_t {     // tuning knob, see profiles/
    const char* env = getenv("PSA_FFT_MAX_POINTS");
    int64_t v = env ? atoll(env) : kDefaultSmemPoints;
    if (v < 64 || v > kMaxSmemPoints) v = kDefaultSmemPoints;
    return v;
  }();
  for (int64_t R = (n_fft + max_points - 1) / max_points; R <= 256 && R <= n_fft; ++R) {
    if (n_fft % R) continue;
    const int64_t m = n_fft / R;
    if (m <= max_points && make_passes((int)m).n_pass >= 0) {
      if (R_out) *R_out = (int)R;
      return m;
    }
  }
  return 0;
}
static bool direct_length(int64_t n) { return n >= 4 && sub_length(n, nullptr) > 0; }

static int64_t pass_table_entries(int64_t n_fft) { return make_passes((int)sub_length(n_fft, nullptr)).total; }

int64_t fft4_table_offset(int64_t n_t) { return n_t + pass_table_entries(n_t); }   // in double2 entries from the plan start

static FftGeom make_geom(int64_t n_fft, const double2* tw) {   // tw = start of the plan: [tw | pass tables | ...]
  FftGeom g;
  g.m = (int)sub_length(n_fft, &g.R);
  g.n_fft = (int)n_fft;
  g.tw = tw;
  g.pass_tw = tw + n_fft;
  g.pp = make_passes(g.m);
  return g;
}

static size_t smem_bytes(const FftGeom& g, bool with_acc) {
  const size_t padded = (size_t)(g.m + (g.m >> 4));
  return padded * sizeof(sc) + (with_acc ? padded * sizeof(float) : 0);
}

template <class K>
static int allow_smem(K kernel, size_t bytes) {
  PSA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return PSA_OK;
}

static int build_tables(int64_t n_fft, double2* tw, cudaStream_t s) {
  twiddle_kernel<<<(unsigned)((n_fft + 255) / 256), 256, 0, s>>>(n_fft, tw);
  const PassPlan pp = make_passes((int)sub_length(n_fft, nullptr));
  double2* pass = tw + n_fft;
  for (int p = 0; p < pp.n_pass; ++p) {
    const int entries = pp.L[p] / pp.r[p];
    pass_table_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, s>>>(pp.L[p], pp.r[p], pass + pp.tw_off[p]);
  }
  return launch_status("fft table kernels");
}

int fft_plan_bytes(int64_t n_t, int64_t* bytes) {
  if (n_t < 1 || n_t > kMaxTransform / 2) {
    set_error("psa_fft: n_t=%lld is outside the supported range [1, %lld]", (long long)n_t, (long long)(kMaxTransform / 2));
    return PSA_ERR_UNSUPPORTED;
  }
  if (direct_length(n_t)) {
    *bytes = (n_t + pass_table_entries(n_t) + (fft4_supported(n_t) ? fft4_table_entries(n_t) : 0)) * (int64_t)sizeof(double2);
  } else {
    const int64_t M = bluestein_length(n_t);
    *bytes = (3 * M + pass_table_entries(M) + n_t) * (int64_t)sizeof(double2);
  }
  return PSA_OK;
}

int launch_fft_plan(int64_t n_t, void* plan_buf, cudaStream_t s) {
  int64_t bytes = 0;
  int st = fft_plan_bytes(n_t, &bytes);
  if (st != PSA_OK) return st;
  double2* plan = reinterpret_cast<double2*>(plan_buf);
  if (direct_length(n_t)) {
    if ((st = build_tables(n_t, plan, s)) != PSA_OK) return st;
    return fft4_supported(n_t) ? launch_fft4_tables(n_t, plan + fft4_table_offset(n_t), s) : PSA_OK;
  }
  const int64_t M = bluestein_length(n_t);
  double2* tw = plan;
  double2* chirp = tw + M + pass_table_entries(M);
  double2* bhat = chirp + n_t;
  double2* work = bhat + M;
  if ((st = build_tables(M, tw, s)) != PSA_OK) return st;
  chirp_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(n_t, M, chirp, work);
  FftGeom g = make_geom(M, tw);
  st = allow_smem(fft_c2c_kernel, smem_bytes(g, false));
  if (st != PSA_OK) return st;
  fft_c2c_kernel<<<(unsigned)g.R, kFftThreads, smem_bytes(g, false), s>>>(work, bhat, g);
  return launch_status("fft plan kernels");
}

int fft_workspace_bytes(int64_t n_t, int64_t n_k, int64_t n_groups, int64_t* bytes) {
  int64_t plan = 0;
  int st = fft_plan_bytes(n_t, &plan);
  if (st != PSA_OK) return st;
  // 8192 / 16384 / 32768 frames: the four-step kernel's L2-resident intermediate + tile counters (coherent assembly)
  if (fft4_supported(n_t)) *bytes = fft4_workspace_bytes(n_t, n_k);
  else *bytes = direct_length(n_t) ? 0 : n_groups * n_k * 3 * bluestein_length(n_t) * (int64_t)sizeof(double2);
  return PSA_OK;
}

int launch_fft(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
               const void* plan_buf, void* workspace, int64_t workspace_bytes, const float* window, int mode, void* out,
               int64_t n_k_total, int64_t k_offset, cudaStream_t s) {
  if (n_k == 0 || n_t == 0) return PSA_OK;
  PSA_REQUIRE(mode == PSA_MODE_COHERENT || mode == PSA_MODE_INCOHERENT, "psa_fft_sed: unknown mode %d", mode);
  int64_t need = 0;
  int st = fft_workspace_bytes(n_t, n_k, n_groups, &need);
  if (st != PSA_OK) return st;
  PSA_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need),
              "psa_fft_sed: workspace of %lld bytes required for n_t=%lld (got %lld)", (long long)need,
              (long long)n_t, (long long)workspace_bytes);
  const double2* plan = reinterpret_cast<const double2*>(plan_buf);
  SedArgs a{P, (int)n_groups, group_stride, ldp, (int)n_t, out, n_k_total, k_offset, 1.0 / (double)n_t, window};
  const bool coherent = mode == PSA_MODE_COHERENT;

  if (coherent && fft4_supported(n_t))
    return launch_fft4(P, n_k, n_t, ldp, plan_buf, workspace, workspace_bytes, window, out, n_k_total, k_offset, s);

  if (need == 0 || direct_length(n_t)) {                 // mixed-radix lengths: one fused kernel
    FftGeom g = make_geom(n_t, plan);
    const size_t smem = smem_bytes(g, !coherent);
    // 3000 columns of 16384 points: 0.790 ms with direct 8-byte stores, 0.769 with pairs, 0.757 with clusters of 4
    static const int cluster_env = getenv("PSA_FFT_CLUSTER") ? atoi(getenv("PSA_FFT_CLUSTER")) : 4;
    const int C = (cluster_env == 2 || cluster_env == 4) ? cluster_env : 1;
    if (coherent && C > 1 && g.m / g.pp.blk <= kFftThreads && g.m % C == 0 && n_k * 3 >= C) {
      if ((st = allow_smem(fft_sed_cluster_kernel, smem)) != PSA_OK) return st;
      const int64_t groups = (n_k * 3 + C - 1) / C;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(groups * g.R * C));
      cfg.blockDim = dim3(kFftThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = (unsigned)C;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      PSA_CUDA(cudaLaunchKernelEx(&cfg, fft_sed_cluster_kernel, a, g, (int)(n_k * 3)));
    } else if (coherent) {
      if ((st = allow_smem(fft_sed_kernel<PSA_MODE_COHERENT>, smem)) != PSA_OK) return st;
      fft_sed_kernel<PSA_MODE_COHERENT><<<(unsigned)(n_k * 3 * g.R), kFftThreads, smem, s>>>(a, g);
    } else {
      if ((st = allow_smem(fft_sed_kernel<PSA_MODE_INCOHERENT>, smem)) != PSA_OK) return st;
      fft_sed_kernel<PSA_MODE_INCOHERENT><<<(unsigned)(n_k * g.R), kFftThreads, smem, s>>>(a, g);
    }
    return launch_status("fft_sed_kernel");
  }

  const int64_t M = bluestein_length(n_t);
  const double2* tw = plan;
  const double2* chirp = tw + M + pass_table_entries(M);
  const double2* bhat = chirp + n_t;
  double2* scratch = reinterpret_cast<double2*>(workspace);
  FftGeom g = make_geom(M, tw);
  if ((st = allow_smem(bluestein_forward_kernel, smem_bytes(g, false))) != PSA_OK) return st;
  bluestein_forward_kernel<<<(unsigned)(n_groups * n_k * 3 * g.R), kFftThreads, smem_bytes(g, false), s>>>(
      a, g, (int)n_k, chirp, bhat, scratch);
  if ((st = launch_status("bluestein_forward_kernel")) != PSA_OK) return st;
  const size_t smem = smem_bytes(g, !coherent);
  if (coherent) {
    if ((st = allow_smem(bluestein_inverse_kernel<PSA_MODE_COHERENT>, smem)) != PSA_OK) return st;
    bluestein_inverse_kernel<PSA_MODE_COHERENT><<<(unsigned)(n_k * 3 * g.R), kFftThreads, smem, s>>>(a, g, (int)n_k, chirp, scratch);
  } else {
    if ((st = allow_smem(bluestein_inverse_kernel<PSA_MODE_INCOHERENT>, smem)) != PSA_OK) return st;
    bluestein_inverse_kernel<PSA_MODE_INCOHERENT><<<(unsigned)(n_k * g.R), kFftThreads, smem, s>>>(a, g, (int)n_k, chirp, scratch);
  }
  return launch_status("bluestein_inverse_kernel");
}

}  // namespace psa
