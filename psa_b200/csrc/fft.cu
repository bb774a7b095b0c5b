// Time-axis FFT of the projected columns, fused with the 1/n_t scale and the SED assembly.
//
// One CTA transforms one column (k, pol [, group]) entirely in shared memory and writes the spectrum
// straight into the result layout:
//   coherent   : complex64 out[f][k][pol]                   (reference: sed_calculator.py:296-311)
//   incoherent : float32  out[f][k] = sum_g sum_pol |S|^2    (reference: sed_calculator.py:313-327)
//
// Power-of-two core (forward, decimation in frequency, in place, m = 2^s points, 32 <= m <= 16384):
//   * shared-memory passes: one radix-2 pass if s-5 is odd, then radix-4 passes down to blocks of 32.
//     Every butterfly leg is >= 32 elements away from the next, so a warp always touches 32
//     consecutive elements: conflict-free.
//   * final stage: each thread pulls one contiguous 32-point block into registers, finishes it with
//     radix 4 x 4 x 2 and hands the 32 results (digit-reversed frequency index) to a sink.  The array
//     is padded by one element per 32 (index p lives at p + p/32), which makes the per-thread
//     contiguous block reads conflict-free as well.
//   * transforms longer than 16384 points do not fit one CTA's shared memory: they are split by a
//     radix-R decimation-in-frequency step applied while loading, giving R independent
//     sub-transforms that produce the frequencies f = R f' + r.
//   * twiddles come from a correctly rounded float32 table (computed in float64), like pocketfft's.
//
// Frame counts that are not a power of two (the reference accepts any n_t through pocketfft) use
// Bluestein's chirp-z identity on top of the same core: with b_t = exp(i pi t^2 / n),
//   X_f = conj(b_f) * sum_t (x_t conj(b_t)) b_{f-t},
// i.e. one forward transform of length M >= 2n-1 (M = 2^s), a point-wise product with the
// precomputed spectrum of b, and one inverse transform; the chirped spectrum of each column makes
// one round trip through a scratch buffer.
#include <stdlib.h>

#include "common.cuh"

namespace psa {

constexpr int kFftThreads = 512;
constexpr int64_t kMaxSmemPoints = 16384;   // complex64 points that fit one CTA (128 KiB + padding)
constexpr int64_t kMaxTransform = (int64_t)1 << 20;
constexpr int kBlk = 32;                    // points finished in registers per thread

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ int phys(int p) { return p + (p >> 5); }

// forward radix-4 DIF butterfly on four legs (no twiddles)
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// ---------------------------------------------------------------------------------------------
// the power-of-two core
// ---------------------------------------------------------------------------------------------
// Shared-memory passes: reduce the m-point problem to m/32 independent contiguous 32-point blocks.
// Radix-4 pass twiddles come from per-pass tables (w_L^j, w_L^2j, w_L^3j stored contiguously in j), so
// a warp reads three short contiguous runs instead of 96 scattered entries of the length-n table.
__host__ __device__ inline int pass_table_base(int L) { return 3 * ((L >> 2) - 16); }   // entries before pass L (L >= 64)

__device__ void fft_smem_passes(float2* __restrict__ s, int m, int log2m, const float2* __restrict__ tw, int tw_n,
                                const float2* __restrict__ pass_tw) {
  int L = m;
  if ((log2m - 5) & 1) {   // radix-2 pass
    const int half = L >> 1, tstep = tw_n / L;
#pragma unroll 4
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int i0 = phys(j), i1 = phys(j + half);
      float2 a = s[i0], b = s[i1];
      s[i0] = cadd(a, b);
      s[i1] = cmul(csub(a, b), __ldg(tw + (int64_t)j * tstep));
    }
    L = half;
    __syncthreads();
  }
  for (; L > kBlk; L >>= 2) {
    const int q = L >> 2;
    const float2* __restrict__ t1 = pass_tw + pass_table_base(L);
    const float2* __restrict__ t2 = t1 + q;
    const float2* __restrict__ t3 = t2 + q;
#pragma unroll 4
    for (int b = threadIdx.x; b < (m >> 2); b += blockDim.x) {
      const int j = b & (q - 1);
      const int base = ((b - j) << 2) + j;            // (b / q) * L + j
      const int i0 = phys(base), i1 = phys(base + q), i2 = phys(base + 2 * q), i3 = phys(base + 3 * q);
      float2 a0 = s[i0], a1 = s[i1], a2 = s[i2], a3 = s[i3];
      bfly4(a0, a1, a2, a3);
      a1 = cmul(a1, __ldg(t1 + j));
      a2 = cmul(a2, __ldg(t2 + j));
      a3 = cmul(a3, __ldg(t3 + j));
      s[i0] = a0; s[i1] = a1; s[i2] = a2; s[i3] = a3;
    }
    __syncthreads();
  }
}

// Finish one contiguous 32-point block held in registers: radix 4 (L=32), radix 4 (L=8), radix 2.
// Register e then holds the block-local frequency (e>>3) + 4*((e>>1)&3) + 16*(e&1).
__device__ __forceinline__ void fft32_registers(float2 (&x)[kBlk], const float2* __restrict__ tw, int tw_n) {
  const int t32 = tw_n >> 5;   // w_32^k = tw[k * t32]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bfly4(x[j], x[j + 8], x[j + 16], x[j + 24]);
    if (j != 0) {
      x[j + 8] = cmul(x[j + 8], __ldg(tw + (int64_t)(j * 1) * t32));
      x[j + 16] = cmul(x[j + 16], __ldg(tw + (int64_t)(j * 2) * t32));
      x[j + 24] = cmul(x[j + 24], __ldg(tw + (int64_t)(j * 3) * t32));
    }
  }
  const float2 w8_1 = __ldg(tw + (int64_t)4 * t32), w8_2 = __ldg(tw + (int64_t)8 * t32),
               w8_3 = __ldg(tw + (int64_t)12 * t32);
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    const int o = blk * 8;
    bfly4(x[o], x[o + 2], x[o + 4], x[o + 6]);
    bfly4(x[o + 1], x[o + 3], x[o + 5], x[o + 7]);
    x[o + 3] = cmul(x[o + 3], w8_1);
    x[o + 5] = cmul(x[o + 5], w8_2);
    x[o + 7] = cmul(x[o + 7], w8_3);
  }
#pragma unroll
  for (int i = 0; i < kBlk; i += 2) {
    float2 a = x[i], b = x[i + 1];
    x[i] = cadd(a, b);
    x[i + 1] = csub(a, b);
  }
}

// frequency (within the m-point sub-transform) of element 0 of block b; element e adds (m/32) * rev(e)
__device__ __forceinline__ int block_base_frequency(int b, int log2m) {
  int bits = log2m - 5;        // bits of the block index, consumed most-significant first
  int f = 0, shift = 0;
  if (bits & 1) {
    bits -= 1;
    f = (b >> bits) & 1;
    shift = 1;
  }
  while (bits > 0) {
    bits -= 2;
    f += ((b >> bits) & 3) << shift;
    shift += 2;
  }
  return f;
}

// Geometry of one launch: transform length n_fft = m * R, twiddle table of length n_fft.
struct FftGeom {
  int m, log2m, R, n_fft;
  const float2* tw;        // w_n^j, j < n_fft
  const float2* pass_tw;   // per-pass radix-4 tables, see pass_table_base
};

// Fill shared memory with sub-sequence r of the radix-R split of fetch(0..n_fft) (R == 1: plain copy).
template <class Fetch>
__device__ void load_column(float2* __restrict__ s, const Fetch& fetch, const FftGeom& g, int r) {
  if (g.R == 1) {
#pragma unroll 4
    for (int t = threadIdx.x; t < g.m; t += blockDim.x) s[phys(t)] = fetch(t);
    return;
  }
  for (int t = threadIdx.x; t < g.m; t += blockDim.x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < g.R; ++j) {
      float2 x = fetch(t + j * g.m);
      int wi = ((j * r) & (g.R - 1)) * g.m;                        // w_R^{jr} = w_n^{(jr mod R) m}
      acc = cadd(acc, wi ? cmul(x, __ldg(g.tw + wi)) : x);
    }
    s[phys(t)] = r ? cmul(acc, __ldg(g.tw + (int64_t)t * r)) : acc;   // w_n^{tr}, t r < n
  }
}

// Transform the column in shared memory and feed every (slot, frequency, value) to the sink.
// slot = padded in-place position, owned by the same thread on every call with the same geometry.
template <class Sink>
__device__ void transform_and_emit(float2* __restrict__ s, const FftGeom& g, int r, Sink& sink) {
  fft_smem_passes(s, g.m, g.log2m, g.tw, g.n_fft, g.pass_tw);
  const int n_blocks = g.m >> 5, fstep = g.m >> 5;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
    float2 x[kBlk];
#pragma unroll
    for (int e = 0; e < kBlk; ++e) x[e] = s[b * (kBlk + 1) + e];
    fft32_registers(x, g.tw, g.n_fft);
    const int f0 = block_base_frequency(b, g.log2m);
#pragma unroll
    for (int e = 0; e < kBlk; ++e) {
      const int rev = (e >> 3) + 4 * ((e >> 1) & 3) + 16 * (e & 1);
      sink(b * (kBlk + 1) + e, (f0 + fstep * rev) * g.R + r, x[e]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fetchers
// ---------------------------------------------------------------------------------------------
struct FetchPlanar {            // rows of P: Re and Im as separate float rows
  const float* re;
  const float* im;
  __device__ float2 operator()(int t) const { return make_float2(__ldg(re + t), __ldg(im + t)); }
};
struct FetchChirped {           // x_t * conj(b_t) for t < n, zero padding up to the transform length
  const float* re;
  const float* im;
  const float2* chirp;
  int n;
  __device__ float2 operator()(int t) const {
    if (t >= n) return make_float2(0.f, 0.f);
    return cmul_conj(make_float2(__ldg(re + t), __ldg(im + t)), __ldg(chirp + t));
  }
};
struct FetchConj {              // conj of an interleaved complex column (inverse transform by conjugation)
  const float2* src;
  __device__ float2 operator()(int t) const {
    float2 v = __ldg(src + t);
    return make_float2(v.x, -v.y);
  }
};
struct FetchComplex {
  const float2* src;
  __device__ float2 operator()(int t) const { return __ldg(src + t); }
};

// ---------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------
// value -> S = value * scale (power-of-two path) or the Bluestein unchirp; then the SED assembly.
struct Unchirp {                // Bluestein: X_f = conj(b_f) * conj(y_f) / M, then / n   (f < n only)
  const float2* chirp;
  int n;
  float inv_m, n_f;
  __device__ bool operator()(int f, float2 y, float2& out) const {
    if (f >= n) return false;
    float2 conv = make_float2(y.x * inv_m, -y.y * inv_m);
    float2 X = cmul_conj(conv, __ldg(chirp + f));
    out = make_float2(X.x / n_f, X.y / n_f);            // divide by n_t like the reference
    return true;
  }
};
struct ScaleOnly {              // power of two: multiply by the exact reciprocal 1 / n_t
  float inv_n;
  __device__ bool operator()(int, float2 y, float2& out) const {
    out = make_float2(y.x * inv_n, y.y * inv_n);
    return true;
  }
};

template <class Post>
struct SinkCoherent {
  float2* o;                    // already offset to (k, pol)
  int64_t fstride;
  Post post;
  __device__ void operator()(int, int f, float2 v) const {
    float2 S;
    if (post(f, v, S)) o[(int64_t)f * fstride] = S;
  }
};
template <class Post>
struct SinkAccumulate {         // s_acc[slot] += |S|^2; the slot is private to the calling thread
  float* s_acc;
  Post post;
  __device__ void operator()(int slot, int f, float2 v) const {
    float2 S;
    if (post(f, v, S)) s_acc[slot] += S.x * S.x + S.y * S.y;
  }
};
struct SinkFlush {              // write the accumulated intensities of one k column
  const float* s_acc;
  float* o;                     // already offset to k
  int64_t fstride;
  int n_valid;
  __device__ void operator()(int slot, int f, float2) const {
    if (f < n_valid) o[(int64_t)f * fstride] = s_acc[slot];
  }
};
struct SinkTimesSpectrum {      // natural-order store of value * bhat[f] (Bluestein forward leg)
  float2* dst;
  const float2* bhat;           // nullptr: plain store
  __device__ void operator()(int, int f, float2 v) const { dst[f] = bhat ? cmul(v, __ldg(bhat + f)) : v; }
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
};

__device__ __forceinline__ void column_rows(const SedArgs& a, int g, int k, int pol, const float*& re, const float*& im) {
  const float* base = a.P + (int64_t)g * a.group_stride;
  re = base + ((int64_t)(2 * k) * 3 + pol) * a.ldp;
  im = base + ((int64_t)(2 * k + 1) * 3 + pol) * a.ldp;
}

// Power-of-two n_t: P -> result in one kernel.
template <int kMode>
__global__ void __launch_bounds__(kFftThreads) fft_sed_kernel(SedArgs a, FftGeom g) {
  extern __shared__ float2 s_data[];
  const ScaleOnly post{1.0f / (float)a.n_t};
  const int r = blockIdx.x % g.R;
  if (kMode == PSA_MODE_COHERENT) {          // block -> (k, pol, r)
    const int pol = (blockIdx.x / g.R) % 3, k = blockIdx.x / (3 * g.R);
    FetchPlanar fetch;
    column_rows(a, 0, k, pol, fetch.re, fetch.im);
    load_column(s_data, fetch, g, r);
    __syncthreads();
    SinkCoherent<ScaleOnly> sink{reinterpret_cast<float2*>(a.out) + (a.k_offset + k) * 3 + pol, a.n_k_total * 3, post};
    transform_and_emit(s_data, g, r, sink);
  } else {                                   // block -> (k, r); loop over groups and polarisations
    const int k = blockIdx.x / g.R;
    const int padded = g.m + (g.m >> 5);
    float* s_acc = reinterpret_cast<float*>(s_data + padded);
    for (int i = threadIdx.x; i < padded; i += blockDim.x) s_acc[i] = 0.f;
    SinkAccumulate<ScaleOnly> acc{s_acc, post};
    for (int grp = 0; grp < a.n_groups; ++grp)
      for (int pol = 0; pol < 3; ++pol) {
        FetchPlanar fetch;
        column_rows(a, grp, k, pol, fetch.re, fetch.im);
        __syncthreads();
        load_column(s_data, fetch, g, r);
        __syncthreads();
        transform_and_emit(s_data, g, r, acc);
      }
    // every slot was accumulated and is flushed by the same thread: no barrier needed
    SinkFlush flush{s_acc, reinterpret_cast<float*>(a.out) + a.k_offset + k, a.n_k_total, g.n_fft};
    const int n_blocks = g.m >> 5, fstep = g.m >> 5;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
      const int f0 = block_base_frequency(b, g.log2m);
#pragma unroll
      for (int e = 0; e < kBlk; ++e) {
        const int rev = (e >> 3) + 4 * ((e >> 1) & 3) + 16 * (e & 1);
        flush(b * (kBlk + 1) + e, (f0 + fstep * rev) * g.R + r, make_float2(0.f, 0.f));
      }
    }
  }
}

// Bluestein leg 1: column -> chirp, zero-pad, forward transform, times the chirp spectrum -> scratch.
// block -> (column, r), column = (group, k, pol) flattened.
__global__ void __launch_bounds__(kFftThreads) bluestein_forward_kernel(SedArgs a, FftGeom g, int n_k,
                                                                        const float2* __restrict__ chirp,
                                                                        const float2* __restrict__ bhat,
                                                                        float2* __restrict__ scratch) {
  extern __shared__ float2 s_data[];
  const int r = blockIdx.x % g.R;
  const int col = blockIdx.x / g.R;
  const int pol = col % 3, k = (col / 3) % n_k, grp = col / (3 * n_k);
  FetchChirped fetch;
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
__global__ void __launch_bounds__(kFftThreads) bluestein_inverse_kernel(SedArgs a, FftGeom g, int n_k,
                                                                        const float2* __restrict__ chirp,
                                                                        const float2* __restrict__ scratch) {
  extern __shared__ float2 s_data[];
  const Unchirp post{chirp, a.n_t, 1.0f / (float)g.n_fft, (float)a.n_t};
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
    const int padded = g.m + (g.m >> 5);
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
    SinkFlush flush{s_acc, reinterpret_cast<float*>(a.out) + a.k_offset + k, a.n_k_total, a.n_t};
    const int n_blocks = g.m >> 5, fstep = g.m >> 5;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
      const int f0 = block_base_frequency(b, g.log2m);
#pragma unroll
      for (int e = 0; e < kBlk; ++e) {
        const int rev = (e >> 3) + 4 * ((e >> 1) & 3) + 16 * (e & 1);
        flush(b * (kBlk + 1) + e, (f0 + fstep * rev) * g.R + r, make_float2(0.f, 0.f));
      }
    }
  }
}

// plain complex-to-complex forward transform of one column, natural order (used once per plan)
__global__ void __launch_bounds__(kFftThreads) fft_c2c_kernel(const float2* __restrict__ src, float2* __restrict__ dst,
                                                              FftGeom g) {
  extern __shared__ float2 s_data[];
  const int r = blockIdx.x % g.R;
  FetchComplex fetch{src};
  load_column(s_data, fetch, g, r);
  __syncthreads();
  SinkTimesSpectrum sink{dst, nullptr};
  transform_and_emit(s_data, g, r, sink);
}

// ---------------------------------------------------------------------------------------------
// plan: twiddles (+ chirp and its spectrum for non-power-of-two lengths), one caller-owned buffer
//   power of two : [ tw (n_t) | pass tables ]
//   otherwise    : [ tw (M)   | pass tables | chirp (n_t) | bhat (M) | work (M) ]      all float2
// ---------------------------------------------------------------------------------------------
static int64_t pass_table_entries(int64_t n_fft) {
  const int64_t lmax = n_fft < kMaxSmemPoints ? n_fft : kMaxSmemPoints;
  return lmax >= 64 ? pass_table_base((int)lmax) + 3 * (lmax >> 2) : 0;
}

__global__ void pass_table_kernel(int L, float2* __restrict__ table) {   // table already offset to pass L
  const int q = L >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * q) return;
  const int r = i / q + 1, j = i % q;
  double s, c;
  sincospi(2.0 * (double)((r * j) % L) / (double)L, &s, &c);
  table[i] = make_float2((float)c, (float)(-s));
}

__global__ void twiddle_kernel(int64_t n, float2* __restrict__ tw);

static int build_tables(int64_t n_fft, float2* tw, cudaStream_t s) {
  twiddle_kernel<<<(unsigned)((n_fft + 255) / 256), 256, 0, s>>>(n_fft, tw);
  float2* pass = tw + n_fft;
  for (int64_t L = 64; L <= n_fft && L <= kMaxSmemPoints; L <<= 1)
    pass_table_kernel<<<(unsigned)((3 * (L >> 2) + 255) / 256), 256, 0, s>>>((int)L, pass + pass_table_base((int)L));
  return launch_status("fft table kernels");
}

__global__ void twiddle_kernel(int64_t n, float2* __restrict__ tw) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s, c;
  sincospi(2.0 * (double)j / (double)n, &s, &c);
  tw[j] = make_float2((float)c, (float)(-s));
}

// chirp[t] = exp(+i pi t^2 / n) with t^2 reduced mod 2n in integers; padded[] = the circular kernel of length M
__global__ void chirp_kernel(int64_t n, int64_t M, float2* __restrict__ chirp, float2* __restrict__ padded) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M) return;
  float2 val = make_float2(0.f, 0.f);
  int64_t src = t < n ? t : (M - t < n ? M - t : -1);        // b_{-t} = b_t wraps to M - t
  if (src >= 0) {
    int64_t q = (src * src) % (2 * n);
    double s, c;
    sincospi((double)q / (double)n, &s, &c);
    val = make_float2((float)c, (float)s);
    if (t < n) chirp[t] = val;
  }
  padded[t] = val;
}

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

static int64_t bluestein_length(int64_t n) {
  int64_t m = 32;
  while (m < 2 * n - 1) m <<= 1;
  return m;
}

static FftGeom make_geom(int64_t n_fft, const float2* tw) {   // tw = start of the plan: [tw | pass tables | ...]
  static const int64_t max_points = []() -> int64_t {     // tuning knob, see profiles/
    const char* env = getenv("PSA_FFT_MAX_POINTS");
    int64_t v = env ? atoll(env) : kMaxSmemPoints;
    if (v < 64 || v > kMaxSmemPoints || (v & (v - 1))) v = kMaxSmemPoints;
    return v;
  }();
  FftGeom g;
  int64_t m = n_fft;
  int R = 1;
  while (m > max_points) { m >>= 1; R <<= 1; }
  g.m = (int)m;
  g.R = R;
  g.n_fft = (int)n_fft;
  g.log2m = 0;
  while ((1 << g.log2m) < m) ++g.log2m;
  g.tw = tw;
  g.pass_tw = tw + n_fft;
  return g;
}

static size_t smem_bytes(const FftGeom& g, bool with_acc) {
  const size_t padded = (size_t)(g.m + (g.m >> 5));
  return padded * sizeof(float2) + (with_acc ? padded * sizeof(float) : 0);
}

template <class K>
static int allow_smem(K kernel, size_t bytes) {
  PSA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return PSA_OK;
}

int fft_plan_bytes(int64_t n_t, int64_t* bytes) {
  if (n_t < 2 || n_t > kMaxTransform / 2) {
    set_error("psa_fft: n_t=%lld is outside the supported range [2, %lld]", (long long)n_t, (long long)(kMaxTransform / 2));
    return PSA_ERR_UNSUPPORTED;
  }
  if (is_pow2(n_t) && n_t >= kBlk) {
    *bytes = (n_t + pass_table_entries(n_t)) * (int64_t)sizeof(float2);
  } else {
    const int64_t M = bluestein_length(n_t);
    *bytes = (3 * M + pass_table_entries(M) + round_up(n_t, 2)) * (int64_t)sizeof(float2);
  }
  return PSA_OK;
}

int launch_fft_plan(int64_t n_t, float2* plan, cudaStream_t s) {
  int64_t bytes = 0;
  int st = fft_plan_bytes(n_t, &bytes);
  if (st != PSA_OK) return st;
  if (is_pow2(n_t) && n_t >= kBlk) return build_tables(n_t, plan, s);
  const int64_t M = bluestein_length(n_t);
  float2* tw = plan;
  float2* chirp = tw + M + pass_table_entries(M);
  float2* bhat = chirp + round_up(n_t, 2);
  float2* work = bhat + M;
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
  *bytes = (is_pow2(n_t) && n_t >= kBlk) ? 0 : n_groups * n_k * 3 * bluestein_length(n_t) * (int64_t)sizeof(float2);
  return PSA_OK;
}

int launch_fft(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
               const float2* plan, void* workspace, int64_t workspace_bytes, int mode, void* out, int64_t n_k_total,
               int64_t k_offset, cudaStream_t s) {
  if (n_k == 0 || n_t == 0) return PSA_OK;
  PSA_REQUIRE(mode == PSA_MODE_COHERENT || mode == PSA_MODE_INCOHERENT, "psa_fft_sed: unknown mode %d", mode);
  int64_t need = 0;
  int st = fft_workspace_bytes(n_t, n_k, n_groups, &need);
  if (st != PSA_OK) return st;
  PSA_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need),
              "psa_fft_sed: workspace of %lld bytes required for n_t=%lld (got %lld)", (long long)need,
              (long long)n_t, (long long)workspace_bytes);
  SedArgs a{P, (int)n_groups, group_stride, ldp, (int)n_t, out, n_k_total, k_offset};
  const bool coherent = mode == PSA_MODE_COHERENT;

  if (need == 0) {                                       // power of two: one fused kernel
    FftGeom g = make_geom(n_t, plan);
    const size_t smem = smem_bytes(g, !coherent);
    if (coherent) {
      if ((st = allow_smem(fft_sed_kernel<PSA_MODE_COHERENT>, smem)) != PSA_OK) return st;
      fft_sed_kernel<PSA_MODE_COHERENT><<<(unsigned)(n_k * 3 * g.R), kFftThreads, smem, s>>>(a, g);
    } else {
      if ((st = allow_smem(fft_sed_kernel<PSA_MODE_INCOHERENT>, smem)) != PSA_OK) return st;
      fft_sed_kernel<PSA_MODE_INCOHERENT><<<(unsigned)(n_k * g.R), kFftThreads, smem, s>>>(a, g);
    }
    return launch_status("fft_sed_kernel");
  }

  const int64_t M = bluestein_length(n_t);
  const float2* tw = plan;
  const float2* chirp = tw + M + pass_table_entries(M);
  const float2* bhat = chirp + round_up(n_t, 2);
  float2* scratch = reinterpret_cast<float2*>(workspace);
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
