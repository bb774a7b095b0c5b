// Element-wise consumers of the SED: chiral phase, intensity, inverse projection (iSED) and the
// reductions iSED's 'auto' rescale needs.  All are HBM-bound streaming kernels.
#include "common.cuh"

namespace psa {

// ---------------------------------------------------------------------------------------------
// Chiral phase (reference: sed_calculator.py:338-371).  Option "C" is evaluated in float32 with the
// reference's operation order: angle difference, wrap to [-pi, pi) with a floored modulo, then fold
// the outer quadrants back.  Options "A"/"B" follow the reference's scalar loop (float32 products,
// threshold 1e-18 on |Z|^2, clip, acos / asin).
// ---------------------------------------------------------------------------------------------
__global__ void chiral_kernel(const float2* __restrict__ z1, const float2* __restrict__ z2, int64_t n,
                              int64_t stride1, int64_t stride2, int opt, float* __restrict__ out) {
  const float PI = 3.14159274101257324f, TWO_PI = 6.28318548202514648f, HALF_PI = 1.57079637050628662f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 a = z1[i * stride1], b = z2[i * stride2];
    float res;
    if (opt == 'C') {
      float d = __fsub_rn(atan2f(a.y, a.x), atan2f(b.y, b.x));
      float x = __fadd_rn(d, PI);
      float mod = fmodf(x, TWO_PI);                 // floored modulo, like numpy's %
      if (mod != 0.f && mod < 0.f) mod = __fadd_rn(mod, TWO_PI);
      d = __fsub_rn(mod, PI);
      if (d > HALF_PI) d = __fsub_rn(PI, d);
      else if (d < -HALF_PI) d = __fsub_rn(-PI, d);
      res = d;
    } else {
      float m1 = a.x * a.x + a.y * a.y, m2 = b.x * b.x + b.y * b.y;
      if (m1 < 1e-18f || m2 < 1e-18f) {
        res = 0.f;
      } else {
        float den = sqrtf(m1) * sqrtf(m2);
        float arg = (opt == 'A') ? (a.x * b.x + a.y * b.y) / den : (a.x * b.y - a.y * b.x) / den;
        arg = fminf(1.f, fmaxf(-1.f, arg));
        res = (opt == 'A') ? acosf(arg) : asinf(arg);
      }
    }
    out[i] = res;
  }
}

int launch_chiral(const float2* z1, const float2* z2, int64_t n, int64_t stride1, int64_t stride2, int opt,
                  float* out, cudaStream_t s) {
  if (n == 0) return PSA_OK;
  PSA_REQUIRE(opt == 'A' || opt == 'B' || opt == 'C', "psa_chiral_phase: option must be 'A', 'B' or 'C'");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  chiral_kernel<<<(unsigned)blocks, 256, 0, s>>>(z1, z2, n, stride1, stride2, opt, out);
  return launch_status("chiral_kernel");
}

// ---------------------------------------------------------------------------------------------
// intensity = sum over the last axis of |sed|^2 (reference: sed.py:22-24)
// ---------------------------------------------------------------------------------------------
__global__ void intensity_kernel(const float2* __restrict__ sed, int64_t n_rows, int n_pol, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < n_pol; ++p) {
      float2 v = sed[i * n_pol + p];
      float mag = hypotf(v.x, v.y);          // |z| first, then squared, as np.abs(z)**2 does
      acc += mag * mag;
    }
    out[i] = acc;
  }
}

int launch_intensity(const float2* sed, int64_t n_rows, int n_pol, float* out, cudaStream_t s) {
  if (n_rows == 0) return PSA_OK;
  int64_t blocks = (n_rows + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  intensity_kernel<<<(unsigned)blocks, 256, 0, s>>>(sed, n_rows, n_pol, out);
  return launch_status("intensity_kernel");
}

// ---------------------------------------------------------------------------------------------
// Inverse projection, batched over (k, omega) points (reference: sed_calculator.py:440-441, 494-533).
//
// For point p and group g the reference adds   Re( A[p][g][pol] * exp(i tau_f - i k_p (mean_a . khat)) )
// into wiggles[f][a][pol] for the atoms a of g, group after group (float64 term, float32 running sum), then
// rescales and adds the mean.  With c_a, s_a = cos, sin of the float32 spatial phase k_p * (mean_a . khat) and
// C_f, S_f = cos, sin of tau_f = 2 pi f / n_frames,
//     Re(A e^{i(tau_f - x_a)}) = C_f (Ar c_a + Ai s_a) + S_f (Ar s_a - Ai c_a) = C_f U + S_f V,
// so one float64 sincos per (atom, point) and two float64 FMAs per output value: the kernel is bound by the
// 12 bytes it writes per (point, frame, atom).  A thread owns one atom of one point and walks the frames;
// the groups an atom belongs to come as a CSR list in processing order (atoms in no group keep the mean).
// ---------------------------------------------------------------------------------------------
constexpr int kIsedMaxFrames = 65536;           // frames of one reconstruction (grid z = frames / 32)

constexpr int kIsedFrameChunk = 32;             // frames per block (grid z): small blocks keep the last wave short

__device__ __forceinline__ void ised_fill_phasors(double2* tab, int f_begin, int f_end, int n_frames) {
  for (int f = f_begin + threadIdx.x; f < f_end; f += blockDim.x) {
    double s, c;
    sincos(6.283185307179586 * (double)f / (double)n_frames, &s, &c);   // np.linspace(0, 2 pi, n, endpoint=False)
    tab[f - f_begin] = make_double2(c, s);
  }
}

// U, V of one atom for one of its groups (float64)
__device__ __forceinline__ void ised_uv(const IsedBatch& b, int p, int g, double ca, double sa, double (&U)[3], double (&V)[3]) {
#pragma unroll
  for (int pol = 0; pol < 3; ++pol) {
    const float2 A = __ldg(b.amp + ((int64_t)p * b.n_groups + g) * 3 + pol);
    const double ar = (double)A.x, ai = (double)A.y;
    U[pol] = ar * ca + ai * sa;
    V[pol] = ar * sa - ai * ca;
  }
}

// One output value: the float32 running sum over the atom's groups, rescaled and added to the mean in the
// reference's order of operations.
struct IsedAtom {
  float mean[3];
  double U[3], V[3];          // single-group fast path
  double ca, sa;
  int m_begin, m_end;
};

// w / d in float32, correctly rounded (== __fdiv_rn, the reference's `wiggles /= max`), for a divisor that is the same
// for a whole launch: r = RN(1 / d) is formed once, then q = w r is corrected twice with exact FMA residuals
// (Markstein).  5 FMA-pipe instructions instead of the ~15 of a general IEEE division; operands outside a safe
// exponent window (denormal quotients, overflow) take the general path.
struct FastDiv {
  float d, r;
  bool usable;
};
__device__ __forceinline__ FastDiv make_fast_div(float div) {
  FastDiv f;
  f.d = div;
  f.r = __frcp_rn(div);
  const float ad = fabsf(div);
  f.usable = ad > 1e-12f && ad < 1e12f;
  return f;
}
__device__ __forceinline__ float fast_div(const FastDiv& f, float w) {
  const float aw = fabsf(w);
  if (f.usable && aw > 1e-24f && aw < 1e24f) {
    float q = __fmul_rn(w, f.r);
    q = __fmaf_rn(__fmaf_rn(-q, f.d, w), f.r, q);
    q = __fmaf_rn(__fmaf_rn(-q, f.d, w), f.r, q);
    return q;
  }
  return __fdiv_rn(w, f.d);
}

__device__ __forceinline__ void ised_values(const IsedBatch& b, const IsedAtom& at, int p, double2 cs, float (&w)[3],
                                            float& running_max) {
  w[0] = w[1] = w[2] = 0.f;
  for (int m = at.m_begin; m < at.m_end; ++m) {            // float32 running sum, group by group
    double U[3], V[3];
    ised_uv(b, p, __ldg(b.member_grp + m), at.ca, at.sa, U, V);
#pragma unroll
    for (int pol = 0; pol < 3; ++pol) {
      w[pol] = (float)((double)w[pol] + (cs.x * U[pol] + cs.y * V[pol]));
      running_max = fmaxf(running_max, fabsf(w[pol]));     // the reference's running maximum (after every group)
    }
  }
}

// Store one frame of a warp: 32 atoms x 3 values = 384 contiguous bytes of the output.  kWide (n_a % 4 == 0): staged in
// shared memory and written as 24 aligned 16-byte stores (three full 128-byte lines) instead of 96 4-byte pieces.
template <bool kWide>
__device__ __forceinline__ void ised_store_frame(float* __restrict__ row, float* __restrict__ st, int lane, int n_warp,
                                                 bool live, const float (&val)[3]) {
  if (kWide) {
    st[lane * 3] = val[0];
    st[lane * 3 + 1] = val[1];
    st[lane * 3 + 2] = val[2];
    __syncwarp();
    if (lane * 4 < n_warp * 3)                                              // n_warp % 4 == 0 here: whole float4s only
      __stcs(reinterpret_cast<float4*>(row) + lane, reinterpret_cast<const float4*>(st)[lane]);
    // the caller alternates two staging buffers: this one is rewritten two frames from now, after another __syncwarp
  } else if (live) {
    row[lane * 3] = val[0];
    row[lane * 3 + 1] = val[1];
    row[lane * 3 + 2] = val[2];
  }
}

// kWrite = false: max over (frame, atom of a group, pol) of |running sum after that group| per point -> wmax[p]
//                 (the 'auto' rescale's max_wiggle_amp_all, sed_calculator.py:502-504), nothing is stored;
// kWrite = true : out[p][f][a][pol] = mean + ((sum / div[p]) * mul[p]) in float32, the reference's order of operations
//                 (sed_calculator.py:517-533); div = mul = 1 leaves the sum untouched bit for bit.
// A block whose atoms all belong to at most one group (the usual, disjoint-group case) runs a tight loop: per output
// value two float64 FMAs, one conversion, the FMA-corrected division and the add - ~45 instructions per (atom, frame);
// blocks with overlapping groups take the general running-sum loop.
template <bool kWrite, bool kWide>
__global__ void __launch_bounds__(128) ised_batch_kernel(IsedBatch b, const float* __restrict__ div,
                                                         const float* __restrict__ mul, float* __restrict__ out,
                                                         float* __restrict__ wmax) {
  __shared__ double2 phasor[kIsedFrameChunk];
  __shared__ __align__(16) float stage[4][2][96];          // [warp][double buffer][32 atoms x 3]
  const int f_begin = blockIdx.z * kIsedFrameChunk, f_end = min(b.n_frames, f_begin + kIsedFrameChunk);
  ised_fill_phasors(phasor, f_begin, f_end, b.n_frames);
  const int p = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t a_warp = a - lane;                         // first atom of this warp
  const bool live = a < b.n_a;
  float local_max = 0.f;
  IsedAtom at = {};
  if (live) {
    at.mean[0] = __ldg(b.mean + a * 3);
    at.mean[1] = __ldg(b.mean + a * 3 + 1);
    at.mean[2] = __ldg(b.mean + a * 3 + 2);
    const float xproj = __fmaf_rn(at.mean[2], __ldg(b.khat + 2), __fmaf_rn(at.mean[1], __ldg(b.khat + 1),
                                                                          __fmul_rn(at.mean[0], __ldg(b.khat))));
    sincos((double)__fmul_rn(__ldg(b.k_act + p), xproj), &at.sa, &at.ca);
    at.m_begin = __ldg(b.member_off + a);
    at.m_end = __ldg(b.member_off + a + 1);
    if (at.m_end - at.m_begin == 1) ised_uv(b, p, __ldg(b.member_grp + at.m_begin), at.ca, at.sa, at.U, at.V);
  }
  const bool overlapping = __syncthreads_or(live && at.m_end - at.m_begin > 1);   // also: the phasor table is complete
  const float dv = kWrite ? __ldg(div + p) : 1.f, ml = kWrite ? __ldg(mul + p) : 1.f;
  const bool rescale = dv != 1.f || ml != 1.f;                               // uniform over the block
  const FastDiv fdiv = make_fast_div(dv);
  float* o = kWrite ? out + (int64_t)p * b.n_frames * b.n_a * 3 : nullptr;
  const int n_warp = (int)min((int64_t)32, b.n_a - a_warp);                 // atoms of this warp that exist (<= 0: none)
  if (!overlapping) {
    // tight loop; atoms in no group have U = V = 0 and come out as the mean
    const float r = fdiv.r, d = fdiv.d;
    const bool exact_fast = fdiv.usable;
    for (int f = f_begin; f < f_end; ++f) {
      const double2 cs = phasor[f - f_begin];
      float val[3];
#pragma unroll
      for (int pol = 0; pol < 3; ++pol) {
        const float w = (float)(cs.x * at.U[pol] + cs.y * at.V[pol]);
        if (!kWrite) {
          local_max = fmaxf(local_max, fabsf(w));
        } else {
          float q = w;
          if (rescale) {
            if (exact_fast) {                                                // FMA-corrected division, see FastDiv
              q = __fmul_rn(w, r);
              q = __fmaf_rn(__fmaf_rn(-q, d, w), r, q);
              q = __fmaf_rn(__fmaf_rn(-q, d, w), r, q);
            } else {
              q = __fdiv_rn(w, d);
            }
            q = __fmul_rn(q, ml);
          }
          val[pol] = __fadd_rn(at.mean[pol], q);
        }
      }
      if (kWrite)
        ised_store_frame<kWide>(o + ((int64_t)f * b.n_a + a_warp) * 3, stage[warp][f & 1], lane, n_warp, live, val);
    }
  } else {
    for (int f = f_begin; f < f_end; ++f) {
      const double2 cs = phasor[f - f_begin];
      float w[3] = {0.f, 0.f, 0.f};
      if (live) ised_values(b, at, p, cs, w, local_max);
      if (!kWrite) continue;
      float val[3];
#pragma unroll
      for (int pol = 0; pol < 3; ++pol) {
        const float scaled = rescale ? __fmul_rn(fast_div(fdiv, w[pol]), ml) : w[pol];
        val[pol] = __fadd_rn(at.mean[pol], scaled);
      }
      ised_store_frame<kWide>(o + ((int64_t)f * b.n_a + a_warp) * 3, stage[warp][f & 1], lane, n_warp, live, val);
    }
  }
  if (!kWrite) {
    for (int o2 = 16; o2 > 0; o2 >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o2));
    if (lane == 0 && local_max > 0.f) atomicMax(reinterpret_cast<int*>(wmax + p), __float_as_int(local_max));
  }
}

static int ised_check(const IsedBatch& b) {
  PSA_REQUIRE(b.n_frames <= kIsedMaxFrames, "psa_ised: at most %d reconstruction frames per call (got %d)", kIsedMaxFrames,
              b.n_frames);
  static_assert(kIsedMaxFrames / kIsedFrameChunk <= 65535, "grid z");
  PSA_REQUIRE(b.n_points <= 65535, "psa_ised: at most 65535 points per call");
  return PSA_OK;
}

int launch_ised_absmax(const IsedBatch& b, float* wmax, cudaStream_t s) {
  int st = ised_check(b);
  if (st != PSA_OK) return st;
  PSA_CUDA(cudaMemsetAsync(wmax, 0, sizeof(float) * (size_t)(b.n_points > 0 ? b.n_points : 0), s));
  if (b.n_a == 0 || b.n_frames == 0 || b.n_points == 0) return PSA_OK;
  dim3 grid((unsigned)((b.n_a + 127) / 128), (unsigned)b.n_points, (unsigned)((b.n_frames + kIsedFrameChunk - 1) / kIsedFrameChunk));
  ised_batch_kernel<false, false><<<grid, 128, 0, s>>>(b, nullptr, nullptr, nullptr, wmax);
  return launch_status("ised_batch_kernel<max>");
}

int launch_ised_frames(const IsedBatch& b, const float* div, const float* mul, float* out, cudaStream_t s) {
  int st = ised_check(b);
  if (st != PSA_OK) return st;
  if (b.n_a == 0 || b.n_frames == 0 || b.n_points == 0) return PSA_OK;
  dim3 grid((unsigned)((b.n_a + 127) / 128), (unsigned)b.n_points, (unsigned)((b.n_frames + kIsedFrameChunk - 1) / kIsedFrameChunk));
  if (b.n_a % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    ised_batch_kernel<true, true><<<grid, 128, 0, s>>>(b, div, mul, out, nullptr);
  else
    ised_batch_kernel<true, false><<<grid, 128, 0, s>>>(b, div, mul, out, nullptr);
  return launch_status("ised_batch_kernel<write>");
}

// amplitudes of the matched bins: out[p][pol] = sed[w_idx[p]][k_idx[p]][pol]  (sed_calculator.py:494-496)
__global__ void gather_bins_kernel(const float2* __restrict__ sed, int64_t n_k, const int32_t* __restrict__ w_idx,
                                   const int32_t* __restrict__ k_idx, int n_points, int64_t out_stride,
                                   float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points * 3) return;
  const int p = i / 3, pol = i % 3;
  out[(int64_t)p * out_stride + pol] = sed[((int64_t)__ldg(w_idx + p) * n_k + __ldg(k_idx + p)) * 3 + pol];
}

int launch_gather_bins(const float2* sed, int64_t n_k, const int32_t* w_idx, const int32_t* k_idx, int n_points,
                       int64_t out_stride, float2* out, cudaStream_t s) {
  if (n_points == 0) return PSA_OK;
  gather_bins_kernel<<<(unsigned)((n_points * 3 + 127) / 128), 128, 0, s>>>(sed, n_k, w_idx, k_idx, n_points, out_stride, out);
  return launch_status("gather_bins_kernel");
}

// ---------------------------------------------------------------------------------------------
// Device-side consumers of an intensity map (N3 of SURVEY.md 8f): what the reference's plotter and GUI compute on
// the host from the full (n_f, n_k) array - intensity scaling (sed_plotter.py:160-181, psa_gui.py:_apply_intensity_
// scaling), the global colour range (psa_gui.py:2424-2441: nanmin / nanmax) and percentile limits
// (sed_plotter.py:211-215: np.percentile over the finite values) - so that only a heat map, or only two numbers,
// cross PCIe.
// ---------------------------------------------------------------------------------------------
// mode 0 linear, 1 log10(max(x, 1e-12)), 2 sqrt(max(x, 0)), 3 sqrt(sqrt(max(x, 0)))   (float32, in place)
__global__ void scale_intensity_kernel(float* __restrict__ x, int64_t n, int mode) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (mode == 1) v = log10f(fmaxf(v, 1e-12f));
    else if (mode == 2) v = sqrtf(fmaxf(v, 0.f));
    else if (mode == 3) v = sqrtf(sqrtf(fmaxf(v, 0.f)));
    x[i] = v;
  }
}

int launch_scale_intensity(float* x, int64_t n, int mode, cudaStream_t s) {
  PSA_REQUIRE(mode >= 0 && mode <= 3, "psa_scale_intensity: mode must be 0 (linear), 1 (log), 2 (sqrt) or 3 (dsqrt)");
  if (n == 0 || mode == 0) return PSA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_intensity_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, mode);
  return launch_status("scale_intensity_kernel");
}

// order-preserving key of a float32: unsigned comparison of keys == numeric comparison of the floats
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ bool is_finite_bits(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

// out[0] = min, out[1] = max over the non-NaN values (np.nanmin / np.nanmax), as float keys; out[2] = finite count
__global__ void minmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ key_min,
                              unsigned int* __restrict__ key_max, unsigned long long* __restrict__ n_finite) {
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  unsigned long long cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    if (v == v) {                                   // not NaN
      const uint32_t k = float_key(v);
      lo = min(lo, k);
      hi = max(hi, k);
    }
    cnt += is_finite_bits(v) ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(key_min, lo);
    atomicMax(key_max, hi);
    atomicAdd(n_finite, cnt);
  }
}

// One pass of a most-significant-byte-first radix select over the FINITE values of x: for each of `m` searches,
// hist[j][b] = number of values whose key agrees with prefix[j] on its top `done_bits` bits and whose next byte is b.
__global__ void select_pass_kernel(const float* __restrict__ x, int64_t n, int done_bits, const uint32_t* __restrict__ prefix,
                                   int m, unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_hist[4][256];
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  uint32_t pre[4];
  for (int j = 0; j < 4; ++j) pre[j] = j < m ? __ldg(prefix + j) : 0u;
  const uint32_t mask = done_bits == 0 ? 0u : 0xFFFFFFFFu << (32 - done_bits);
  const int shift = 24 - done_bits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    if (!is_finite_bits(v)) continue;
    const uint32_t k = float_key(v);
    for (int j = 0; j < m; ++j)
      if ((k & mask) == (pre[j] & mask)) atomicAdd(&s_hist[j][(k >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m * 256; i += blockDim.x) {
    const unsigned int c = (&s_hist[0][0])[i];
    if (c) atomicAdd(hist + i, c);
  }
}

int launch_minmax(const float* x, int64_t n, void* out3, cudaStream_t s) {
  // out3: [key_min u32][key_max u32][n_finite u64]
  unsigned int init[4] = {0xFFFFFFFFu, 0u, 0u, 0u};
  PSA_CUDA(cudaMemcpyAsync(out3, init, sizeof(init), cudaMemcpyHostToDevice, s));
  PSA_CUDA(cudaStreamSynchronize(s));                       // `init` is on the stack
  if (n == 0) return PSA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  unsigned int* u = reinterpret_cast<unsigned int*>(out3);
  minmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, u, u + 1, reinterpret_cast<unsigned long long*>(u + 2));
  return launch_status("minmax_kernel");
}

int launch_select_pass(const float* x, int64_t n, int done_bits, const uint32_t* prefix, int m, unsigned int* hist,
                       cudaStream_t s) {
  PSA_REQUIRE(m >= 1 && m <= 4, "psa_select_pass: 1 to 4 simultaneous searches");
  PSA_REQUIRE(done_bits == 0 || done_bits == 8 || done_bits == 16 || done_bits == 24, "psa_select_pass: done_bits must be 0, 8, 16 or 24");
  PSA_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned int) * 256 * (size_t)m, s));
  if (n == 0) return PSA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  select_pass_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, done_bits, prefix, m, hist);
  return launch_status("select_pass_kernel");
}

// ---------------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------------
// Sum and sum of squares (float64) of the float32 displacements pos - mean over (frame, selected atom, component).
// grid.x tiles the columns of a frame row, grid.y the frames; a thread keeps its mean values in registers and walks
// down the frames, so a warp reads 512 contiguous bytes per frame (whole rows: one float4 per thread) and there is no
// index arithmetic in the loop.
template <int kMode>   // 0: whole rows as float4 (n_a * 3 divisible by 4), 1: whole rows scalar, 2: gathered atoms
__global__ void __launch_bounds__(256) disp_moments_kernel(const float* __restrict__ pos, const float* __restrict__ mean,
                                                           const int32_t* __restrict__ idx, int64_t n_t, int64_t n_a,
                                                           int64_t n_sel, int64_t t_per_block, double* __restrict__ out2) {
  constexpr int kVals = kMode == 0 ? 4 : (kMode == 1 ? 1 : 3);
  const int64_t row = n_a * 3;
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // float4 / float / selected atom
  const int64_t n_items = kMode == 0 ? row / 4 : (kMode == 1 ? row : n_sel);
  const int64_t t0 = (int64_t)blockIdx.y * t_per_block, t1 = t0 + t_per_block < n_t ? t0 + t_per_block : n_t;
  double s1 = 0., s2 = 0.;
  if (item < n_items) {
    const int64_t col = kMode == 0 ? item * 4 : (kMode == 1 ? item : (int64_t)__ldg(idx + item) * 3);
    float m[kVals];
#pragma unroll
    for (int p = 0; p < kVals; ++p) m[p] = __ldg(mean + col + p);
    const float* src = pos + t0 * row + col;
#pragma unroll 8
    for (int64_t t = t0; t < t1; ++t, src += row) {
      float v[kVals];
      if constexpr (kMode == 0) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int p = 0; p < kVals; ++p) v[p] = __ldg(src + p);
      }
#pragma unroll
      for (int p = 0; p < kVals; ++p) {
        const double d = (double)__fsub_rn(v[p], m[p]);
        s1 += d;
        s2 = fma(d, d, s2);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out2, s1);
    atomicAdd(out2 + 1, s2);
  }
}

int launch_disp_moments(const float* pos, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                        int64_t n_sel, double* out2, cudaStream_t s) {
  PSA_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(double), s));
  if (n_t * n_sel == 0) return PSA_OK;
  const int64_t row = n_a * 3;
  const bool vec = idx == nullptr && (row & 3) == 0 && ((uintptr_t)pos & 15) == 0 && ((uintptr_t)mean & 15) == 0;
  const int64_t n_items = idx != nullptr ? n_sel : (vec ? row / 4 : row);
  const int64_t bx = (n_items + 255) / 256;
  int64_t by = (148 * 8 + bx - 1) / bx;                     // ~8 CTAs per SM in total
  if (by > n_t) by = n_t;
  if (by > 65535) by = 65535;
  const int64_t t_per_block = (n_t + by - 1) / by;
  by = (n_t + t_per_block - 1) / t_per_block;
  const dim3 grid((unsigned)bx, (unsigned)by);
  if (idx != nullptr) disp_moments_kernel<2><<<grid, 256, 0, s>>>(pos, mean, idx, n_t, n_a, n_sel, t_per_block, out2);
  else if (vec) disp_moments_kernel<0><<<grid, 256, 0, s>>>(pos, mean, idx, n_t, n_a, n_sel, t_per_block, out2);
  else disp_moments_kernel<1><<<grid, 256, 0, s>>>(pos, mean, idx, n_t, n_a, n_sel, t_per_block, out2);
  return launch_status("disp_moments_kernel");
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0
}

int launch_absmax(const float* x, int64_t n, float* out, cudaStream_t s) {
  PSA_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  if (n == 0) return PSA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, out);
  return launch_status("absmax_kernel");
}

}  // namespace psa
