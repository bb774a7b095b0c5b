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
// Inverse projection (reference: sed_calculator.py:440-441, 494-499, 533).  The spatial phase
// k_act * (mean . khat) is formed in float32 like the reference's float32 products, the rotating
// exponential in float64 (the reference promotes to complex128), the sum with the mean in float32.
// ---------------------------------------------------------------------------------------------
__global__ void ised_kernel(const float* __restrict__ mean, const double* __restrict__ amp,
                            const float* __restrict__ khat, float k_act, double scale, int add_mean,
                            int64_t n_a, int64_t n_frames, float* __restrict__ out) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_a) return;
  const float h0 = __ldg(khat), h1 = __ldg(khat + 1), h2 = __ldg(khat + 2);
  const float m0 = mean[a * 3], m1 = mean[a * 3 + 1], m2 = mean[a * 3 + 2];
  const float xproj = __fmaf_rn(m2, h2, __fmaf_rn(m1, h1, __fmul_rn(m0, h0)));
  const double spatial = (double)__fmul_rn(k_act, xproj);
  const double ar[3] = {amp[a * 6 + 0], amp[a * 6 + 2], amp[a * 6 + 4]};
  const double ai[3] = {amp[a * 6 + 1], amp[a * 6 + 3], amp[a * 6 + 5]};
  const float mm[3] = {m0, m1, m2};
  const bool active = ar[0] != 0. || ai[0] != 0. || ar[1] != 0. || ai[1] != 0. || ar[2] != 0. || ai[2] != 0.;
  for (int64_t f = blockIdx.y; f < n_frames; f += gridDim.y) {
    double s = 0., c = 1.;
    if (active) {
      double tau = 6.283185307179586 * (double)f / (double)n_frames;
      sincos(tau - spatial, &s, &c);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float w = active ? (float)(scale * (ar[p] * c - ai[p] * s)) : 0.f;
      out[(f * n_a + a) * 3 + p] = add_mean ? __fadd_rn(mm[p], w) : w;
    }
  }
}

int launch_ised(const float* mean, const double* amp, const float* khat, float k_act, double scale, int add_mean,
                int64_t n_a, int64_t n_frames, float* out, cudaStream_t s) {
  if (n_a == 0 || n_frames == 0) return PSA_OK;
  dim3 grid((unsigned)((n_a + 127) / 128), (unsigned)(n_frames < 64 ? n_frames : 64));
  ised_kernel<<<grid, 128, 0, s>>>(mean, amp, khat, k_act, scale, add_mean, n_a, n_frames, out);
  return launch_status("ised_kernel");
}

// ---------------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------------
__global__ void disp_moments_kernel(const float* __restrict__ pos, const float* __restrict__ mean,
                                    const int32_t* __restrict__ idx, int64_t n_t, int64_t n_a, int64_t n_sel,
                                    double* __restrict__ out2) {
  double s1 = 0., s2 = 0.;
  const int64_t total = n_t * n_sel;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e / n_sel, j = e % n_sel;
    int64_t atom = idx ? (int64_t)__ldg(idx + j) : j;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      float d = __fsub_rn(__ldg(pos + (t * n_a + atom) * 3 + p), __ldg(mean + atom * 3 + p));
      s1 += (double)d;
      s2 += (double)d * (double)d;
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
  int64_t blocks = (n_t * n_sel + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  disp_moments_kernel<<<(unsigned)blocks, 256, 0, s>>>(pos, mean, idx, n_t, n_a, n_sel, out2);
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
