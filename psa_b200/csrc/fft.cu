// Time-axis FFT of the projected columns, fused with the 1/n_t scale and the SED assembly.
//
// One CTA transforms one column (k, pol [, group]) entirely in shared memory: coalesced float4-free
// planar loads of Re/Im rows of P, in-place decimation-in-frequency passes (radix 4, one leading
// radix-2 pass when log2 is odd), digit-reversed read-out straight into the result layout
//   coherent   : complex64 out[f][k][pol]                   (reference: sed_calculator.py:296-311)
//   incoherent : float32  out[f][k] = sum_g sum_pol |S|^2    (reference: sed_calculator.py:313-327)
// so spectra never round-trip through HBM.  Columns longer than 16384 points (128 KiB of complex64)
// do not fit one CTA's shared memory: they are split by a radix-R decimation-in-frequency step done
// while loading, giving R independent sub-transforms that each produce the frequencies f = R f' + r.
// Twiddles come from a correctly rounded float32 table (computed in float64), like pocketfft's.
#include "common.cuh"

namespace psa {

constexpr int kFftThreads = 512;
constexpr int64_t kMaxSmemPoints = 16384;

__global__ void twiddle_kernel(int64_t n, float2* __restrict__ tw) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s, c;
  sincospi(2.0 * (double)j / (double)n, &s, &c);
  tw[j] = make_float2((float)c, (float)(-s));
}

int launch_twiddles(int64_t n, float2* tw, cudaStream_t s) {
  if (n <= 0) return PSA_OK;
  twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, tw);
  return launch_status("twiddle_kernel");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// In-place forward DIF transform of s[0..m).  Result index p holds frequency digit_reverse(p).
__device__ void fft_dif_inplace(float2* __restrict__ s, int m, int log2m, const float2* __restrict__ tw, int tw_n) {
  int L = m;
  if (log2m & 1) {   // leading radix-2 pass
    const int half = L >> 1, tstep = tw_n / L;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      float2 a = s[j], b = s[j + half];
      s[j] = cadd(a, b);
      s[j + half] = cmul(csub(a, b), __ldg(tw + (int64_t)j * tstep));
    }
    L = half;
    __syncthreads();
  }
  for (; L >= 4; L >>= 2) {
    const int q = L >> 2, tstep = tw_n / L;
    for (int b = threadIdx.x; b < (m >> 2); b += blockDim.x) {
      const int j = b & (q - 1);
      const int base = (b - j) * 4 + j;            // (b / q) * L + j
      float2 a0 = s[base], a1 = s[base + q], a2 = s[base + 2 * q], a3 = s[base + 3 * q];
      float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
      float2 y0 = cadd(t0, t2), y1 = cadd(t1, t3), y2 = csub(t0, t2), y3 = csub(t1, t3);
      if (j != 0) {
        const int64_t w = (int64_t)j * tstep;
        y1 = cmul(y1, __ldg(tw + w));
        y2 = cmul(y2, __ldg(tw + 2 * w));
        y3 = cmul(y3, __ldg(tw + 3 * w));
      }
      s[base] = y0; s[base + q] = y1; s[base + 2 * q] = y2; s[base + 3 * q] = y3;
    }
    __syncthreads();
  }
}

// frequency held at in-place position p after fft_dif_inplace
__device__ __forceinline__ int dif_frequency(int p, int m, int log2m) {
  int f = 0, weight = 1, L = m;
  if (log2m & 1) {
    L >>= 1;
    f += (p / L) & 1;
    weight = 2;
  }
  for (; L >= 4; L >>= 2) {
    f += ((p / (L >> 2)) & 3) * weight;
    weight <<= 2;
  }
  return f;
}

// Load one column into shared memory, applying the radix-R split for sub-transform r (R == 1: plain copy).
__device__ void load_column(float2* __restrict__ s, const float* __restrict__ re, const float* __restrict__ im,
                            int m, int R, int r, const float2* __restrict__ tw, int n_t) {
  if (R == 1) {
    for (int t = threadIdx.x; t < m; t += blockDim.x) s[t] = make_float2(__ldg(re + t), __ldg(im + t));
    return;
  }
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < R; ++j) {
      float2 x = make_float2(__ldg(re + t + (int64_t)j * m), __ldg(im + t + (int64_t)j * m));
      int wi = (int)(((int64_t)j * r) % R) * m;           // w_R^{jr} = w_n^{(jr mod R) m}
      acc = cadd(acc, wi ? cmul(x, __ldg(tw + wi)) : x);
    }
    s[t] = r ? cmul(acc, __ldg(tw + (int64_t)t * r)) : acc;   // w_n^{tr}, t r < n
  }
}

template <int kMode>
__global__ void __launch_bounds__(kFftThreads) fft_sed_kernel(
    const float* __restrict__ P, int n_groups, int64_t group_stride, int n_k, int n_t, int64_t ldp,
    const float2* __restrict__ tw, void* __restrict__ out, int64_t n_k_total, int64_t k_offset, int m, int log2m,
    int R) {
  extern __shared__ float2 s_data[];
  const float n_f = (float)n_t;   // divide like the reference does (exact for powers of two anyway)

  if (kMode == PSA_MODE_COHERENT) {
    // block -> (k, pol, r)
    const int r = blockIdx.x % R;
    const int pol = (blockIdx.x / R) % 3;
    const int k = blockIdx.x / (3 * R);
    const float* re = P + ((int64_t)(2 * k) * 3 + pol) * ldp;
    const float* im = P + ((int64_t)(2 * k + 1) * 3 + pol) * ldp;
    load_column(s_data, re, im, m, R, r, tw, n_t);
    __syncthreads();
    fft_dif_inplace(s_data, m, log2m, tw, n_t);
    float2* o = reinterpret_cast<float2*>(out);
    for (int p = threadIdx.x; p < m; p += blockDim.x) {
      const int64_t f = (int64_t)dif_frequency(p, m, log2m) * R + r;
      float2 v = s_data[p];
      o[(f * n_k_total + k_offset + k) * 3 + pol] = make_float2(v.x / n_f, v.y / n_f);
    }
  } else {
    // block -> (k, r); loop over groups and polarisations, accumulate |S|^2 per in-place position
    float* s_acc = reinterpret_cast<float*>(s_data + m);
    const int r = blockIdx.x % R;
    const int k = blockIdx.x / R;
    for (int p = threadIdx.x; p < m; p += blockDim.x) s_acc[p] = 0.f;
    for (int g = 0; g < n_groups; ++g) {
      for (int pol = 0; pol < 3; ++pol) {
        const float* base = P + (int64_t)g * group_stride;
        const float* re = base + ((int64_t)(2 * k) * 3 + pol) * ldp;
        const float* im = base + ((int64_t)(2 * k + 1) * 3 + pol) * ldp;
        __syncthreads();
        load_column(s_data, re, im, m, R, r, tw, n_t);
        __syncthreads();
        fft_dif_inplace(s_data, m, log2m, tw, n_t);
        for (int p = threadIdx.x; p < m; p += blockDim.x) {
          float2 v = s_data[p];
          float vr = v.x / n_f, vi = v.y / n_f;
          s_acc[p] += vr * vr + vi * vi;
        }
      }
    }
    __syncthreads();
    float* o = reinterpret_cast<float*>(out);
    for (int p = threadIdx.x; p < m; p += blockDim.x) {
      const int64_t f = (int64_t)dif_frequency(p, m, log2m) * R + r;
      o[f * n_k_total + k_offset + k] = s_acc[p];
    }
  }
}

int launch_fft(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
               const float2* tw, int mode, void* out, int64_t n_k_total, int64_t k_offset, cudaStream_t s) {
  if (n_k == 0 || n_t == 0) return PSA_OK;
  PSA_REQUIRE(mode == PSA_MODE_COHERENT || mode == PSA_MODE_INCOHERENT, "psa_fft_sed: unknown mode %d", mode);
  if ((n_t & (n_t - 1)) != 0 || n_t < 16 || n_t > (int64_t)kMaxSmemPoints * 64) {
    set_error("psa_fft_sed: n_t=%lld is not a supported length (power of two, 16 <= n_t <= 2^20)", (long long)n_t);
    return PSA_ERR_UNSUPPORTED;
  }
  int R = 1;
  int64_t m = n_t;
  while (m > kMaxSmemPoints) { m >>= 1; R <<= 1; }
  int log2m = 0;
  while ((1 << log2m) < m) ++log2m;

  size_t smem = (size_t)m * sizeof(float2) + (mode == PSA_MODE_INCOHERENT ? (size_t)m * sizeof(float) : 0);
  if (mode == PSA_MODE_COHERENT) {
    PSA_CUDA(cudaFuncSetAttribute(fft_sed_kernel<PSA_MODE_COHERENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned blocks = (unsigned)(n_k * 3 * R);
    fft_sed_kernel<PSA_MODE_COHERENT><<<blocks, kFftThreads, smem, s>>>(P, (int)n_groups, group_stride, (int)n_k, (int)n_t,
                                                                      ldp, tw, out, n_k_total, k_offset, (int)m, log2m, R);
  } else {
    PSA_CUDA(cudaFuncSetAttribute(fft_sed_kernel<PSA_MODE_INCOHERENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned blocks = (unsigned)(n_k * R);
    fft_sed_kernel<PSA_MODE_INCOHERENT><<<blocks, kFftThreads, smem, s>>>(P, (int)n_groups, group_stride, (int)n_k, (int)n_t,
                                                                        ldp, tw, out, n_k_total, k_offset, (int)m, log2m, R);
  }
  return launch_status("fft_sed_kernel");
}

}  // namespace psa
