// Time-axis FFT of the projected columns, fused with the 1/n_t scale and the SED assembly.
//
// One CTA transforms one column (k, pol [, group]) entirely in shared memory and writes the spectrum
// straight into the result layout, so spectra never round-trip through HBM:
//   coherent   : complex64 out[f][k][pol]                   (reference: sed_calculator.py:296-311)
//   incoherent : float32  out[f][k] = sum_g sum_pol |S|^2    (reference: sed_calculator.py:313-327)
//
// Transform structure (forward, decimation in frequency, in place, m = 2^s points, 32 <= m <= 16384):
//   * shared-memory passes: one radix-2 pass if s-5 is odd, then radix-4 passes down to blocks of 32.
//     Every butterfly leg is >= 32 elements away from the next, so a warp always touches 32
//     consecutive elements: conflict-free.
//   * final stage: each thread pulls one contiguous 32-point block into registers, finishes it with
//     radix 4 x 4 x 2 and stores the 32 results directly to global memory (digit-reversed frequency
//     index).  The array is padded by one element per 32 (index p lives at p + p/32), which makes the
//     per-thread contiguous block reads conflict-free as well.
// Columns longer than 16384 points do not fit one CTA's shared memory: they are split by a radix-R
// decimation-in-frequency step applied while loading, giving R independent sub-transforms that
// produce the frequencies f = R f' + r.
// Twiddles come from a correctly rounded float32 table (computed in float64), like pocketfft's.
#include "common.cuh"

namespace psa {

constexpr int kFftThreads = 512;
constexpr int64_t kMaxSmemPoints = 16384;
constexpr int kBlk = 32;   // points finished in registers per thread

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
__device__ __forceinline__ int phys(int p) { return p + (p >> 5); }

// forward radix-4 DIF butterfly on four legs (no twiddles)
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// Shared-memory passes: reduce the m-point problem to m/32 independent contiguous 32-point blocks.
__device__ void fft_smem_passes(float2* __restrict__ s, int m, int log2m, const float2* __restrict__ tw, int tw_n) {
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
    const int q = L >> 2, tstep = tw_n / L;
#pragma unroll 4
    for (int b = threadIdx.x; b < (m >> 2); b += blockDim.x) {
      const int j = b & (q - 1);
      const int base = ((b - j) << 2) + j;            // (b / q) * L + j
      const int i0 = phys(base), i1 = phys(base + q), i2 = phys(base + 2 * q), i3 = phys(base + 3 * q);
      float2 a0 = s[i0], a1 = s[i1], a2 = s[i2], a3 = s[i3];
      bfly4(a0, a1, a2, a3);
      if (j != 0) {
        const int64_t w = (int64_t)j * tstep;
        a1 = cmul(a1, __ldg(tw + w));
        a2 = cmul(a2, __ldg(tw + 2 * w));
        a3 = cmul(a3, __ldg(tw + 3 * w));
      }
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

// Load one column into shared memory, applying the radix-R split for sub-transform r (R == 1: plain copy).
__device__ void load_column(float2* __restrict__ s, const float* __restrict__ re, const float* __restrict__ im,
                            int m, int R, int r, const float2* __restrict__ tw) {
  if (R == 1) {
#pragma unroll 8
    for (int t = threadIdx.x; t < m; t += blockDim.x) s[phys(t)] = make_float2(__ldg(re + t), __ldg(im + t));
    return;
  }
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < R; ++j) {
      float2 x = make_float2(__ldg(re + t + (int64_t)j * m), __ldg(im + t + (int64_t)j * m));
      int wi = (int)(((int64_t)j * r) % R) * m;           // w_R^{jr} = w_n^{(jr mod R) m}
      acc = cadd(acc, wi ? cmul(x, __ldg(tw + wi)) : x);
    }
    s[phys(t)] = r ? cmul(acc, __ldg(tw + (int64_t)t * r)) : acc;   // w_n^{tr}, t r < n
  }
}

template <int kMode>
__global__ void __launch_bounds__(kFftThreads) fft_sed_kernel(
    const float* __restrict__ P, int n_groups, int64_t group_stride, int n_k, int n_t, int64_t ldp,
    const float2* __restrict__ tw, void* __restrict__ out, int64_t n_k_total, int64_t k_offset, int m, int log2m,
    int R) {
  extern __shared__ float2 s_data[];
  const float inv_n = 1.0f / (float)n_t;   // n_t is a power of two: the product equals the reference's division
  const int n_blocks = m >> 5;
  const int fstep = m >> 5;                // frequency step between register elements with rev(e) = 1

  if (kMode == PSA_MODE_COHERENT) {
    // block -> (k, pol, r)
    const int r = blockIdx.x % R;
    const int pol = (blockIdx.x / R) % 3;
    const int k = blockIdx.x / (3 * R);
    const float* re = P + ((int64_t)(2 * k) * 3 + pol) * ldp;
    const float* im = P + ((int64_t)(2 * k + 1) * 3 + pol) * ldp;
    load_column(s_data, re, im, m, R, r, tw);
    __syncthreads();
    fft_smem_passes(s_data, m, log2m, tw, n_t);
    float2* o = reinterpret_cast<float2*>(out) + (k_offset + k) * 3 + pol;
    const int64_t fstride = n_k_total * 3;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
      float2 x[kBlk];
#pragma unroll
      for (int e = 0; e < kBlk; ++e) x[e] = s_data[b * (kBlk + 1) + e];
      fft32_registers(x, tw, n_t);
      const int f0 = block_base_frequency(b, log2m);
#pragma unroll
      for (int e = 0; e < kBlk; ++e) {
        const int rev = (e >> 3) + 4 * ((e >> 1) & 3) + 16 * (e & 1);
        const int64_t f = (int64_t)(f0 + fstep * rev) * R + r;
        o[f * fstride] = make_float2(x[e].x * inv_n, x[e].y * inv_n);
      }
    }
  } else {
    // block -> (k, r); loop over groups and polarisations, accumulate |S|^2 per (padded) position
    float* s_acc = reinterpret_cast<float*>(s_data + m + (m >> 5));
    const int r = blockIdx.x % R;
    const int k = blockIdx.x / R;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x)
#pragma unroll
      for (int e = 0; e < kBlk; ++e) s_acc[b * (kBlk + 1) + e] = 0.f;
    for (int g = 0; g < n_groups; ++g) {
      for (int pol = 0; pol < 3; ++pol) {
        const float* base = P + (int64_t)g * group_stride;
        const float* re = base + ((int64_t)(2 * k) * 3 + pol) * ldp;
        const float* im = base + ((int64_t)(2 * k + 1) * 3 + pol) * ldp;
        __syncthreads();
        load_column(s_data, re, im, m, R, r, tw);
        __syncthreads();
        fft_smem_passes(s_data, m, log2m, tw, n_t);
        for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
          float2 x[kBlk];
#pragma unroll
          for (int e = 0; e < kBlk; ++e) x[e] = s_data[b * (kBlk + 1) + e];
          fft32_registers(x, tw, n_t);
#pragma unroll
          for (int e = 0; e < kBlk; ++e) {
            const float vr = x[e].x * inv_n, vi = x[e].y * inv_n;
            s_acc[b * (kBlk + 1) + e] += vr * vr + vi * vi;     // same thread owns this slot every time
          }
        }
      }
    }
    float* o = reinterpret_cast<float*>(out) + k_offset + k;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
      const int f0 = block_base_frequency(b, log2m);
#pragma unroll
      for (int e = 0; e < kBlk; ++e) {
        const int rev = (e >> 3) + 4 * ((e >> 1) & 3) + 16 * (e & 1);
        const int64_t f = (int64_t)(f0 + fstep * rev) * R + r;
        o[f * n_k_total] = s_acc[b * (kBlk + 1) + e];
      }
    }
  }
}

int launch_fft(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
               const float2* tw, int mode, void* out, int64_t n_k_total, int64_t k_offset, cudaStream_t s) {
  if (n_k == 0 || n_t == 0) return PSA_OK;
  PSA_REQUIRE(mode == PSA_MODE_COHERENT || mode == PSA_MODE_INCOHERENT, "psa_fft_sed: unknown mode %d", mode);
  if ((n_t & (n_t - 1)) != 0 || n_t < kBlk || n_t > (int64_t)kMaxSmemPoints * 64) {
    set_error("psa_fft_sed: n_t=%lld is not a supported length (power of two, 32 <= n_t <= 2^20)", (long long)n_t);
    return PSA_ERR_UNSUPPORTED;
  }
  int R = 1;
  int64_t m = n_t;
  while (m > kMaxSmemPoints) { m >>= 1; R <<= 1; }
  int log2m = 0;
  while ((1 << log2m) < m) ++log2m;

  const size_t padded = (size_t)(m + (m >> 5));
  size_t smem = padded * sizeof(float2) + (mode == PSA_MODE_INCOHERENT ? padded * sizeof(float) : 0);
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
