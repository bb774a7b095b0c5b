// Phase table exp(+i k.r) for one k-chunk, written directly as int8 digit planes.
//
// The reference forms theta = k.r in float32 (OpenBLAS sgemm with K = 3, i.e. an FMA chain) and
// then takes the complex64 exponential (reference: sed_calculator.py:78).  Both steps are
// replicated: the same FMA chain, then cos/sin evaluated in float64 and rounded once to float32,
// which is the correctly rounded float32 value (|theta| reaches hundreds of radians, so fast-math
// intrinsics are out).  The float32 cos/sin is then scaled by 2^30 (exact) and split into balanced
// base-256 digits.  The table for a chunk (8 bytes per (k, atom)) is scratch: it is sized to stay
// L2-resident and is consumed by the projection kernel through TMA.
#include "common.cuh"

namespace psa {

__global__ void __launch_bounds__(256) phase_digits_kernel(const float* __restrict__ kvecs, int64_t n_k,
                                                           const float* __restrict__ mean,
                                                           const int32_t* __restrict__ idx, int64_t n_sel,
                                                           int64_t pitch, int64_t rows_alloc,
                                                           int8_t* __restrict__ adig) {
  const int64_t k = blockIdx.y;
  const float k0 = __ldg(kvecs + k * 3 + 0), k1 = __ldg(kvecs + k * 3 + 1), k2 = __ldg(kvecs + k * 3 + 2);
  const int64_t plane = rows_alloc * pitch;
  const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= pitch) return;

  uint32_t wc[kSlices] = {}, ws[kSlices] = {};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int64_t j = j0 + q;
    if (j < n_sel) {
      int64_t atom = idx ? (int64_t)__ldg(idx + j) : j;
      float r0 = __ldg(mean + atom * 3 + 0), r1 = __ldg(mean + atom * 3 + 1), r2 = __ldg(mean + atom * 3 + 2);
      float theta = __fmaf_rn(k2, r2, __fmaf_rn(k1, r1, __fmul_rn(k0, r0)));
      double sd, cd;
      sincos((double)theta, &sd, &cd);
      float cf = (float)cd, sf = (float)sd;
      int8_t dc[kSlices], ds[kSlices];
      balanced_digits(__float2int_rn(cf * 1073741824.0f), dc);
      balanced_digits(__float2int_rn(sf * 1073741824.0f), ds);
#pragma unroll
      for (int sl = 0; sl < kSlices; ++sl) {
        wc[sl] |= (uint32_t)(uint8_t)dc[sl] << (8 * q);
        ws[sl] |= (uint32_t)(uint8_t)ds[sl] << (8 * q);
      }
    }
  }
#pragma unroll
  for (int sl = 0; sl < kSlices; ++sl) {
    *reinterpret_cast<uint32_t*>(adig + sl * plane + (2 * k) * pitch + j0) = wc[sl];
    *reinterpret_cast<uint32_t*>(adig + sl * plane + (2 * k + 1) * pitch + j0) = ws[sl];
  }
}

int launch_phase_digits(const float* kvecs, int64_t n_k, const float* mean, const int32_t* idx,
                        int64_t n_sel, int64_t pitch, int64_t rows_alloc, int8_t* adig, cudaStream_t s) {
  if (n_k == 0 || pitch == 0) return PSA_OK;
  dim3 grid((unsigned)((pitch / 4 + 255) / 256), (unsigned)n_k);
  phase_digits_kernel<<<grid, 256, 0, s>>>(kvecs, n_k, mean, idx, n_sel, pitch, rows_alloc, adig);
  return launch_status("phase_digits_kernel");
}

}  // namespace psa
