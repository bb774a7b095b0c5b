// Trajectory ingest: float32 mean positions (bit-exact with NumPy) and the one-time split of the
// projected time series into int8 digit planes laid out for TMA / tcgen05.  Both are HBM-bound.
#include "common.cuh"

namespace psa {

// ---------------------------------------------------------------------------------------------
// Mean position.  np.mean(positions, axis=0, dtype=float32) reduces the OUTER axis of a
// C-contiguous array, i.e. one sequential float32 accumulation per (atom, xyz) column in frame
// order, followed by a float32 division (reference: sed_calculator.py:205; recipe verified in
// SURVEY.md appendix A).  One thread owns one column; loads are issued in independent batches so
// only the adds are serial.
// ---------------------------------------------------------------------------------------------
constexpr int kMeanBatch = 16;

__global__ void __launch_bounds__(128) mean_positions_kernel(const float* __restrict__ pos, int64_t n_t,
                                                             int64_t n_cols, float* __restrict__ mean) {
  int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  const float* p = pos + col;
  float acc = 0.0f;
  int64_t t = 0;
  for (; t + kMeanBatch <= n_t; t += kMeanBatch) {
    float v[kMeanBatch];
#pragma unroll
    for (int i = 0; i < kMeanBatch; ++i) v[i] = __ldg(p + (t + i) * n_cols);
#pragma unroll
    for (int i = 0; i < kMeanBatch; ++i) acc = __fadd_rn(acc, v[i]);
  }
  for (; t < n_t; ++t) acc = __fadd_rn(acc, __ldg(p + t * n_cols));
  mean[col] = __fdiv_rn(acc, (float)n_t);
}

int launch_mean_positions(const float* pos, int64_t n_t, int64_t n_a, float* mean, cudaStream_t s) {
  int64_t n_cols = n_a * 3;
  if (n_cols == 0) return PSA_OK;
  int64_t blocks = (n_cols + 127) / 128;
  mean_positions_kernel<<<(unsigned)blocks, 128, 0, s>>>(pos, n_t, n_cols, mean);
  return launch_status("mean_positions_kernel");
}

// ---------------------------------------------------------------------------------------------
// Digitise.  One CTA per frame.  Pass 1 finds, per polarisation, the exponent e with
// max_a |x| < 2^e; pass 2 (the row is L2-hot) writes the balanced base-256 digits of
// rint(x * 2^(30-e)) as four int8 planes dig[pol][slice][t][atom].  A thread handles four
// consecutive atoms so that every plane store is a packed 32-bit word.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_value(const float* __restrict__ row, const float* __restrict__ mean,
                                            int64_t atom, int pol) {
  float v = __ldg(row + atom * 3 + pol);
  if (mean != nullptr) v = __fsub_rn(v, __ldg(mean + atom * 3 + pol));
  return v;
}

__global__ void __launch_bounds__(256) digitize_kernel(const float* __restrict__ data,
                                                       const float* __restrict__ mean,
                                                       const int32_t* __restrict__ idx, int64_t n_t,
                                                       int64_t n_a, int64_t n_sel, int64_t pitch,
                                                       int8_t* __restrict__ dig, int32_t* __restrict__ expo) {
  const int64_t t = blockIdx.x;
  const float* row = data + t * n_a * 3;
  __shared__ float s_max[3][8];
  __shared__ int s_exp[3];

  float mx[3] = {0.f, 0.f, 0.f};
  for (int64_t j = threadIdx.x; j < n_sel; j += blockDim.x) {
    int64_t atom = idx ? (int64_t)__ldg(idx + j) : j;
#pragma unroll
    for (int p = 0; p < 3; ++p) mx[p] = fmaxf(mx[p], fabsf(load_value(row, mean, atom, p)));
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    float m = mx[p];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[p][threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[threadIdx.x][w]);
    int e = kExpMin;
    if (m > 0.f && isfinite(m)) {
      frexpf(m, &e);                       // m = f * 2^e with f in [0.5,1)  =>  m < 2^e
      e = max(kExpMin, min(kExpMax, e));
    }
    s_exp[threadIdx.x] = e;
    expo[threadIdx.x * n_t + t] = e;
  }
  __syncthreads();

  float scale[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) scale[p] = exp2f((float)(kFracBits - s_exp[p]));   // exact power of two

  const int64_t plane = n_t * pitch;                 // bytes of one (pol, slice) plane
  for (int64_t j0 = (int64_t)threadIdx.x * 4; j0 < pitch; j0 += (int64_t)blockDim.x * 4) {
    uint32_t word[3][kSlices] = {};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int64_t j = j0 + q;
      if (j < n_sel) {
        int64_t atom = idx ? (int64_t)__ldg(idx + j) : j;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          float v = load_value(row, mean, atom, p);
          int8_t d[kSlices];
          balanced_digits(__float2int_rn(v * scale[p]), d);
#pragma unroll
          for (int sl = 0; sl < kSlices; ++sl) word[p][sl] |= (uint32_t)(uint8_t)d[sl] << (8 * q);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int sl = 0; sl < kSlices; ++sl)
        *reinterpret_cast<uint32_t*>(dig + (int64_t)(p * kSlices + sl) * plane + t * pitch + j0) = word[p][sl];
  }
}

int launch_digitize(const float* data, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                    int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, cudaStream_t s) {
  if (n_t == 0) return PSA_OK;
  digitize_kernel<<<(unsigned)n_t, 256, 0, s>>>(data, mean, idx, n_t, n_a, n_sel, pitch, dig, expo);
  return launch_status("digitize_kernel");
}

}  // namespace psa
