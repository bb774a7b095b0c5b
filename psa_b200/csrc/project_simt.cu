// CUDA-core (dp4a) projection kernel.  Same contract and bit-identical output as the tcgen05 kernel
// in project_tc2.cu: it is the on-device cross-check used by the parity tests and by bring-up of the
// tensor-core path (selected with PSA_PROJECT_SIMT); it is not a fallback - the Python layer always
// asks for PSA_PROJECT_TENSOR.
#include "project_common.cuh"

namespace psa {

constexpr int TM = 64, TN = 64, TKW = 16;   // tile rows, tile columns, K words (4 atoms each) per step
constexpr int LDW = TKW + 1;                // padded row length in words

__global__ void __launch_bounds__(256) project_simt_kernel(
    const int8_t* __restrict__ adig, int64_t rows, int64_t rows_alloc, const int8_t* __restrict__ bdig,
    const int32_t* __restrict__ expo, int64_t n_t, int64_t n_t_total, int64_t pitch, int64_t a_begin, int64_t a_end,
    int accumulate, float* __restrict__ P, int64_t ldp) {
  __shared__ uint32_t As[kSlices][TM][LDW];
  __shared__ uint32_t Bs[kSlices][TN][LDW];

  const int pol = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.y * TM;
  const int64_t t0 = (int64_t)blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t a_plane = rows_alloc * pitch;
  const int64_t b_plane = n_t_total * pitch;

  int32_t acc[kClasses][4][4] = {};

  for (int64_t a0 = a_begin; a0 < a_end; a0 += TKW * 4) {
    // cooperative tile load: 64 rows x 16 words per slice for A and for B
    for (int e = threadIdx.x; e < kSlices * TM * TKW; e += 256) {
      int kw = e % TKW, r = (e / TKW) % TM, sl = e / (TKW * TM);
      int64_t atom = a0 + kw * 4;
      uint32_t va = 0, vb = 0;
      if (atom < a_end) {   // a_end is a multiple of 64 or the padded pitch; planes are zero padded
        if (m0 + r < rows) va = __ldg(reinterpret_cast<const uint32_t*>(adig + sl * a_plane + (m0 + r) * pitch + atom));
        if (t0 + r < n_t)
          vb = __ldg(reinterpret_cast<const uint32_t*>(bdig + (int64_t)(pol * kSlices + sl) * b_plane + (t0 + r) * pitch + atom));
      }
      As[sl][r][kw] = va;
      Bs[sl][r][kw] = vb;
    }
    __syncthreads();
#pragma unroll 2
    for (int kw = 0; kw < TKW; ++kw) {
      int a[kSlices][4], b[kSlices][4];
#pragma unroll
      for (int sl = 0; sl < kSlices; ++sl)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[sl][i] = (int)As[sl][ty * 4 + i][kw];
          b[sl][i] = (int)Bs[sl][tx * 4 + i][kw];
        }
#pragma unroll
      for (int si = 0; si < kSlices; ++si)
#pragma unroll
        for (int sj = 0; sj < kSlices; ++sj) {
          if (si + sj < kMinClass) continue;
          const int c = si + sj - kMinClass;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][i][j] = __dp4a(a[si][i], b[sj][j], acc[c][i][j]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t t = t0 + tx * 4 + j;
      if (t >= n_t) continue;
      int e = __ldg(expo + pol * n_t_total + t);
      float v = combine_classes(acc[0][i][j], acc[1][i][j], acc[2][i][j], acc[3][i][j], e);
      float* dst = P + (m * 3 + pol) * ldp + t;
      *dst = accumulate ? __fadd_rn(*dst, v) : v;
    }
  }
}

int launch_project_simt(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                        const int32_t* expo, int64_t n_t, int64_t n_t_total, int64_t n_sel, int64_t pitch, float* P,
                        int64_t ldp, cudaStream_t s) {
  if (rows == 0 || n_t == 0) return PSA_OK;
  dim3 grid((unsigned)((n_t + TN - 1) / TN), (unsigned)((rows + TM - 1) / TM), 3);
  int pass = 0;
  for (int64_t a0 = 0; a0 < pitch && (a0 < n_sel || pass == 0); a0 += kMaxAtomsPerPass, ++pass) {
    int64_t a1 = a0 + kMaxAtomsPerPass < pitch ? a0 + kMaxAtomsPerPass : pitch;
    project_simt_kernel<<<grid, 256, 0, s>>>(adig, rows, rows_alloc, bdig, expo, n_t, n_t_total, pitch, a0, a1,
                                             pass > 0, P, ldp);
    int st = launch_status("project_simt_kernel");
    if (st != PSA_OK) return st;
  }
  return PSA_OK;
}

}  // namespace psa
