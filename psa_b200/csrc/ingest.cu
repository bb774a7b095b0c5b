// Trajectory ingest: float32 mean positions (bit-exact with NumPy) and the one-time split of the
// projected time series into int8 digit planes laid out for TMA / tcgen05.  Both are HBM-bound.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace psa {

// ---------------------------------------------------------------------------------------------
// Mean position.  np.mean(positions, axis=0, dtype=float32) reduces the OUTER axis of a
// C-contiguous array, i.e. one sequential float32 accumulation per (atom, xyz) column in frame
// order, followed by a float32 division (reference: sed_calculator.py:205; recipe verified in
// SURVEY.md appendix A).
// ---------------------------------------------------------------------------------------------
// The adds of one column are inherently serial, so bandwidth has to come from memory-level
// parallelism instead: a CTA owns kMeanCols adjacent columns; all 16 warps stream a tile of kMeanRows
// frames into registers (independent coalesced loads, issued while the previous tile is being
// consumed), park it in shared memory, and one warp per 32 columns then performs the ordered float32
// additions from shared memory.
#ifndef PSA_MEAN_COLS
#define PSA_MEAN_COLS 32
#endif
#ifndef PSA_MEAN_ROWS
#define PSA_MEAN_ROWS 256
#endif
constexpr int kMeanCols = PSA_MEAN_COLS;        // columns per CTA (multiple of 32): one consumer warp per 32
constexpr int kMeanRows = PSA_MEAN_ROWS;        // frames per tile
constexpr int kMeanThreads = 512;
constexpr int kMeanPerThread = kMeanRows * kMeanCols / kMeanThreads;   // loads in flight per thread
constexpr int kMeanConsumers = kMeanCols / 32;

// out = (acc_in ? acc_in : 0) (+) the rows in order, then / divisor when divisor > 0 (a multi-GPU ingest
// passes the running sums of a frame slice from rank to rank: the additions stay in frame order).
__global__ void __launch_bounds__(kMeanThreads) mean_positions_kernel(const float* __restrict__ pos, int64_t n_t,
                                                                      int64_t n_cols, const float* acc_in, float divisor,
                                                                      float* mean) {
  extern __shared__ float mean_smem[];
  float (*tile)[kMeanRows][kMeanCols] = reinterpret_cast<float (*)[kMeanRows][kMeanCols]>(mean_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t col0 = (int64_t)blockIdx.x * kMeanCols;
  const int64_t n_tiles = (n_t + kMeanRows - 1) / kMeanRows;
  // element e = threadIdx.x + i * kMeanThreads of a tile: row e / kMeanCols, column e % kMeanCols
  const int my_col = threadIdx.x % kMeanCols, my_row0 = threadIdx.x / kMeanCols;
  constexpr int kRowStep = kMeanThreads / kMeanCols;
  const bool load_ok = col0 + my_col < n_cols;
  const float* p = pos + col0 + my_col;

  float v[kMeanPerThread];
  auto fetch = [&](int64_t tile_idx) {
    const int64_t t0 = tile_idx * kMeanRows + my_row0;
#pragma unroll
    for (int i = 0; i < kMeanPerThread; ++i) {
      const int64_t t = t0 + (int64_t)i * kRowStep;
      v[i] = (load_ok && t < n_t) ? __ldg(p + t * n_cols) : 0.0f;
    }
  };

  float acc = 0.0f;
  if (acc_in != nullptr && warp < kMeanConsumers && col0 + warp * 32 + lane < n_cols) acc = acc_in[col0 + warp * 32 + lane];
  if (n_tiles > 0) fetch(0);
  for (int64_t it = 0; it < n_tiles; ++it) {
    const int buf = (int)(it & 1);
#pragma unroll
    for (int i = 0; i < kMeanPerThread; ++i) tile[buf][my_row0 + i * kRowStep][my_col] = v[i];
    __syncthreads();
    if (it + 1 < n_tiles) fetch(it + 1);              // in flight while the consumer warps add this tile
    if (warp < kMeanConsumers) {
      const int c = warp * 32 + lane;
      const int64_t rows = (n_t - it * kMeanRows) < kMeanRows ? (n_t - it * kMeanRows) : kMeanRows;
      if (rows == kMeanRows) {
#pragma unroll 16
        for (int r = 0; r < kMeanRows; ++r) acc = __fadd_rn(acc, tile[buf][r][c]);
      } else {
        for (int r = 0; r < (int)rows; ++r) acc = __fadd_rn(acc, tile[buf][r][c]);
      }
    }
    // the store into tile[buf] two iterations from now is ordered after this read by the
    // __syncthreads of the next iteration
  }
  if (warp < kMeanConsumers) {
    const int64_t col = col0 + warp * 32 + lane;
    if (col < n_cols) mean[col] = divisor > 0.f ? __fdiv_rn(acc, divisor) : acc;
  }
}

// TMA variant (the product path whenever the row pitch is a multiple of 16 bytes, i.e. n_atoms % 4 == 0).
// The register-staged kernel above spends 20 instructions per loaded word on addresses, predicates and
// the shared-memory hand-over (ncu: issue slots 60 % busy, DRAM 47 %).  Here the copy engine does all
// of that: one elected thread streams [kTmaRows x kTmaCols] boxes of the frame-major array into a ring
// of shared-memory stages (cp.async.bulk.tensor.2d + mbarrier transaction counts), and the only other
// threads of the CTA are the consumers - one per column - which perform the ordered float32 additions.
// A CTA keeps kTmaStages * 16 KiB in flight, three CTAs per SM (scripts/mean_tune.py: 64 x 64 x 4 stages).
#ifndef PSA_MEAN_TMA_COLS
#define PSA_MEAN_TMA_COLS 64
#endif
#ifndef PSA_MEAN_TMA_ROWS
#define PSA_MEAN_TMA_ROWS 64
#endif
#ifndef PSA_MEAN_TMA_STAGES
#define PSA_MEAN_TMA_STAGES 4
#endif
constexpr int kTmaCols = PSA_MEAN_TMA_COLS;      // columns per CTA (multiple of 32, <= 256)
constexpr int kTmaRows = PSA_MEAN_TMA_ROWS;      // frames per stage (<= 256)
constexpr int kTmaStages = PSA_MEAN_TMA_STAGES;
constexpr int kTmaConsumerWarps = kTmaCols / 32;
constexpr int kTmaThreads = (kTmaConsumerWarps + 1) * 32;
constexpr uint32_t kTmaStageBytes = kTmaCols * kTmaRows * sizeof(float);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(kTmaThreads) mean_positions_tma_kernel(const __grid_constant__ CUtensorMap map, int n_t,
                                                                         int64_t n_cols, const float* acc_in, float divisor,
                                                                         float* mean) {
  extern __shared__ __align__(128) float tma_tiles[];        // [stage][row][col]
  __shared__ uint64_t full_bar[kTmaStages], empty_bar[kTmaStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = blockIdx.x * kTmaCols;
  const int n_tiles = (n_t + kTmaRows - 1) / kTmaRows;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {                            // producer: one thread feeds the ring
    if (lane == 0) {
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % kTmaStages;
        mbar_wait(&empty_bar[s], ((i / kTmaStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kTmaStageBytes);         // rows past n_t are zero-filled and still counted
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                smem_addr(tma_tiles + (size_t)s * kTmaRows * kTmaCols)),
            "l"(&map), "r"(smem_addr(&full_bar[s])), "r"(col0), "r"(i * kTmaRows)
            : "memory");
      }
    }
    return;
  }

  const int c = warp * 32 + lane;
  float acc = (acc_in != nullptr && col0 + c < n_cols) ? acc_in[col0 + c] : 0.0f;
  for (int i = 0; i < n_tiles; ++i) {
    const int s = i % kTmaStages;
    mbar_wait(&full_bar[s], (i / kTmaStages) & 1);
    const float* tp = tma_tiles + (size_t)s * kTmaRows * kTmaCols + c;
    const int rows = min(kTmaRows, n_t - i * kTmaRows);
    if (rows == kTmaRows) {
#pragma unroll 16
      for (int r = 0; r < kTmaRows; ++r) acc = __fadd_rn(acc, tp[r * kTmaCols]);
    } else {
      for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, tp[r * kTmaCols]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);                // this warp is done reading the stage
  }
  if (col0 + c < n_cols) mean[col0 + c] = divisor > 0.f ? __fdiv_rn(acc, divisor) : acc;
}

static bool mean_use_tma(const float* pos, int64_t n_t, int64_t n_cols) {
  static const bool disabled = []() {
    const char* e = getenv("PSA_MEAN_IMPL");
    return e != nullptr && strcmp(e, "simt") == 0;
  }();
  return !disabled && n_cols % 4 == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 && n_t < (1 << 30) &&
         n_cols < ((int64_t)1 << 31) && tensor_map_encoder() != nullptr;
}

int launch_mean_positions(const float* pos, int64_t n_t, int64_t n_a, float* mean, cudaStream_t s) {
  return launch_mean_accumulate(pos, n_t, n_a, nullptr, n_t, mean, s);
}

int launch_mean_accumulate(const float* pos, int64_t n_t, int64_t n_a, const float* acc_in, int64_t divide_by, float* mean,
                           cudaStream_t s) {
  int64_t n_cols = n_a * 3;
  if (n_cols == 0) return PSA_OK;
  DeviceGuard guard(mean);
  const float divisor = divide_by > 0 ? (float)divide_by : 0.f;
  if (n_t > 0 && mean_use_tma(pos, n_t, n_cols)) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_t};
    cuuint64_t strides[1] = {(cuuint64_t)n_cols * sizeof(float)};
    cuuint32_t box[2] = {kTmaCols, kTmaRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(pos), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed with CUresult %d (mean positions, n_t=%lld n_cols=%lld)", (int)r,
                (long long)n_t, (long long)n_cols);
      return PSA_ERR_CUDA;
    }
    constexpr int smem = kTmaStages * (int)kTmaStageBytes;
    PSA_CUDA(cudaFuncSetAttribute(mean_positions_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t blocks = (n_cols + kTmaCols - 1) / kTmaCols;
    mean_positions_tma_kernel<<<(unsigned)blocks, kTmaThreads, smem, s>>>(map, (int)n_t, n_cols, acc_in, divisor, mean);
    return launch_status("mean_positions_tma_kernel");
  }
  int64_t blocks = (n_cols + kMeanCols - 1) / kMeanCols;
  constexpr int smem = 2 * kMeanRows * kMeanCols * (int)sizeof(float);   // 64 KiB: three CTAs per SM
  PSA_CUDA(cudaFuncSetAttribute(mean_positions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mean_positions_kernel<<<(unsigned)blocks, kMeanThreads, smem, s>>>(pos, n_t, n_cols, acc_in, divisor, mean);
  return launch_status("mean_positions_kernel");
}

// ---------------------------------------------------------------------------------------------
// Digitise.  One CTA per frame.  Pass 1 finds, per polarisation, the exponent e with
// max_a |x| < 2^e; pass 2 (the row is L2-hot) writes the balanced base-256 digits of
// rint(x * 2^(30-e)) as four int8 planes dig[pol][slice][t][atom].  A thread handles four
// consecutive atoms so that every plane store is a packed 32-bit word.
// ---------------------------------------------------------------------------------------------
// Values of four consecutive selected atoms (12 floats).  Without a gather list the 48 bytes are
// contiguous and, when 16-byte aligned, fetched as three float4 loads; otherwise scalar loads.
// kStaged: `row` is the copy of the frame in shared memory (plain loads), otherwise global memory (__ldg).
template <bool kStaged>
__device__ __forceinline__ float row_value(const float* p) { return kStaged ? *p : __ldg(p); }

// `weight` (or NULL): one float32 factor per atom, applied after the mean subtraction - the README facade's
// mass weighting sqrt(m) v (reference: README.md:83-101; the shipped source is unweighted, weight == NULL).
template <bool kStaged>
__device__ __forceinline__ void load_quad(const float* __restrict__ row, const float* __restrict__ mean,
                                          const float* __restrict__ weight, const int32_t* __restrict__ idx, int64_t j0,
                                          int64_t n_sel, float (&v)[12]) {
  const float* src = row + j0 * 3;
  if (idx == nullptr && j0 + 4 <= n_sel && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4 a, b, c;
    if (kStaged) { a = s4[0]; b = s4[1]; c = s4[2]; }
    else { a = __ldg(s4); b = __ldg(s4 + 1); c = __ldg(s4 + 2); }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
    v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    if (mean != nullptr) {
      const float* m = mean + j0 * 3;
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = __fsub_rn(v[i], __ldg(m + i));
    }
    if (weight != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float w = __ldg(weight + j0 + q);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[q * 3 + c] = __fmul_rn(v[q * 3 + c], w);
      }
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t j = j0 + q;
    if (j < n_sel) {
      const int64_t atom = idx ? (int64_t)__ldg(idx + j) : j;
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        float x = row_value<kStaged>(row + atom * 3 + p);
        if (mean != nullptr) x = __fsub_rn(x, __ldg(mean + atom * 3 + p));
        if (weight != nullptr) x = __fmul_rn(x, __ldg(weight + atom));
        v[q * 3 + p] = x;
      }
    } else {
      v[q * 3] = v[q * 3 + 1] = v[q * 3 + 2] = 0.f;
    }
  }
}

// Row maxima are tracked as the bit pattern of |x| (an unsigned integer max orders finite floats like fmaxf does,
// and - unlike fmaxf, which drops NaN - lets NaN and Inf win, so that a corrupt frame cannot pass silently).
__device__ __forceinline__ uint32_t abs_bits(float x) { return __float_as_uint(x) & 0x7fffffffu; }

// exponent e with m < 2^e (kExpMin for an all-zero row); kExpPoison when the row holds a NaN, an infinity or a value
// >= 2^kExpMax: the reference propagates those into the spectrum (sed_calculator.py:81-84), and so does the
// projection epilogue for a poisoned row.
__device__ __forceinline__ int row_exponent(uint32_t m_bits) {
  if (m_bits >= 0x7f800000u) return kExpPoison;
  const float m = __uint_as_float(m_bits);
  int e = kExpMin;
  if (m > 0.f) {
    frexpf(m, &e);                         // m = f * 2^e with f in [0.5,1)  =>  m < 2^e
    if (e > kExpMax) return kExpPoison;
    e = max(kExpMin, e);
  }
  return e;
}
__device__ __forceinline__ float row_scale(int e) {      // 2^(30 - e), exact; a poisoned row digitises as zeros
  return e == kExpPoison ? 0.f : exp2f((float)(kFracBits - e));
}

// Digit-plane words of four atoms.  Digits without a carry chain: with Y = X + 0x00808080 the low three bytes
// of Y are d_i + 128 and the top byte is d3, so the bytes of Z = Y ^ 0x00808080 ARE the balanced digits
// (identical to balanced_digits(); |X| <= 2^30 leaves room for the bias).  The four atoms are then turned
// into four plane words by a 4 x 4 byte transpose (eight PRMTs).
__device__ __forceinline__ void quad_words(const float (&v)[12], const float (&scale)[3], uint32_t (&word)[3][kSlices]) {
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    uint32_t z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      z[q] = ((uint32_t)__float2int_rn(v[q * 3 + p] * scale[p]) + 0x00808080u) ^ 0x00808080u;
    const uint32_t lo01 = __byte_perm(z[0], z[1], 0x5140), hi01 = __byte_perm(z[0], z[1], 0x7362);
    const uint32_t lo23 = __byte_perm(z[2], z[3], 0x5140), hi23 = __byte_perm(z[2], z[3], 0x7362);
    word[p][0] = __byte_perm(lo01, lo23, 0x5410);
    word[p][1] = __byte_perm(lo01, lo23, 0x7632);
    word[p][2] = __byte_perm(hi01, hi23, 0x5410);
    word[p][3] = __byte_perm(hi01, hi23, 0x7632);
  }
}

// kStaged (gathered selections whose frame row fits shared memory): the row is fetched once with one bulk
// copy (cp.async.bulk + mbarrier) and both passes gather from shared memory; scalar gathers from global memory
// kept the LSU/L1 busier than HBM (ncu: 172 M sector requests, issue slots 57 %, DRAM 61 %).
template <bool kStaged>
__global__ void __launch_bounds__(512) digitize_kernel(const float* __restrict__ data,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ weight,
                                                       const int32_t* __restrict__ idx, int64_t n_t,
                                                       int64_t n_a, int64_t n_sel, int64_t pitch,
                                                       DigDests dst, int64_t t0, bool compact) {
  // `data` holds the rows [t0, t0 + gridDim.x) of a trajectory of n_t frames; dig / expo describe all n_t frames
  const int64_t t = t0 + blockIdx.x;
  const float* row = data + (int64_t)blockIdx.x * n_a * 3;
  __shared__ uint32_t s_max[3][32];
  __shared__ int s_exp[3];
  if (kStaged) {
    extern __shared__ __align__(128) float s_row[];
    __shared__ uint64_t row_bar;
    const uint32_t bytes = (uint32_t)(n_a * 3 * sizeof(float));
    if (threadIdx.x == 0) {
      mbar_init(&row_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&row_bar, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(s_row)),
                   "l"(row), "r"(bytes), "r"(smem_addr(&row_bar))
                   : "memory");
    }
    __syncthreads();                                   // the barrier is initialised before anyone polls it
    mbar_wait(&row_bar, 0);
    row = s_row;
  }

  uint32_t mx[3] = {0u, 0u, 0u};
  if (kStaged && idx != nullptr && compact) {
    // Gathered selection (PSA_DIG_COMPACT=1; measured 6 % slower than gathering twice, so off by default - the bank
    // conflicts it removes were not what bounds the kernel): the maxima pass also COMPACTS the selected atoms (mean subtracted, weight applied) into a
    // second shared-memory array, one atom per thread (stride-3 writes: conflict-free).  The digit pass then reads
    // whole quads as three 16-byte loads.  Gathering quads straight from the staged row put 8 lanes on one bank for
    // the regular "every other group of four" selections of a two-sublattice crystal (ncu: 26.8 M conflicts).
    extern __shared__ __align__(128) float s_row2[];
    float* sel = s_row2 + ((n_a * 3 + 3) & ~(int64_t)3);
    for (int64_t j = threadIdx.x; j < n_sel; j += blockDim.x) {
      const int64_t atom = (int64_t)__ldg(idx + j);
      const float wt = weight != nullptr ? __ldg(weight + atom) : 1.f;
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        float x = row[atom * 3 + p];
        if (mean != nullptr) x = __fsub_rn(x, __ldg(mean + atom * 3 + p));
        if (weight != nullptr) x = __fmul_rn(x, wt);
        sel[j * 3 + p] = x;
        mx[p] = max(mx[p], abs_bits(x));
      }
    }
    row = sel;                                          // from here on: a contiguous row of n_sel atoms, ready to digitise
    mean = nullptr;
    weight = nullptr;
    idx = nullptr;
  } else if (idx == nullptr) {
    for (int64_t j0 = (int64_t)threadIdx.x * 4; j0 < n_sel; j0 += (int64_t)blockDim.x * 4) {
      float v[12];
      load_quad<kStaged>(row, mean, weight, idx, j0, n_sel, v);
#pragma unroll
      for (int i = 0; i < 12; ++i) mx[i % 3] = max(mx[i % 3], abs_bits(v[i]));
    }
  } else {   // gather: one selected atom per thread keeps neighbouring lanes on neighbouring atoms
    for (int64_t j = threadIdx.x; j < n_sel; j += blockDim.x) {
      const int64_t atom = (int64_t)__ldg(idx + j);
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        float x = row_value<kStaged>(row + atom * 3 + p);
        if (mean != nullptr) x = __fsub_rn(x, __ldg(mean + atom * 3 + p));
        if (weight != nullptr) x = __fmul_rn(x, __ldg(weight + atom));
        mx[p] = max(mx[p], abs_bits(x));
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    uint32_t m = mx[p];
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[p][threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t m = 0u;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = max(m, s_max[threadIdx.x][w]);
    const int e = row_exponent(m);
    s_exp[threadIdx.x] = e;
    for (int d = 0; d < dst.n; ++d) dst.expo[d][threadIdx.x * n_t + t] = e;
  }
  __syncthreads();

  float scale[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) scale[p] = row_scale(s_exp[p]);   // exact power of two

  const int64_t plane = n_t * pitch;                 // bytes of one (pol, slice) plane
  for (int64_t j0 = (int64_t)threadIdx.x * 4; j0 < pitch; j0 += (int64_t)blockDim.x * 4) {
    uint32_t word[3][kSlices] = {};
    if (j0 < n_sel) {
      float v[12];
      load_quad<kStaged>(row, mean, weight, idx, j0, n_sel, v);      // entries past n_sel come back as zeros -> zero digits
      quad_words(v, scale, word);
    }
    // every destination gets the row: this GPU's planes and, in a multi-GPU sliced ingest, the peers' planes through
    // NVLink (a warp stores 128 contiguous bytes per plane: one full-size NVLink write each)
    for (int d = 0; d < dst.n; ++d) {
      int8_t* dig = dst.dig[d];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int sl = 0; sl < kSlices; ++sl)
          *reinterpret_cast<uint32_t*>(dig + (int64_t)(p * kSlices + sl) * plane + t * pitch + j0) = word[p][sl];
    }
  }
}

// Contiguous rows (no gather list): a cluster of C = 1, 2, 4 or 8 CTAs owns one frame.  Each CTA fetches its share of
// the row (<= 100 KB, two CTAs per SM) with bulk copies into shared memory and produces the digits from the staged
// copy, so the frame is read from HBM exactly once; the per-polarisation maxima of the CTAs are exchanged through
// distributed shared memory.  The share arrives as kDigChunks bulk copies with their own mbarriers: the maxima pass
// starts on the first quarter while the rest is still in flight (waiting for the whole 96 KB slice of a 64 000-atom
// frame held the kernel at 67 % of the HBM roofline).
constexpr int kDigChunks = 4;

__global__ void __launch_bounds__(256)
digitize_cluster_kernel(const float* __restrict__ data, const float* __restrict__ mean,
                        const float* __restrict__ weight, int64_t n_t, int64_t n_a,
                        int64_t pitch, int slice_atoms, DigDests dst, int64_t t0) {
  extern __shared__ __align__(128) float s_row[];     // this CTA's slice of the frame
  __shared__ uint64_t row_bar[kDigChunks];
  __shared__ uint32_t s_max[3][8], s_loc[3];
  uint32_t rank, kDigCluster;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(kDigCluster));
  const int64_t frame = blockIdx.x / kDigCluster, t = t0 + frame;
  const int64_t a0 = (int64_t)rank * slice_atoms;
  const int n_loc = (int)max((int64_t)0, min(n_a, a0 + slice_atoms) - a0);     // multiple of 4, > 0
  const int chunk_atoms = ((n_loc + kDigChunks - 1) / kDigChunks + 3) / 4 * 4;  // whole quads: 48-byte multiples
  if (threadIdx.x == 0) {
    for (int c = 0; c < kDigChunks; ++c) mbar_init(&row_bar[c], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const float* src = data + (frame * n_a + a0) * 3;
    for (int c = 0; c < kDigChunks; ++c) {
      const int c0 = min(c * chunk_atoms, n_loc), c1 = min(c0 + chunk_atoms, n_loc);
      const uint32_t bytes = (uint32_t)(c1 - c0) * 3 * sizeof(float);
      mbar_expect_tx(&row_bar[c], bytes);
      if (bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(s_row + (size_t)c0 * 3)),
                     "l"(src + (size_t)c0 * 3), "r"(bytes), "r"(smem_addr(&row_bar[c]))
                     : "memory");
    }
  }
  __syncthreads();
  const float* lmean = mean != nullptr ? mean + a0 * 3 : nullptr;
  const float* lweight = weight != nullptr ? weight + a0 : nullptr;

  uint32_t mx[3] = {0u, 0u, 0u};
  for (int c = 0; c < kDigChunks; ++c) {
    const int c0 = min(c * chunk_atoms, n_loc), c1 = min(c0 + chunk_atoms, n_loc);
    mbar_wait(&row_bar[c], 0);
    for (int j0 = c0 + threadIdx.x * 4; j0 < c1; j0 += blockDim.x * 4) {
      float v[12];
      load_quad<true>(s_row, lmean, lweight, nullptr, j0, n_loc, v);
#pragma unroll
      for (int i = 0; i < 12; ++i) mx[i % 3] = max(mx[i % 3], abs_bits(v[i]));
    }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    uint32_t m = mx[p];
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[p][threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t m = 0u;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = max(m, s_max[threadIdx.x][w]);
    s_loc[threadIdx.x] = m;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  float scale[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {                        // every thread folds the CTAs' maxima (3 x cluster size remote reads)
    uint32_t m = 0u;
    for (uint32_t c = 0; c < kDigCluster; ++c) {
      uint32_t remote, val;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr(&s_loc[p])), "r"(c));
      asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(val) : "r"(remote) : "memory");
      m = max(m, val);
    }
    const int e = row_exponent(m);
    if (rank == 0 && threadIdx.x == 0)
      for (int d = 0; d < dst.n; ++d) dst.expo[d][p * n_t + t] = e;
    scale[p] = row_scale(e);
  }

  const int64_t plane = n_t * pitch;
  // this CTA's atoms, plus - on the last CTA - the zero padding up to the pitch
  const int64_t j_end = rank == kDigCluster - 1 ? pitch - a0 : n_loc;
  for (int64_t j0 = (int64_t)threadIdx.x * 4; j0 < j_end; j0 += (int64_t)blockDim.x * 4) {
    uint32_t word[3][kSlices] = {};
    if (j0 < n_loc) {
      float v[12];
      load_quad<true>(s_row, lmean, lweight, nullptr, j0, n_loc, v);
      quad_words(v, scale, word);
    }
    for (int d = 0; d < dst.n; ++d) {
      int8_t* dig = dst.dig[d];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int sl = 0; sl < kSlices; ++sl)
          *reinterpret_cast<uint32_t*>(dig + (int64_t)(p * kSlices + sl) * plane + t * pitch + a0 + j0) = word[p][sl];
    }
  }
  // nobody leaves while a peer may still read its maxima
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int launch_digitize(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_t,
                    int64_t n_a, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, cudaStream_t s) {
  DigDests dst{};
  dst.n = 1;
  dst.dig[0] = dig;
  dst.expo[0] = expo;
  return launch_digitize_rows(data, mean, weight, idx, n_t, n_a, n_sel, pitch, dst, n_t, 0, s, false);
}

int launch_digitize_rows(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                         int64_t n_a, int64_t n_sel, int64_t pitch, const DigDests& dst, int64_t n_t_total, int64_t t0,
                         cudaStream_t s, bool light) {
  if (n_rows == 0) return PSA_OK;
  DeviceGuard guard(data);
  if (light) {
    // A ring step of the multi-GPU exchange runs UNDER the projection kernel, which leaves 10 240 registers and
    // ~30 KB of shared memory per SM: 128-thread CTAs of the plain two-pass kernel (64 registers, no staging) fit
    // next to it; the row is L2-hot for the second pass and the step is NVLink-bound anyway.
    digitize_kernel<false><<<(unsigned)n_rows, 128, 0, s>>>(data, mean, weight, idx, n_t_total, n_a, n_sel, pitch, dst, t0,
                                                            false);
    return launch_status("digitize_kernel<light>");
  }
  const size_t row_bytes = (size_t)n_a * 3 * sizeof(float);
  static const bool no_stage = getenv("PSA_DIGITIZE_NO_STAGE") != nullptr;
  if (idx != nullptr && !no_stage && row_bytes <= 100 * 1024 && row_bytes % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(data) & 15) == 0) {            // gathered selection, row fits: stage it
    static const bool compact = getenv("PSA_DIG_COMPACT") != nullptr && atoi(getenv("PSA_DIG_COMPACT")) != 0;
    const size_t staged_bytes = ((row_bytes + 15) & ~(size_t)15) + (compact ? (size_t)((n_sel + 3) / 4 * 4) * 3 * sizeof(float) : 0);
    PSA_CUDA(cudaFuncSetAttribute(digitize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_bytes));
    digitize_kernel<true><<<(unsigned)n_rows, 256, staged_bytes, s>>>(data, mean, weight, idx, n_t_total, n_a, n_sel, pitch,
                                                                     dst, t0, compact);
    return launch_status("digitize_kernel<staged>");
  }
  // Rows longer than 200 KB (64 000 atoms: 768 KB) take the cluster kernel - the two-pass kernel below would read them
  // from HBM twice.  Shorter rows stay in L1/L2 between the two passes of the plain kernel, which measured faster than
  // staging them (13 824 atoms: 1.07 vs 1.16 ms; 4096 atoms: equal); PSA_DIG_CLUSTER=1|2|4|8 forces a cluster size.
  static const int forced_cluster = getenv("PSA_DIG_CLUSTER") ? atoi(getenv("PSA_DIG_CLUSTER")) : 0;
  if (idx == nullptr && !no_stage && n_a % 4 == 0 && n_a >= 64 && (reinterpret_cast<uintptr_t>(data) & 15) == 0 &&
      (row_bytes > 200 * 1024 || forced_cluster)) {
    const int forced = forced_cluster;
    for (int C : {1, 2, 4, 8}) {
      if (forced && C != forced) continue;
      const int slice_atoms = (int)(((n_a + C - 1) / C + 3) / 4 * 4);
      const size_t slice_bytes = (size_t)slice_atoms * 3 * sizeof(float);
      if (slice_bytes > 100u * 1024 || (int64_t)slice_atoms * (C - 1) >= n_a) continue;
      PSA_CUDA(cudaFuncSetAttribute(digitize_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_bytes));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(n_rows * C));
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = slice_bytes;
      cfg.stream = s;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = (unsigned)C;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      PSA_CUDA(cudaLaunchKernelEx(&cfg, digitize_cluster_kernel, data, mean, weight, n_t_total, n_a, pitch, slice_atoms, dst,
                                  t0));
      return launch_status("digitize_cluster_kernel");
    }
  }
  // 512 threads for contiguous rows, 256 for gathered ones (bench.py on C1 / C2: 0.131 vs 0.142 ms, 0.242 vs 0.252 ms)
  digitize_kernel<false><<<(unsigned)n_rows, idx == nullptr ? 512 : 256, 0, s>>>(data, mean, weight, idx, n_t_total, n_a,
                                                                                   n_sel, pitch, dst, t0, false);
  return launch_status("digitize_kernel");
}

}  // namespace psa
