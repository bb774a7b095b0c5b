// Projection on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   P[row][pol][t] = sum_atoms phase[row][atom] * data[t][atom][pol]
//
// Both operands arrive as four int8 digit planes (balanced base 256, see common.cuh), so the
// float32-grade contraction becomes ten exact s8 x s8 -> s32 tensor-core products per K step:
// every digit pair (i, j) with i + j >= 3, accumulated per class i + j in its own TMEM columns.
// The epilogue recombines the four int32 class sums in int64 and rounds ONCE to float32, so the
// result does not depend on tile shapes, K order or clock - it is reproducible bit for bit by
// the integer model in tests/.
//
// Tile: M = 128 frames of one polarisation (TMEM lanes) x N <= 128 phase rows (TMEM columns; a row
// is cos or sin of one k-point).  Frames are the M side because n_t is always a large multiple of
// 128 while the number of rows is ragged: N is rounded up to 16 per tile, so a 400-row k-path costs
// 3 x 128 + 16 columns rather than 4 x 128, the row exponent is one value per thread, and a warp's
// store of one output row is a coalesced 128-byte line.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0 lane 0 : TMA producer  - 2 bulk tensor copies per stage (trajectory digits 4x128x64 B,
//                   phase digits 4x128x64 B), 64-byte swizzle, 3-stage mbarrier ring
//   warp 1 lane 0 : MMA issuer    - 20 tcgen05.mma.kind::i8 (M128 N<=128 K32) per stage into
//                   4 x 128 TMEM columns, tcgen05.commit releases the stage / publishes the tile
//   warps 2..9    : epilogue      - two warps per TMEM lane quarter (each takes half the columns):
//                   tcgen05.ld 32x32b, int64 recombination, float32 store
// TMEM is fully used by the four accumulator classes (4 x 128 columns), so the epilogue of a tile
// is not overlapped with the next tile's MMAs; the producer does keep prefetching through it.
#include <cuda.h>

#include "project_common.cuh"
#include "tma.cuh"

namespace psa {
namespace tc {

constexpr int BM = 128;            // frames per tile == TMEM lanes
constexpr int BN = 128;            // phase rows (2 per k-point) per tile == TMEM columns per class
constexpr int BK = 64;             // atoms per stage == one 64-byte swizzle row
constexpr int UMMA_K = 32;         // atoms per tcgen05.mma.kind::i8
constexpr int STAGES = 3;
constexpr int SLICE_BYTES = 128 * BK;                  // one digit plane of a tile
constexpr int OPERAND_BYTES = kSlices * SLICE_BYTES;   // 32 KiB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;         // 64 KiB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // + alignment slack
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One lane of a fully converged warp.  ptxas only emits straight-line UTCIMMA / UTMALDG code when the
// single issuing thread is chosen with elect.sync; under a plain `lane == 0` branch it wraps every
// uniform-datapath instruction in an ELECT/branch loop that costs more than the MMA itself.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// MMA with descriptors given as (low word, shared high word): the low word is linear in the smem
// address, so stepping through slices / K steps is one 32-bit add per operand.
__device__ __forceinline__ void tc_mma_i8_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n"
      "}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kDescHi = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);   // SBO, version 1, SWIZZLE_64B

#define PSA_TMEM_LD16(r, addr)                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(addr))

// Shared-memory matrix descriptors (K-major operand, 64-byte swizzle: rows of 64 B, 8-row groups 512 B
// apart = SBO, descriptor version 1 for sm_100, layout type 4 = SWIZZLE_64B) are built from desc_lo /
// kDescHi above.
// Instruction descriptor: D = s32, A = B = s8, both K-major, M = 128, N = n (multiple of 16).
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TileCoord {
  int r_tile, pol, t_tile, row0, n_cols;   // row0: first phase row; n_cols: rows in this tile, rounded up to 16
};
// Phase rows are spread evenly over the row tiles (each a multiple of 16 columns, at most 128): a
// 400-row k-path becomes 4 x 112 columns instead of 3 x 128 + 16 - a 16-column tile costs almost as
// much as a full one because the M-side operand read does not shrink with N.
__device__ __forceinline__ TileCoord decode_tile(int tile, int r_tiles, int t_tiles, int rows) {
  TileCoord c;
  // Row tiles fastest (concurrent CTAs share the same trajectory strip), rotated by the strip index: the last
  // row tile is the narrow one, and with a fixed order a CTA whose stride is even in r_tiles would only ever
  // see wide tiles (C2: 74 CTA pairs, 4 row tiles - half the pairs did 11 wide tiles, the others mixed).
  int n = tile / r_tiles;
  c.r_tile = (tile + n) % r_tiles;
  c.pol = n / t_tiles;
  c.t_tile = n % t_tiles;
  const int per_tile = (((rows + r_tiles - 1) / r_tiles) + 15) & ~15;
  c.row0 = c.r_tile * per_tile;
  int left = rows - c.row0;
  left = left < 0 ? 0 : left;
  c.n_cols = left >= per_tile ? per_tile : ((left + 15) & ~15);
  return c;
}

__global__ void __launch_bounds__(THREADS, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const int32_t* __restrict__ expo, float* __restrict__ P, int rows, int n_t, int64_t ldp,
                  int a_begin, int a_end, int accumulate, int r_tiles, int t_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 32 * EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int total_tiles = r_tiles * t_tiles * 3;
  const int num_kb = (a_end - a_begin + BK - 1) / BK;

  if (warp == 0) {                                             // ---------------- TMA producer (one elected lane)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const uint32_t dst = smem_u32(smem + stage * STAGE_BYTES);
          const int atom0 = a_begin + kb * BK;
          tma_load_3d(dst, &tmap_b, &full_bar[stage], atom0, tc.t_tile * BM, tc.pol * kSlices);   // M side
          tma_load_3d(dst + OPERAND_BYTES, &tmap_a, &full_bar[stage], atom0, tc.row0, 0);            // N side
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {                                      // ---------------- MMA issuer (one elected lane)
    int stage = 0;
    uint32_t phase = 0, tile_phase = 0;
    const uint32_t a_lo0 = desc_lo(smem_u32(smem));            // stage 0, slice 0, K step 0 of the M-side operand
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
      const uint32_t idesc = make_idesc(tc.n_cols);
      mbar_wait(tmem_empty, tile_phase ^ 1);                   // epilogue has drained the accumulators
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((stage * STAGE_BYTES) >> 4);   // trajectory digits (M side)
          const uint32_t b_lo = a_lo + (uint32_t)(OPERAND_BYTES >> 4);            // phase digits (N side)
          const uint32_t first = kb > 0 ? 1u : 0u;
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
#pragma unroll
            for (int si = 0; si < kSlices; ++si) {
#pragma unroll
              for (int sj = 0; sj < kSlices; ++sj) {
                if (si + sj < kMinClass) continue;
                const uint32_t d = tmem_base + (uint32_t)((si + sj - kMinClass) * BN);
                // the first product issued into each class of a tile (ks == 0, sj == 3) overwrites
                const uint32_t acc = (ks > 0 || sj != kSlices - 1) ? 1u : first;
                tc_mma_i8_lohi(d, a_lo + (uint32_t)((si * SLICE_BYTES + ks * UMMA_K) >> 4),
                               b_lo + (uint32_t)((sj * SLICE_BYTES + ks * UMMA_K) >> 4), kDescHi, idesc, acc);
              }
            }
          }
          tc_commit(&empty_bar[stage]);                        // stage reusable once these MMAs retire
          if (kb == num_kb - 1) tc_commit(tmem_full);          // accumulators complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tile_phase ^= 1;
    }
  } else {                                                     // ---------------- epilogue warps 2..9
    const int quarter = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                          // which 64 columns of each class it drains
    uint32_t tile_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
      mbar_wait(tmem_full, tile_phase);
      tc_fence_after();
      const int t = tc.t_tile * BM + quarter * 32 + lane;      // this thread's frame
      const bool t_ok = t < n_t;
      const int e = t_ok ? __ldg(expo + (int64_t)tc.pol * n_t + t) : kExpMin;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int c_end = min(tc.n_cols, half * 64 + 64);
#pragma unroll 1
      for (int c0 = half * 64; c0 < c_end; c0 += 16) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        PSA_TMEM_LD16(r0, lane_addr + 0 * BN + c0);
        PSA_TMEM_LD16(r1, lane_addr + 1 * BN + c0);
        PSA_TMEM_LD16(r2, lane_addr + 2 * BN + c0);
        PSA_TMEM_LD16(r3, lane_addr + 3 * BN + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (t_ok) {
          const int row0 = tc.row0 + c0;
          float* dst = P + ((int64_t)row0 * 3 + tc.pol) * ldp + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (row0 + i < rows) {
              float v = combine_classes((int32_t)r0[i], (int32_t)r1[i], (int32_t)r2[i], (int32_t)r3[i], e);
              float* d = dst + (int64_t)i * 3 * ldp;             // a warp writes one 128-byte line per row
              *d = accumulate ? __fadd_rn(*d, v) : v;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      tile_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------
static EncodeTiledFn encode_fn() { return tensor_map_encoder(); }

// 3-D map over int8 digit planes [outer][mid][n_sel], row pitch `pitch` bytes, box 64 x 128 x 4.
static int make_map(CUtensorMap* map, const int8_t* base, int64_t n_sel, int64_t pitch, int64_t mid, int64_t mid_alloc,
                    int64_t outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return PSA_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)n_sel, (cuuint64_t)mid, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(mid_alloc * pitch)};
  cuuint32_t box[3] = {BK, 128, kSlices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (n_sel=%lld pitch=%lld mid=%lld outer=%lld)", (int)r,
              (long long)n_sel, (long long)pitch, (long long)mid, (long long)outer);
    return PSA_ERR_CUDA;
  }
  return PSA_OK;
}

}  // namespace tc

int launch_project_tc(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                      const int32_t* expo, int64_t n_t, int64_t n_sel, int64_t pitch, float* P, int64_t ldp,
                      cudaStream_t s) {
  using namespace tc;
  if (rows == 0 || n_t == 0) return PSA_OK;
  PSA_REQUIRE(n_sel > 0, "psa_project: empty atom selection");
  PSA_REQUIRE(rows < (1 << 30) && n_t < (1 << 30) && n_sel < (1 << 30), "psa_project: extent too large");
  CUtensorMap map_a, map_b;
  int st = make_map(&map_a, adig, n_sel, pitch, rows, rows_alloc, kSlices);
  if (st != PSA_OK) return st;
  st = make_map(&map_b, bdig, n_sel, pitch, n_t, n_t, 3 * kSlices);
  if (st != PSA_OK) return st;

  static bool attr_set = false;   // per-process, idempotent
  if (!attr_set) {
    PSA_CUDA(cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  int dev = 0, sms = 0;
  PSA_CUDA(cudaGetDevice(&dev));
  PSA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int r_tiles = (int)((rows + BN - 1) / BN);
  const int t_tiles = (int)((n_t + BM - 1) / BM);
  const int total = r_tiles * t_tiles * 3;
  const int grid = total < sms ? total : sms;

  int pass = 0;
  for (int64_t a0 = 0; a0 < n_sel; a0 += kMaxAtomsPerPass, ++pass) {
    int64_t a1 = a0 + kMaxAtomsPerPass < n_sel ? a0 + kMaxAtomsPerPass : n_sel;
    project_tc_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, expo, P, (int)rows, (int)n_t, ldp, (int)a0,
                                                        (int)a1, pass > 0, r_tiles, t_tiles);
    st = launch_status("project_tc_kernel");
    if (st != PSA_OK) return st;
  }
  return PSA_OK;
}

}  // namespace psa
