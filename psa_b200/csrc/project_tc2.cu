// Projection on the tensor cores (tcgen05 cta_group::2, sm_100a):
//   P[row][pol][t] = sum_atoms phase[row][atom] * series[pol][t][atom]     (reference: sed_calculator.py:80-81)
// as an exact integer contraction of int8 digit planes (common.cuh), rounded once to float32.
//
// Two CTAs of a cluster (one TPC) compute M = 256 frames x N <= 128 phase rows together.  Each CTA
// stages its own 128 frames of trajectory digits but only HALF of the phase-digit tile (64 rows); the
// pair's MMA reads both halves.  Shared-memory traffic per MMA is (128 + 64) operand rows per CTA, the
// stage is 48 KiB (4 stages).  A single-CTA cta_group::1 kernel (round 1, removed) was bound by
// tensor-core shared-memory reads (ncu: l1tex tc wavefronts at 100 % in the MMA phase with the tensor
// pipe at ~70 %) and 5-8 % slower.
//
// Protocol (rank 0 = leader of the pair):
//   producers (warp 0 lane 0 of BOTH CTAs)   TMA into their own smem with .cta_group::2, completing
//                                            on the LEADER's full barrier (peer bit cleared)
//   MMA issuer (warp 1 lane 0 of the leader) tcgen05.mma.cta_group::2 M256; tcgen05.commit multicast
//                                            to both CTAs' empty / tmem_full barriers
//   epilogue (warps 2..17 of BOTH CTAs)      drain their own TMEM half, then arrive on the leader's
//                                            tmem_empty barrier (remote arrive from rank 1)
#include <cuda.h>

#include "project_common.cuh"
#include "tma.cuh"

namespace psa {
namespace tc2 {

constexpr int BM = 128;            // frames per CTA == TMEM lanes (256 per pair)
constexpr int BN = 128;            // phase rows per tile == TMEM columns per class
constexpr int BNH = BN / 2;        // phase rows staged by each CTA
constexpr int BK = 64;
constexpr int UMMA_K = 32;
constexpr int STAGES = 4;
constexpr int A_SLICE_BYTES = BM * BK;                   // 8 KiB
constexpr int B_SLICE_BYTES = BNH * BK;                  // 4 KiB
constexpr int A_BYTES = kSlices * A_SLICE_BYTES;         // 32 KiB
constexpr int B_BYTES = kSlices * B_SLICE_BYTES;         // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;           // 48 KiB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
#ifndef PSA_EPI_WARPS
#define PSA_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = PSA_EPI_WARPS;      // 4 TMEM lane quarters x (EPI_WARPS / 4) column parts
constexpr int EPI_COLS = BN / (EPI_WARPS / 4); // columns of each class drained by one warp
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;            // shared::cluster address of the same offset in the even CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `target` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t target) {
  asm volatile(
      "{\n"
      ".reg .b32 rem;\n"
      "mapa.shared::cluster.u32 rem, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(target)
      : "memory");
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
// TMA into this CTA's smem, transaction bytes reported to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// completion of all prior MMAs of this thread -> arrive on `bar` in both CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
// elect.sync (instead of lane == 0) keeps the issue code straight-line
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
__device__ __forceinline__ void tc_mma_i8_pair_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %4, p;\n"
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

// D = s32, A = B = s8, K-major, M = 256 (pair), N = n (multiple of 16)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

struct TileCoord {
  int r_tile, pol, t_tile, row0, n_cols;
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

// Where the projection rows go.  n == 0: all of them into P (row r at P + r * 3 * ldp).  n > 0 (frame-sharded multi-GPU
// run): rows [begin[q], begin[q + 1]) belong to destination q - another rank's projection buffer, mapped through CUDA
// IPC - and land there as its rows 0, 1, ...: ONE launch projects this rank's frames for every owner's k-points.
struct RowRoute {
  float* base[kMaxRouteDests];
  int begin[kMaxRouteDests + 1];
  int n;
};

template <bool kRouted>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
project_tc2_kernel(const __grid_constant__ CUtensorMap tmap_phase, const __grid_constant__ CUtensorMap tmap_traj,
                   const int32_t* __restrict__ expo, float* __restrict__ P, const __grid_constant__ RowRoute route,
                   int rows, int n_t, int64_t expo_stride, int64_t ldp, int a_begin, int a_end, int accumulate,
                   int r_tiles, int t_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);          // leader's producer arrives once with the pair's byte count
      mbar_init(&empty_bar[s], 1);         // one multicast commit from the leader's MMA thread
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 2 * 32 * EPI_WARPS);   // every epilogue thread of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_phase) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_traj) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                          // peer barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int total_tiles = r_tiles * t_tiles * 3;
  const int num_kb = (a_end - a_begin + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {                                             // ---------------- TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t dst = smem_u32(smem + stage * STAGE_BYTES);
          const int atom0 = a_begin + kb * BK;
          tma_load_3d_pair(dst, &tmap_traj, &full_bar[stage], atom0, tc.t_tile * 2 * BM + (int)rank * BM, tc.pol * kSlices);
          // the pair's N = n_cols columns are split in halves: this CTA stages rows [rank * n_cols/2, +n_cols/2)
          tma_load_3d_pair(dst + A_BYTES, &tmap_phase, &full_bar[stage], atom0, tc.row0 + (int)rank * (tc.n_cols >> 1), 0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {                                              // ---------------- MMA issuer (leader CTA, one elected lane)
      int stage = 0;
      uint32_t phase = 0, tile_phase = 0;
      const uint32_t a_lo0 = desc_lo(smem_u32(smem));
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
        const uint32_t idesc = make_idesc(tc.n_cols);
        mbar_wait(tmem_empty, tile_phase ^ 1);                 // both CTAs have drained their accumulators
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);                  // both CTAs' bytes have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)((stage * STAGE_BYTES) >> 4);   // trajectory digits (M side)
            const uint32_t b_lo = a_lo + (uint32_t)(A_BYTES >> 4);                  // this CTA's half of the phase digits
            const uint32_t first = kb > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
#pragma unroll
              for (int si = 0; si < kSlices; ++si) {
#pragma unroll
                for (int sj = 0; sj < kSlices; ++sj) {
                  if (si + sj < kMinClass) continue;
                  const uint32_t d = tmem_base + (uint32_t)((si + sj - kMinClass) * BN);
                  const uint32_t acc = (ks > 0 || sj != kSlices - 1) ? 1u : first;
                  tc_mma_i8_pair_lohi(d, a_lo + (uint32_t)((si * A_SLICE_BYTES + ks * UMMA_K) >> 4),
                                      b_lo + (uint32_t)((sj * B_SLICE_BYTES + ks * UMMA_K) >> 4), kDescHi, idesc, acc);
                }
              }
            }
            tc_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) tc_commit_pair(tmem_full);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tile_phase ^= 1;
      }
    }
  } else {                                                     // ---------------- epilogue warps 2..17 (both CTAs)
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    uint32_t tile_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const TileCoord tc = decode_tile(tile, r_tiles, t_tiles, rows);
      mbar_wait(tmem_full, tile_phase);
      tc_fence_after();
      const int t = tc.t_tile * 2 * BM + (int)rank * BM + quarter * 32 + lane;
      const bool t_ok = t < n_t;
      const int e = t_ok ? __ldg(expo + (int64_t)tc.pol * expo_stride + t) : kExpMin;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int c_begin = part * EPI_COLS, c_end = min(tc.n_cols, c_begin + EPI_COLS);
      // Drain first, store later: the accumulators are read and recombined into EPI_COLS float32 values
      // per thread, the TMEM is handed back to the MMA issuer, and only then do the (slow) global stores
      // run - under the next tile's MMAs instead of in front of them.
      float v[EPI_COLS];
#pragma unroll
      for (int ch = 0; ch < EPI_COLS / 16; ++ch) {
        const int c0 = c_begin + ch * 16;
        if (c0 < c_end) {                                      // warp-uniform
          uint32_t r0[16], r1[16], r2[16], r3[16];
          PSA_TMEM_LD16(r0, lane_addr + 0 * BN + c0);
          PSA_TMEM_LD16(r1, lane_addr + 1 * BN + c0);
          PSA_TMEM_LD16(r2, lane_addr + 2 * BN + c0);
          PSA_TMEM_LD16(r3, lane_addr + 3 * BN + c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[ch * 16 + i] = combine_classes((int32_t)r0[i], (int32_t)r1[i], (int32_t)r2[i], (int32_t)r3[i], e);
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty, 0);                      // leader's barrier, remote for rank 1
      if (t_ok) {
#pragma unroll
        for (int ch = 0; ch < EPI_COLS / 16; ++ch) {
          const int row0 = tc.row0 + c_begin + ch * 16;
          if (c_begin + ch * 16 < c_end) {
            if constexpr (!kRouted) {
              float* dst = P + ((int64_t)row0 * 3 + tc.pol) * ldp + t;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (row0 + i < rows) {
                  float* d = dst + (int64_t)i * 3 * ldp;
                  *d = accumulate ? __fadd_rn(*d, v[ch * 16 + i]) : v[ch * 16 + i];
                }
              }
            } else {                                           // per-row destination (warp-uniform lookups)
              int q = 0;
              while (q + 1 < route.n && row0 >= route.begin[q + 1]) ++q;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int row = row0 + i;
                if (row < rows) {
                  while (q + 1 < route.n && row >= route.begin[q + 1]) ++q;
                  float* d = route.base[q] + ((int64_t)(row - route.begin[q]) * 3 + tc.pol) * ldp + t;
                  *d = accumulate ? __fadd_rn(*d, v[ch * 16 + i]) : v[ch * 16 + i];
                }
              }
            }
          }
        }
      }
      tile_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                          // nobody exits while the peer may still touch its smem / barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static EncodeTiledFn encode_fn() { return tensor_map_encoder(); }

// 3-D map over int8 digit planes [outer][mid][n_sel], row pitch `pitch` bytes, box 64 x box_rows x 4.
static int make_map(CUtensorMap* map, const int8_t* base, int64_t n_sel, int64_t pitch, int64_t mid, int64_t mid_alloc,
                    int64_t outer, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return PSA_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)n_sel, (cuuint64_t)mid, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(mid_alloc * pitch)};
  cuuint32_t box[3] = {BK, (cuuint32_t)box_rows, kSlices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PSA_ERR_CUDA;
  }
  return PSA_OK;
}

}  // namespace tc2

// bdig / expo / P point at the first frame of the range to project; n_t frames of a trajectory of n_t_total frames
int launch_project_tc2(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                       const int32_t* expo, int64_t n_t, int64_t n_t_total, int64_t n_sel, int64_t pitch, float* P,
                       int64_t ldp, cudaStream_t s, float* const* dests, const int64_t* row_begin, int n_dest) {
  using namespace tc2;
  if (rows == 0 || n_t == 0) return PSA_OK;
  RowRoute route;
  memset(&route, 0, sizeof(route));
  if (n_dest > 0) {
    PSA_REQUIRE(n_dest <= kMaxRouteDests && dests != nullptr && row_begin != nullptr,
                "psa_project_routed: 1 to %d destinations", kMaxRouteDests);
    PSA_REQUIRE(row_begin[0] == 0 && row_begin[n_dest] == rows, "psa_project_routed: row ranges must cover [0, rows)");
    for (int q = 0; q < n_dest; ++q) {
      PSA_REQUIRE(row_begin[q + 1] >= row_begin[q], "psa_project_routed: row ranges must be ascending");
      PSA_REQUIRE(dests[q] != nullptr || row_begin[q + 1] == row_begin[q], "psa_project_routed: null destination %d", q);
      route.base[q] = dests[q];
      route.begin[q] = (int)row_begin[q];
    }
    route.begin[n_dest] = (int)rows;
    route.n = n_dest;
  }
  DeviceGuard guard(adig);
  PSA_REQUIRE(n_sel > 0, "psa_project: empty atom selection");
  PSA_REQUIRE(rows < (1 << 30) && n_t < (1 << 30) && n_sel < (1 << 30), "psa_project: extent too large");
  CUtensorMap map_phase, map_traj;
  int st = make_map(&map_phase, adig, n_sel, pitch, rows, rows_alloc, kSlices, BNH);
  if (st != PSA_OK) return st;
  st = make_map(&map_traj, bdig, n_sel, pitch, n_t, n_t_total, 3 * kSlices, BM);
  if (st != PSA_OK) return st;

  // per device and per context: set on every launch (a process may drive several GPUs from several threads)
  PSA_CUDA(cudaFuncSetAttribute(project_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  PSA_CUDA(cudaFuncSetAttribute(project_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int dev = 0, sms = 0;
  PSA_CUDA(cudaGetDevice(&dev));
  PSA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int r_tiles = (int)((rows + BN - 1) / BN);
  const int t_tiles = (int)((n_t + 2 * BM - 1) / (2 * BM));
  const int total = r_tiles * t_tiles * 3;
  const int clusters = total < sms / 2 ? total : sms / 2;

  int pass = 0;
  for (int64_t a0 = 0; a0 < n_sel; a0 += kMaxAtomsPerPass, ++pass) {
    int64_t a1 = a0 + kMaxAtomsPerPass < n_sel ? a0 + kMaxAtomsPerPass : n_sel;
    if (route.n > 0)
      project_tc2_kernel<true><<<2 * clusters, THREADS, SMEM_BYTES, s>>>(map_phase, map_traj, expo, P, route, (int)rows, (int)n_t,
                                                                         n_t_total, ldp, (int)a0, (int)a1, pass > 0, r_tiles, t_tiles);
    else
      project_tc2_kernel<false><<<2 * clusters, THREADS, SMEM_BYTES, s>>>(map_phase, map_traj, expo, P, route, (int)rows, (int)n_t,
                                                                          n_t_total, ldp, (int)a0, (int)a1, pass > 0, r_tiles, t_tiles);
    st = launch_status("project_tc2_kernel");
    if (st != PSA_OK) return st;
  }
  return PSA_OK;
}

}  // namespace psa
