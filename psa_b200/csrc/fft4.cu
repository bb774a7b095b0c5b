// Four-step time FFT + coherent assembly for long power-of-two columns (n_t = 8192, 16384, 32768) - the
// frame counts of every BASELINE config.  Replaces, for those lengths, the one-CTA-per-(column, residue)
// kernel of fft.cu, whose 8-byte scattered result stores, 4x redundant column reads and ~190 instructions per
// point held it at 13 % of the HBM roofline (profiles/r01g_ncu_full_c2.md).
//
// ONE persistent kernel runs both stages of the four-step scheme (fft4.cuh) as tiles of 4096 points handed out
// through an atomic counter, in an order that software-pipelines the two stages over groups of 16 adjacent
// (k, pol) columns:
//
//     phase p:   A-tiles of group p   (P -> 128-point transforms over n2 -> twiddle -> Y, float64)
//                B-tiles of group p - LAG   (Y -> N1-point transforms over n1 -> / n_t -> result)
//
// Y lives in a ring of RING group slots (16 columns x n_t x 16 bytes each, 20-40 MB in total): it is written
// and read back while still resident in the 126 MB L2, so HBM sees each input sample and each output value
// once.  Per-group completion counters order the stages: a B-tile waits until the 16 columns of its group are
// in Y, an A-tile that reuses a ring slot waits until the slot's previous group has been read.  Tiles are taken
// in increasing order and a tile only ever waits for tiles with smaller numbers, which are running or done -
// the grid needs no co-residency guarantee beyond "a CTA that took a tile is running".
//
// Stage B stores the reference's complex64 (n_f, n_k, 3) layout as 128-byte runs (16 columns of one frequency).
#include <cuda.h>

#include "common.cuh"
#include "fft4.cuh"

namespace psa {
namespace fft4 {

__constant__ double2 c_w128[128];            // w_128^e

struct Args {
  const float* P;
  int64_t ldp;
  const float* window;
  const double2* tw;          // w_n^e, e < n (start of the FFT plan)
  const double2* t1;          // [16][N1]: w_n^(n1 s)
  const double2* t2;          // [N1]: w_n^(16 n1)
  int prefetch;               // L2 prefetch of the next stage-A tile's samples
  double2* ybuf;              // RING slots of 16 columns x n
  int* counter;               // next tile
  int* done_a;                // [n_groups] stage-A tiles finished per group
  int* done_b;                // [n_groups] stage-B tiles whose Y reads are finished
  float2* out;                // already offset to column k_offset * 3
  int64_t fstride;            // elements between consecutive frequencies of the result (n_k_total * 3)
  int n_cols, n_groups, lag, ring, total_items;
  double inv_n;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_count(const int* p, int want) {   // one thread spins; the CTA joins at the barrier after
  while (ld_acquire(p) < want) __nanosleep(64);
}

struct LoadP {                 // one sample of a column of P, widened (and tapered) in float64
  const float* re;
  const float* im;
  const float* win;
  __device__ __forceinline__ c2 operator()(int t) const {
    c2 v = mk((double)__ldcs(re + t), (double)__ldcs(im + t));       // read once: evict first, Y should own the L2
    if (win != nullptr) {
      const double w = (double)__ldg(win + t);
      v.x *= w;
      v.y *= w;
    }
    return v;
  }
};

template <int N1>
struct LoadY {                 // Y[c][k2_0 + k2l][n1] of the tile's group; L2 only (other SMs wrote it in this launch)
  const double2* y_group;
  int k2_0;
  __device__ __forceinline__ c2 operator()(int tau, int n1) const {
    const int c = tau & 15, k2l = tau >> 4;
    const double2 v = __ldcg(y_group + ((int64_t)c * kN2 + k2_0 + k2l) * N1 + n1);
    return mk(v.x, v.y);
  }
};

template <int N1>
struct StoreOut {
  float2* out;                 // offset to the group's first column
  int64_t fstride;
  int k2_0, cols_live;
  double inv_n;
  __device__ __forceinline__ void operator()(int c, int k2l, int k1, c2 v) const {
    if (c < cols_live)
      __stcs(out + (int64_t)(k2_0 + k2l + kN2 * k1) * fstride + c, make_float2((float)(v.x * inv_n), (float)(v.y * inv_n)));
  }
};

// OCC = CTAs per SM the register allocation is capped for: the butterflies hold 16 complex128 values (64 registers)
// and ptxas fits the whole tile loop in 80-96 registers with a few bytes of spill.  Five 128-thread CTAs per SM
// measured best (more CTAs shrink the L1 next to their shared memory; fewer leave latency exposed).
template <int N1, int T, int OCC>
__global__ void __launch_bounds__(T, OCC) fft4_kernel(Args a) {
  using G = Geo<N1, T>;
  constexpr int kThreads = T, kN1Tile = G::w;
  extern __shared__ double2 f4_smem[];
  c2* exch = reinterpret_cast<c2*>(f4_smem);
  c2* twb = exch + G::exchange_elems;                     // w_N1^(j s), rows of 17
  __shared__ int s_item[2];
  const int tid = threadIdx.x;
  for (int i = tid; i < G::q * 17; i += kThreads) {
    const int j = i / 17, s = i % 17;
    const double2 w = __ldg(a.tw + (s < 16 ? kN2 * j * s : 0));
    twb[i] = mk(w.x, w.y);
  }
  const c2* w128 = reinterpret_cast<const c2*>(c_w128);
  const c2* t1 = reinterpret_cast<const c2*>(a.t1);
  const c2* t2 = reinterpret_cast<const c2*>(a.t2);
  constexpr int TPG = G::tiles_per_group;

  // Tiles come from an atomic counter; the NEXT tile is taken while the current one is being processed, so the
  // counter's round trip and the first touch of the next stage-A tile's input (an L2 prefetch) hide under the
  // butterflies.  (Still deadlock-free: every CTA takes tiles in increasing order and works them off in order, so
  // the smallest unfinished tile is always some CTA's current tile, and it only waits for smaller ones.)
  if (tid == 0) s_item[0] = atomicAdd(a.counter, 1);
  __syncthreads();
  int item = s_item[0];
  for (int it = 0; item < a.total_items; ++it) {
    if (tid == 0) s_item[(it + 1) & 1] = atomicAdd(a.counter, 1);
    const int phase = item / (2 * TPG), r = item % (2 * TPG);
    if (r < TPG) {                                         // ---------------- stage A tile
      const int g = phase;
      if (g < a.n_groups) {
        const int c_local = r / G::a_tiles_per_column, n1_0 = (r % G::a_tiles_per_column) * kN1Tile;
        const int col = g * kColsPerGroup + c_local;
        if (g >= a.ring && tid == 0) wait_count(a.done_b + g - a.ring, TPG);   // the slot's previous group has been read
        if (col < a.n_cols) {
          const int k = col / 3, pol = col % 3;
          LoadP load{a.P + ((int64_t)(2 * k) * 3 + pol) * a.ldp, a.P + ((int64_t)(2 * k + 1) * 3 + pol) * a.ldp, a.window};
          c2* y_col = reinterpret_cast<c2*>(a.ybuf) + ((int64_t)(g % a.ring) * kColsPerGroup + c_local) * G::n;
          stage_a_pass1<N1, T>(tid, n1_0, load, w128, exch);       // loads first: they do not depend on the ring slot
          __syncthreads();                                 // exchange complete; tid 0 has seen the slot free
          stage_a_pass2<N1, T>(tid, n1_0, exch, t1, t2, y_col);
        }
        __syncthreads();                                   // every thread's Y stores precede the barrier ...
        if (tid == 0) {
          __threadfence();                                 // ... and become visible GPU-wide before the count goes up
          atomicAdd(a.done_a + g, 1);
        }
      }
    } else {                                               // ---------------- stage B tile
      const int g = phase - a.lag;
      if (g >= 0 && g < a.n_groups) {
        const int k2_0 = (r - TPG) * G::k2_per_tile;
        if (tid == 0) wait_count(a.done_a + g, TPG);       // all 16 columns of the group are in Y
        __syncthreads();
        LoadY<N1> load_y{a.ybuf + (int64_t)(g % a.ring) * kColsPerGroup * G::n, k2_0};
        stage_b_pass1<N1, T>(tid, load_y, twb, exch);
        __syncthreads();                                   // Y of this tile is in registers / shared memory
        if (tid == 0) atomicAdd(a.done_b + g, 1);
        StoreOut<N1> sink{a.out + (int64_t)g * kColsPerGroup, a.fstride, k2_0, min(kColsPerGroup, a.n_cols - g * kColsPerGroup),
                          a.inv_n};
        stage_b_pass2<N1, T>(tid, exch, sink);
      }
    }
    __syncthreads();                                       // this tile's shared-memory reads are finished; s_item is set
    item = s_item[(it + 1) & 1];
    // first touch of the next stage-A tile's samples: pull its lines into L2 while nothing depends on them yet
    if (a.prefetch && item < a.total_items && item % (2 * TPG) < TPG) {
      const int rn = item % (2 * TPG), gn = item / (2 * TPG);
      const int col = gn * kColsPerGroup + rn / G::a_tiles_per_column;
      if (gn < a.n_groups && col < a.n_cols) {
        const int k = col / 3, pol = col % 3, n1_0 = (rn % G::a_tiles_per_column) * kN1Tile;
        const float* re = a.P + ((int64_t)(2 * k) * 3 + pol) * a.ldp + n1_0;
        const float* im = a.P + ((int64_t)(2 * k + 1) * 3 + pol) * a.ldp + n1_0;
        for (int n2 = tid; n2 < 2 * kN2; n2 += T) {
          const float* ptr = (n2 < kN2 ? re : im) + (int64_t)(n2 & (kN2 - 1)) * N1;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
      }
    }
  }
}

static int n1_of(int64_t n_t) {
  if (n_t == 8192) return 64;
  if (n_t == 16384) return 128;
  if (n_t == 32768) return 256;
  return 0;
}

template <int N1, int T>
static size_t smem_bytes() { return (size_t)(Geo<N1, T>::exchange_elems + Geo<N1, T>::q * 17) * sizeof(double2); }

// threads per CTA: 128 where a 2048-point stage-B tile still spans 16 columns, else 256; CTAs per SM: 5 x 128 or
// 2 x 256 threads (PSA_FFT4_THREADS / PSA_FFT4_OCC select the other compiled variants for A/B runs)
static int threads_of(int n1) {
  static const int forced = getenv("PSA_FFT4_THREADS") ? atoi(getenv("PSA_FFT4_THREADS")) : 0;
  if (n1 == 256) return 256;
  return forced == 256 ? 256 : 128;
}
static int occ_of(int threads) {
  static const int forced = getenv("PSA_FFT4_OCC") ? atoi(getenv("PSA_FFT4_OCC")) : 0;
  if (threads == 128) return (forced == 4 || forced == 6) ? forced : 5;     // 16384 x 1024 k: 0.431 / 0.404 / 0.490 ms for 4 / 5 / 6
  return forced == 3 ? 3 : 2;                                               // 32768 x 256 k: 0.247 / 0.266 ms for 2 / 3
}

struct Schedule {
  int n_groups, lag, ring, resident, tpg;
  int64_t ctl_bytes, y_bytes;
};

static Schedule make_schedule(int64_t n_t, int64_t n_cols, int sms) {
  Schedule s;
  const int n1 = n1_of(n_t), threads = threads_of(n1);
  const int tpg = kColsPerGroup * n1 / (threads / 8);       // Geo<N1, T>::tiles_per_group
  s.tpg = tpg;
  s.n_groups = (int)((n_cols + kColsPerGroup - 1) / kColsPerGroup);
  s.resident = occ_of(threads) * sms;
  // Tile numbering: phase p = [A-tiles of group p | B-tiles of group p - lag].  Between the last A-tile of a group
  // and its first B-tile lie lag * 2 tpg tiles; between the last B-tile of a group and the first A-tile that reuses
  // its ring slot, (ring - lag - 1) * 2 tpg.  Both distances exceed the number of resident CTAs, so that a tile
  // normally finds what it waits for already complete.
  const int need = s.resident + tpg / 2;
  s.lag = 1;
  while (s.lag * 2 * tpg < need) ++s.lag;
  s.ring = s.lag + 2;
  while ((s.ring - s.lag - 1) * 2 * tpg < need) ++s.ring;
  if (s.ring > s.n_groups) s.ring = s.n_groups > 0 ? s.n_groups : 1;     // never more slots than groups
  s.ctl_bytes = round_up((int64_t)(1 + 2 * (int64_t)s.n_groups) * sizeof(int), 256);
  s.y_bytes = (int64_t)s.ring * kColsPerGroup * n_t * sizeof(double2);
  return s;
}

}  // namespace fft4

// Plan layout for these lengths: [ w_n (n) | the one-CTA kernel's pass tables (incoherent assembly still uses it) |
// t1 (16 N1) | t2 (N1) ]; fft.cu places the last two behind its own tables.
int64_t fft4_table_entries(int64_t n_t) { return 17 * (int64_t)fft4::n1_of(n_t); }

__global__ void fft4_tables_kernel(int n, int n1_len, double2* __restrict__ t1, double2* __restrict__ t2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 17 * n1_len) return;
  const int s = i / n1_len, n1 = i % n1_len;               // s == 16: the t2 row, exponent 16 n1
  double sn, cs;
  sincospi(2.0 * (double)((int64_t)n1 * s) / (double)n, &sn, &cs);
  (s < 16 ? t1 + s * n1_len : t2)[n1] = make_double2(cs, -sn);
}

int launch_fft4_tables(int64_t n_t, double2* t1, cudaStream_t s) {
  const int n1 = fft4::n1_of(n_t);
  fft4_tables_kernel<<<(17 * n1 + 255) / 256, 256, 0, s>>>((int)n_t, n1, t1, t1 + 16 * n1);
  return launch_status("fft4_tables_kernel");
}

bool fft4_supported(int64_t n_t) {
  static const bool off = getenv("PSA_FFT4") != nullptr && atoi(getenv("PSA_FFT4")) == 0;
  return !off && fft4::n1_of(n_t) != 0;
}

static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  else cudaGetLastError();
  return sms > 0 ? sms : 148;
}

int64_t fft4_workspace_bytes(int64_t n_t, int64_t n_k) {
  if (!fft4_supported(n_t) || n_k <= 0) return 0;
  // sized for 148 SMs when no device is current (the planner is callable without a GPU)
  const fft4::Schedule s = fft4::make_schedule(n_t, n_k * 3, device_sms());
  return s.ctl_bytes + s.y_bytes;
}

template <int N1, int T, int OCC>
static int launch_occ(const fft4::Args& a, const fft4::Schedule& sch, cudaStream_t s) {
  using namespace fft4;
  const size_t smem = smem_bytes<N1, T>();
  PSA_CUDA(cudaFuncSetAttribute(fft4_kernel<N1, T, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (int64_t)sch.n_groups * 2 * Geo<N1, T>::tiles_per_group;
  const int grid = (int)(tiles < sch.resident ? tiles : sch.resident);
  fft4_kernel<N1, T, OCC><<<grid, T, smem, s>>>(a);
  return launch_status("fft4_kernel");
}

template <int N1, int T>
static int launch_one(const fft4::Args& a, const fft4::Schedule& sch, cudaStream_t s) {
  const int occ = fft4::occ_of(T);
  if constexpr (T == 128) {
    if (occ == 4) return launch_occ<N1, T, 4>(a, sch, s);
    if (occ == 5) return launch_occ<N1, T, 5>(a, sch, s);
    return launch_occ<N1, T, 6>(a, sch, s);
  } else {
    if (occ == 2) return launch_occ<N1, T, 2>(a, sch, s);
    return launch_occ<N1, T, 3>(a, sch, s);
  }
}

// coherent assembly only: complex64 out[f][k_offset + k][pol]
int launch_fft4(const float* P, int64_t n_k, int64_t n_t, int64_t ldp, const void* plan_buf, void* workspace,
                int64_t workspace_bytes, const float* window, void* out, int64_t n_k_total, int64_t k_offset,
                cudaStream_t s) {
  using namespace fft4;
  static bool w128_ready[64] = {};
  int dev = 0;
  PSA_CUDA(cudaGetDevice(&dev));
  const Schedule sch = make_schedule(n_t, n_k * 3, device_sms());
  PSA_REQUIRE(workspace != nullptr && workspace_bytes >= sch.ctl_bytes + sch.y_bytes,
              "psa_fft_sed: workspace of %lld bytes required for n_t=%lld (got %lld)",
              (long long)(sch.ctl_bytes + sch.y_bytes), (long long)n_t, (long long)workspace_bytes);
  PSA_REQUIRE(((uintptr_t)workspace & 255) == 0, "psa_fft_sed: workspace must be 256-byte aligned");
  if (dev < 64 && !w128_ready[dev]) {                       // per device; idempotent, so a race only repeats the copy
    double2 host[128];
    for (int e = 0; e < 128; ++e) {
      const double ang = -2.0 * 3.14159265358979323846 * (double)e / 128.0;
      host[e] = make_double2(cos(ang), sin(ang));
    }
    // exact values at the octants
    host[0] = make_double2(1.0, 0.0); host[32] = make_double2(0.0, -1.0); host[64] = make_double2(-1.0, 0.0);
    host[96] = make_double2(0.0, 1.0);
    PSA_CUDA(cudaMemcpyToSymbolAsync(c_w128, host, sizeof(host), 0, cudaMemcpyHostToDevice, s));
    PSA_CUDA(cudaStreamSynchronize(s));                    // `host` is on the stack
    w128_ready[dev] = true;
  }
  PSA_CUDA(cudaMemsetAsync(workspace, 0, (size_t)sch.ctl_bytes, s));
  Args a;
  a.P = P;
  a.ldp = ldp;
  a.window = window;
  a.tw = reinterpret_cast<const double2*>(plan_buf);
  a.t1 = a.tw + fft4_table_offset(n_t);
  a.t2 = a.t1 + 16 * n1_of(n_t);
  static const int prefetch_env = getenv("PSA_FFT4_PREFETCH") ? atoi(getenv("PSA_FFT4_PREFETCH")) : 0;
  a.prefetch = prefetch_env;
  int* ctl = reinterpret_cast<int*>(workspace);
  a.counter = ctl;
  a.done_a = ctl + 1;
  a.done_b = ctl + 1 + sch.n_groups;
  a.ybuf = reinterpret_cast<double2*>(reinterpret_cast<char*>(workspace) + sch.ctl_bytes);
  a.out = reinterpret_cast<float2*>(out) + k_offset * 3;
  a.fstride = n_k_total * 3;
  a.n_cols = (int)(n_k * 3);
  a.n_groups = sch.n_groups;
  a.lag = sch.lag;
  a.ring = sch.ring;
  const int n1 = n1_of(n_t), threads = threads_of(n1);
  a.total_items = (sch.n_groups + sch.lag) * 2 * sch.tpg;
  a.inv_n = 1.0 / (double)n_t;
  if (n1 == 64) return threads == 128 ? launch_one<64, 128>(a, sch, s) : launch_one<64, 256>(a, sch, s);
  if (n1 == 128) return threads == 128 ? launch_one<128, 128>(a, sch, s) : launch_one<128, 256>(a, sch, s);
  return launch_one<256, 256>(a, sch, s);
}

}  // namespace psa
