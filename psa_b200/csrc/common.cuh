// Shared declarations of the psa_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/psa_b200.h"

namespace psa {

// ---- error plumbing: every entry point returns a status, the text is thread-local
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define PSA_CUDA(call)                                            \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::psa::cuda_fail(_e, #call);    \
  } while (0)

#define PSA_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) {                                                \
      ::psa::set_error(__VA_ARGS__);                              \
      return PSA_ERR_BAD_ARG;                                     \
    }                                                             \
  } while (0)

// Every entry point runs on the device that owns its buffers, whatever the calling thread's current device is
// (the reference's GUI calls calculate() from a worker thread, psa_gui.py:1015; the current device is per-thread
// state).  The guard looks the device up from one of the call's pointers and restores the previous one on exit.
class DeviceGuard {
 public:
  explicit DeviceGuard(const void* ptr) {
    cudaPointerAttributes attr;
    if (ptr == nullptr || cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return; }
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev_) != cudaSuccess) { prev_ = -1; return; }
    if (attr.device != prev_) { if (cudaSetDevice(attr.device) == cudaSuccess) switched_ = true; }
  }
  ~DeviceGuard() { if (switched_) cudaSetDevice(prev_); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
 private:
  int prev_ = -1;
  bool switched_ = false;
};

inline int launch_status(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, kernel);
  return PSA_OK;
}

// ---- digit format shared by the digitiser, the phase generator and both projection kernels
//
// A real number x with |x| <= 1 (phase table) or |x| < 2^e (trajectory row) is held as the
// integer X = rint(x * 2^(30-e)), |X| <= 2^30, written in balanced base 256:
//     X = d0 + 256 d1 + 256^2 d2 + 256^3 d3,   d0..d2 in [-128,127], d3 in [-64,64].
// Each digit plane is an int8 matrix the tensor cores multiply exactly (int32 accumulate).
constexpr int kSlices = 4;
constexpr int kFracBits = 30;
// digit-pair classes kept by the projection: i + j >= 3  (10 products, weights 256^(i+j))
constexpr int kMinClass = 3;
constexpr int kClasses = 4;

__host__ __device__ inline void balanced_digits(int32_t x, int8_t d[4]) {
  int32_t r = x;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int8_t lo = (int8_t)(r & 0xFF);
    d[i] = lo;
    r = (r - (int32_t)lo) >> 8;   // exact: r - lo is a multiple of 256
  }
  d[3] = (int8_t)r;
}

constexpr int kExpMin = -80;   // exponent stored for an all-zero row
constexpr int kExpMax = 100;
constexpr int kExpPoison = 0x40000000;   // exponent stored for a row holding NaN / Inf / |x| >= 2^kExpMax: projections are NaN

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// device-side kernels' host launchers (one per .cu file)
int launch_mean_positions(const float* pos, int64_t n_t, int64_t n_a, float* mean, cudaStream_t s);
int launch_mean_accumulate(const float* pos, int64_t n_t, int64_t n_a, const float* acc_in, int64_t divide_by, float* mean,
                           cudaStream_t s);
// Destinations of one digitised row: this GPU's digit planes and, in a multi-GPU sliced ingest, the peers' planes
// mapped through CUDA IPC (psa_ipc_open) - the all-gather of the planes is fused into the kernel that produces them.
constexpr int kMaxPeers = 8;
struct DigDests {
  int n;
  int8_t* dig[kMaxPeers];
  int32_t* expo[kMaxPeers];
};
int launch_digitize_rows(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                         int64_t n_a, int64_t n_sel, int64_t pitch, const DigDests& dst, int64_t n_t_total, int64_t t0,
                         cudaStream_t s, bool light = false);
int launch_digitize(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_t,
                    int64_t n_a, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, cudaStream_t s);
int launch_phase_digits(const float* kvecs, int64_t n_k, const float* mean, const int32_t* idx,
                        int64_t n_sel, int64_t pitch, int64_t rows_alloc, int8_t* adig, cudaStream_t s);
// dests / row_begin / n_dest (host arrays): rows [row_begin[q], row_begin[q + 1]) are stored as rows 0.. of dests[q]
// instead of into P (frame-sharded multi-GPU run: one launch for every owner's k-points); n_dest == 0: everything into P
constexpr int kMaxRouteDests = 8;
int launch_project_tc2(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                       const int32_t* expo, int64_t n_t, int64_t n_t_total, int64_t n_sel, int64_t pitch, float* P,
                       int64_t ldp, cudaStream_t s, float* const* dests = nullptr, const int64_t* row_begin = nullptr,
                       int n_dest = 0);
int launch_project_simt(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                        const int32_t* expo, int64_t n_t, int64_t n_t_total, int64_t n_sel, int64_t pitch, float* P,
                        int64_t ldp, cudaStream_t s);
int fft_plan_bytes(int64_t n_t, int64_t* bytes);
int fft_workspace_bytes(int64_t n_t, int64_t n_k, int64_t n_groups, int64_t* bytes);
int launch_fft_plan(int64_t n_t, void* plan, cudaStream_t s);
int launch_fft(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
               const void* plan, void* workspace, int64_t workspace_bytes, const float* window, int mode, void* out,
               int64_t n_k_total, int64_t k_offset, cudaStream_t s);
// four-step kernel for long power-of-two columns (fft4.cu); coherent assembly only
bool fft4_supported(int64_t n_t);
int64_t fft4_table_entries(int64_t n_t);
int64_t fft4_table_offset(int64_t n_t);
int launch_fft4_tables(int64_t n_t, double2* t1, cudaStream_t s);
int64_t fft4_workspace_bytes(int64_t n_t, int64_t n_k);
int launch_fft4(const float* P, int64_t n_k, int64_t n_t, int64_t ldp, const void* plan, void* workspace,
                int64_t workspace_bytes, const float* window, void* out, int64_t n_k_total, int64_t k_offset,
                cudaStream_t s);
int launch_chiral(const float2* z1, const float2* z2, int64_t n, int64_t stride1, int64_t stride2,
                  int opt, float* out, cudaStream_t s);
int launch_intensity(const float2* sed, int64_t n_rows, int n_pol, float* out, cudaStream_t s);
struct IsedBatch {
  const float* mean;          // [n_a][3]
  const float* khat;          // [3]
  const float* k_act;         // [n_points]
  const float2* amp;          // [n_points][n_groups][3] complex64 amplitudes S_g[w, k, pol]
  const int32_t* member_off;  // [n_a + 1] CSR offsets into member_grp
  const int32_t* member_grp;  // group ids of every atom, ascending (= the reference's group loop order)
  int n_groups;
  int64_t n_a;
  int n_frames, n_points;
};
int launch_ised_absmax(const IsedBatch& b, float* wmax, cudaStream_t s);
int launch_ised_frames(const IsedBatch& b, const float* div, const float* mul, float* out, cudaStream_t s);
int launch_gather_bins(const float2* sed, int64_t n_k, const int32_t* w_idx, const int32_t* k_idx, int n_points,
                       int64_t out_stride, float2* out, cudaStream_t s);
int launch_scale_intensity(float* x, int64_t n, int mode, cudaStream_t s);
int launch_minmax(const float* x, int64_t n, void* out3, cudaStream_t s);
int launch_select_pass(const float* x, int64_t n, int done_bits, const uint32_t* prefix, int m, unsigned int* hist,
                       cudaStream_t s);
int launch_disp_moments(const float* pos, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                        int64_t n_sel, double* out2, cudaStream_t s);
int launch_absmax(const float* x, int64_t n, float* out, cudaStream_t s);

}  // namespace psa
