// extern "C" surface of libpsa_b200.so (declared in include/psa_b200.h): argument validation,
// error plumbing, and dispatch to the kernel launchers.  No state is kept between calls.
#include <stdarg.h>
#include <string.h>

#include <cuda.h>

#include "common.cuh"

namespace psa {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)err, cudaGetErrorString(err), what);
  return PSA_ERR_CUDA;
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace psa

using namespace psa;

extern "C" {

int psa_version(void) { return 100; }   // 0.1.0

const char* psa_last_error(void) { return g_error; }

int psa_device_check(int device) {
  cudaDeviceProp prop;
  PSA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return PSA_ERR_UNSUPPORTED;
  }
  return PSA_OK;
}

int64_t psa_pitch(int64_t n_sel) { return round_up(n_sel > 0 ? n_sel : 1, 64); }

int psa_mean_positions(const float* pos, int64_t n_t, int64_t n_a, float* mean, void* stream) {
  PSA_REQUIRE(n_t >= 0 && n_a >= 0, "psa_mean_positions: negative extent");
  PSA_REQUIRE(n_t == 0 || n_a == 0 || (pos && mean), "psa_mean_positions: null pointer");
  PSA_REQUIRE(n_t > 0 || n_a == 0, "psa_mean_positions: zero frames");
  return launch_mean_positions(pos, n_t, n_a, mean, as_stream(stream));
}

int psa_digitize(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_t,
                 int64_t n_a, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, void* stream) {
  PSA_REQUIRE(data && dig && expo, "psa_digitize: null pointer");
  PSA_REQUIRE(n_t > 0 && n_a > 0 && n_sel > 0, "psa_digitize: empty input (n_t=%lld n_a=%lld n_sel=%lld)",
              (long long)n_t, (long long)n_a, (long long)n_sel);
  PSA_REQUIRE(idx != nullptr || n_sel == n_a, "psa_digitize: n_sel must equal n_a when idx is NULL");
  PSA_REQUIRE(pitch >= n_sel && pitch % 64 == 0, "psa_digitize: pitch must be a multiple of 64 and >= n_sel");
  return launch_digitize(data, mean, weight, idx, n_t, n_a, n_sel, pitch, dig, expo, as_stream(stream));
}

int psa_mean_accumulate(const float* pos, int64_t n_rows, int64_t n_a, const float* acc_in, int64_t divide_by,
                        float* out, void* stream) {
  PSA_REQUIRE(n_rows >= 0 && n_a >= 0 && divide_by >= 0, "psa_mean_accumulate: negative extent");
  PSA_REQUIRE(n_a == 0 || (out && (pos || n_rows == 0)), "psa_mean_accumulate: null pointer");
  return launch_mean_accumulate(pos, n_rows, n_a, acc_in, divide_by, out, as_stream(stream));
}

int psa_digitize_rows(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                      int64_t n_a, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, int64_t n_t_total,
                      int64_t t0, void* stream) {
  PSA_REQUIRE(dig && expo && (data || n_rows == 0), "psa_digitize_rows: null pointer");
  PSA_REQUIRE(n_rows >= 0 && t0 >= 0 && t0 + n_rows <= n_t_total && n_a > 0 && n_sel > 0,
              "psa_digitize_rows: bad extent (rows [%lld, %lld) of %lld frames)", (long long)t0,
              (long long)(t0 + n_rows), (long long)n_t_total);
  PSA_REQUIRE(idx != nullptr || n_sel == n_a, "psa_digitize_rows: n_sel must equal n_a when idx is NULL");
  PSA_REQUIRE(pitch >= n_sel && pitch % 64 == 0, "psa_digitize_rows: pitch must be a multiple of 64 and >= n_sel");
  DigDests dst{};
  dst.n = 1;
  dst.dig[0] = dig;
  dst.expo[0] = expo;
  return launch_digitize_rows(data, mean, weight, idx, n_rows, n_a, n_sel, pitch, dst, n_t_total, t0, as_stream(stream));
}

int psa_digitize_rows_peers(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                            int64_t n_a, int64_t n_sel, int64_t pitch, void* const* dig_all_host,
                            void* const* expo_all_host, int64_t n_dst, int64_t n_t_total, int64_t t0, int light,
                            void* stream) {
  PSA_REQUIRE(dig_all_host && expo_all_host && (data || n_rows == 0), "psa_digitize_rows_peers: null pointer");
  PSA_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers, "psa_digitize_rows_peers: between 1 and %d destinations (got %lld)",
              kMaxPeers, (long long)n_dst);
  PSA_REQUIRE(n_rows >= 0 && t0 >= 0 && t0 + n_rows <= n_t_total && n_a > 0 && n_sel > 0,
              "psa_digitize_rows_peers: bad extent (rows [%lld, %lld) of %lld frames)", (long long)t0,
              (long long)(t0 + n_rows), (long long)n_t_total);
  PSA_REQUIRE(idx != nullptr || n_sel == n_a, "psa_digitize_rows_peers: n_sel must equal n_a when idx is NULL");
  PSA_REQUIRE(pitch >= n_sel && pitch % 64 == 0, "psa_digitize_rows_peers: pitch must be a multiple of 64 and >= n_sel");
  DigDests dst{};
  dst.n = (int)n_dst;
  for (int d = 0; d < dst.n; ++d) {
    PSA_REQUIRE(dig_all_host[d] && expo_all_host[d], "psa_digitize_rows_peers: destination %d is null", d);
    dst.dig[d] = reinterpret_cast<int8_t*>(dig_all_host[d]);
    dst.expo[d] = reinterpret_cast<int32_t*>(expo_all_host[d]);
  }
  return launch_digitize_rows(data, mean, weight, idx, n_rows, n_a, n_sel, pitch, dst, n_t_total, t0, as_stream(stream),
                              light != 0);
}

// ---- page-locking of caller-owned host memory (a result array shared by the ranks of one box)
int psa_host_register(void* host_ptr, int64_t bytes) {
  PSA_REQUIRE(host_ptr && bytes > 0, "psa_host_register: bad arguments");
  PSA_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable));
  return PSA_OK;
}

int psa_host_unregister(void* host_ptr) {
  if (host_ptr == nullptr) return PSA_OK;
  PSA_CUDA(cudaHostUnregister(host_ptr));
  return PSA_OK;
}

// ---- CUDA IPC: let a peer process (one process per GPU) map a buffer of this one
int psa_ipc_export(const void* ptr, void* handle64_host, int64_t* offset_host) {
  PSA_REQUIRE(ptr && handle64_host && offset_host, "psa_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard guard(ptr);
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = []() -> RangeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<RangeFn>(p);
  }();
  if (range_fn == nullptr) {
    set_error("psa_ipc_export: cuMemGetAddressRange is not available from the CUDA driver");
    return PSA_ERR_CUDA;
  }
  CUdeviceptr b = 0;
  size_t size = 0;
  CUresult r = range_fn(&b, &size, (CUdeviceptr)(uintptr_t)ptr);
  if (r != CUDA_SUCCESS) {
    set_error("psa_ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
    return PSA_ERR_CUDA;
  }
  void* base = reinterpret_cast<void*>((uintptr_t)b);
  cudaIpcMemHandle_t h;
  PSA_CUDA(cudaIpcGetMemHandle(&h, base));
  memcpy(handle64_host, &h, sizeof(h));
  *offset_host = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  return PSA_OK;
}

int psa_ipc_open(const void* handle64_host, void** base_out_host) {
  PSA_REQUIRE(handle64_host && base_out_host, "psa_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, sizeof(h));
  void* p = nullptr;
  PSA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out_host = p;
  return PSA_OK;
}

int psa_ipc_close(void* base) {
  if (base == nullptr) return PSA_OK;
  PSA_CUDA(cudaIpcCloseMemHandle(base));
  return PSA_OK;
}

int psa_phase_digits(const float* kvecs, int64_t n_k, const float* mean, const int32_t* idx, int64_t n_sel,
                     int64_t pitch, int64_t rows_alloc, int8_t* adig, void* stream) {
  PSA_REQUIRE(kvecs && mean && adig, "psa_phase_digits: null pointer");
  PSA_REQUIRE(n_k > 0 && n_k <= 65535, "psa_phase_digits: n_k per call must be in [1, 65535] (chunk the k list)");
  PSA_REQUIRE(n_sel > 0 && pitch >= n_sel && pitch % 64 == 0, "psa_phase_digits: bad n_sel/pitch");
  PSA_REQUIRE(rows_alloc >= 2 * n_k, "psa_phase_digits: rows_alloc < 2 n_k");
  DeviceGuard guard(adig);
  return launch_phase_digits(kvecs, n_k, mean, idx, n_sel, pitch, rows_alloc, adig, as_stream(stream));
}

int psa_project_rows(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig, const int32_t* expo,
                     int64_t n_t, int64_t t0, int64_t n_t_rows, int64_t n_sel, int64_t pitch, float* P, int64_t ldp,
                     int impl, void* stream) {
  PSA_REQUIRE(adig && bdig && expo && P, "psa_project: null pointer");
  PSA_REQUIRE(rows > 0 && rows <= rows_alloc && n_t > 0 && n_sel > 0, "psa_project: bad extents");
  PSA_REQUIRE(t0 >= 0 && n_t_rows >= 0 && t0 + n_t_rows <= n_t, "psa_project: frame range [%lld, %lld) outside %lld frames",
              (long long)t0, (long long)(t0 + n_t_rows), (long long)n_t);
  PSA_REQUIRE(pitch >= n_sel && pitch % 64 == 0, "psa_project: pitch must be a multiple of 64 and >= n_sel");
  PSA_REQUIRE(ldp >= n_t && ldp % 4 == 0, "psa_project: ldp must be a multiple of 4 and >= n_t");
  PSA_REQUIRE(((uintptr_t)adig % 16) == 0 && ((uintptr_t)bdig % 16) == 0 && ((uintptr_t)P % 16) == 0,
              "psa_project: buffers must be 16-byte aligned");
  if (n_t_rows == 0) return PSA_OK;
  DeviceGuard guard(adig);       // P may be a peer GPU's buffer (frame-sharded multi-GPU): the phase digits are always local
  const int8_t* b0 = bdig + t0 * pitch;
  const int32_t* e0 = expo + t0;
  float* p0 = P + t0;
  if (impl == PSA_PROJECT_TENSOR)
    return launch_project_tc2(adig, rows, rows_alloc, b0, e0, n_t_rows, n_t, n_sel, pitch, p0, ldp, as_stream(stream));
  if (impl == PSA_PROJECT_SIMT)
    return launch_project_simt(adig, rows, rows_alloc, b0, e0, n_t_rows, n_t, n_sel, pitch, p0, ldp, as_stream(stream));
  set_error("psa_project: unknown impl %d", impl);
  return PSA_ERR_BAD_ARG;
}

int psa_project_routed(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig, const int32_t* expo,
                       int64_t n_t, int64_t n_sel, int64_t pitch, float* const* dests, const int64_t* row_begin,
                       int n_dest, int64_t ldp, void* stream) {
  PSA_REQUIRE(adig && bdig && expo && dests && row_begin, "psa_project_routed: null pointer");
  PSA_REQUIRE(rows > 0 && rows <= rows_alloc && n_t > 0 && n_sel > 0, "psa_project_routed: bad extents");
  PSA_REQUIRE(n_dest >= 1 && n_dest <= kMaxRouteDests, "psa_project_routed: 1 to %d destinations", kMaxRouteDests);
  PSA_REQUIRE(pitch >= n_sel && pitch % 64 == 0, "psa_project_routed: pitch must be a multiple of 64 and >= n_sel");
  PSA_REQUIRE(ldp >= n_t && ldp % 4 == 0, "psa_project_routed: ldp must be a multiple of 4 and >= n_t");
  PSA_REQUIRE(((uintptr_t)adig % 16) == 0 && ((uintptr_t)bdig % 16) == 0, "psa_project_routed: buffers must be 16-byte aligned");
  for (int q = 0; q < n_dest; ++q)
    PSA_REQUIRE(((uintptr_t)dests[q] % 16) == 0, "psa_project_routed: destination %d must be 16-byte aligned", q);
  DeviceGuard guard(adig);       // the destinations may be peer GPUs' buffers; the digits are always local
  return launch_project_tc2(adig, rows, rows_alloc, bdig, expo, n_t, n_t, n_sel, pitch, nullptr, ldp, as_stream(stream),
                            dests, row_begin, n_dest);
}

int psa_project(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig, const int32_t* expo,
                int64_t n_t, int64_t n_sel, int64_t pitch, float* P, int64_t ldp, int impl, void* stream) {
  return psa_project_rows(adig, rows, rows_alloc, bdig, expo, n_t, 0, n_t, n_sel, pitch, P, ldp, impl, stream);
}

int64_t psa_fft_plan_bytes(int64_t n_t) {
  int64_t bytes = 0;
  return fft_plan_bytes(n_t, &bytes) == PSA_OK ? bytes : -1;
}

int64_t psa_fft_workspace_bytes(int64_t n_t, int64_t n_k, int64_t n_groups) {
  int64_t bytes = 0;
  return fft_workspace_bytes(n_t, n_k, n_groups, &bytes) == PSA_OK ? bytes : -1;
}

int psa_fft_plan_init(int64_t n_t, void* plan, void* stream) {
  PSA_REQUIRE(plan != nullptr, "psa_fft_plan_init: null plan buffer");
  PSA_REQUIRE(((uintptr_t)plan % 16) == 0, "psa_fft_plan_init: plan buffer must be 16-byte aligned");
  DeviceGuard guard(plan);
  return launch_fft_plan(n_t, plan, as_stream(stream));
}

int psa_fft_sed(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t, int64_t ldp,
                const void* plan, void* workspace, int64_t workspace_bytes, const float* window, int mode, void* out,
                int64_t n_k_total, int64_t k_offset, void* stream) {
  PSA_REQUIRE(P && plan && out, "psa_fft_sed: null pointer");
  PSA_REQUIRE(n_groups >= 1 && n_k > 0 && n_t > 0 && ldp >= n_t, "psa_fft_sed: bad extents");
  PSA_REQUIRE(k_offset >= 0 && k_offset + n_k <= n_k_total, "psa_fft_sed: k range outside the result");
  PSA_REQUIRE(mode == PSA_MODE_INCOHERENT || n_groups == 1, "psa_fft_sed: coherent mode takes one group");
  DeviceGuard guard(out);
  return launch_fft(P, n_groups, group_stride, n_k, n_t, ldp, plan, workspace, workspace_bytes, window, mode, out,
                    n_k_total, k_offset, as_stream(stream));
}

int psa_chiral_phase(const float* z1, const float* z2, int64_t n, int64_t stride1, int64_t stride2, int opt,
                     float* out, void* stream) {
  PSA_REQUIRE(n >= 0 && (n == 0 || (z1 && z2 && out)), "psa_chiral_phase: bad arguments");
  DeviceGuard guard(out);
  return launch_chiral(reinterpret_cast<const float2*>(z1), reinterpret_cast<const float2*>(z2), n, stride1, stride2,
                       opt, out, as_stream(stream));
}

int psa_intensity(const float* sed, int64_t n_rows, int n_pol, float* out, void* stream) {
  PSA_REQUIRE(n_rows >= 0 && n_pol > 0 && (n_rows == 0 || (sed && out)), "psa_intensity: bad arguments");
  DeviceGuard guard(out);
  return launch_intensity(reinterpret_cast<const float2*>(sed), n_rows, n_pol, out, as_stream(stream));
}

static int ised_batch(IsedBatch* b, const char* who, const float* mean, const float* khat, const float* k_act,
                      const float* amp, const int32_t* member_off, const int32_t* member_grp, int64_t n_groups,
                      int64_t n_a, int64_t n_frames, int64_t n_points) {
  PSA_REQUIRE(n_a >= 0 && n_frames >= 0 && n_points >= 0 && n_groups >= 1, "%s: bad extent", who);
  PSA_REQUIRE(n_a * n_frames * n_points == 0 || (mean && khat && k_act && amp && member_off && member_grp),
              "%s: null pointer", who);
  PSA_REQUIRE(n_frames < (1 << 30) && n_points < (1 << 30) && n_groups < (1 << 30), "%s: extent too large", who);
  *b = IsedBatch{mean, khat, k_act, reinterpret_cast<const float2*>(amp), member_off, member_grp, (int)n_groups, n_a,
                 (int)n_frames, (int)n_points};
  return PSA_OK;
}

int psa_ised_absmax(const float* mean, const float* khat, const float* k_act, const float* amp,
                    const int32_t* member_off, const int32_t* member_grp, int64_t n_groups, int64_t n_a,
                    int64_t n_frames, int64_t n_points, float* wmax, void* stream) {
  IsedBatch b;
  int st = ised_batch(&b, "psa_ised_absmax", mean, khat, k_act, amp, member_off, member_grp, n_groups, n_a, n_frames, n_points);
  if (st != PSA_OK) return st;
  PSA_REQUIRE(wmax || n_points == 0, "psa_ised_absmax: null output");
  DeviceGuard guard(wmax);
  return launch_ised_absmax(b, wmax, as_stream(stream));
}

int psa_ised_frames(const float* mean, const float* khat, const float* k_act, const float* amp,
                    const int32_t* member_off, const int32_t* member_grp, int64_t n_groups, int64_t n_a,
                    int64_t n_frames, int64_t n_points, const float* div, const float* mul, float* out, void* stream) {
  IsedBatch b;
  int st = ised_batch(&b, "psa_ised_frames", mean, khat, k_act, amp, member_off, member_grp, n_groups, n_a, n_frames, n_points);
  if (st != PSA_OK) return st;
  PSA_REQUIRE(n_a * n_frames * n_points == 0 || (div && mul && out), "psa_ised_frames: null pointer");
  DeviceGuard guard(out);
  return launch_ised_frames(b, div, mul, out, as_stream(stream));
}

int psa_gather_bins(const float* sed, int64_t n_k, const int32_t* w_idx, const int32_t* k_idx, int64_t n_points,
                    int64_t out_stride, float* out, void* stream) {
  PSA_REQUIRE(n_points >= 0 && n_points < (1 << 28) && n_k > 0 && out_stride >= 3, "psa_gather_bins: bad extent");
  PSA_REQUIRE(n_points == 0 || (sed && w_idx && k_idx && out), "psa_gather_bins: null pointer");
  DeviceGuard guard(out);
  return launch_gather_bins(reinterpret_cast<const float2*>(sed), n_k, w_idx, k_idx, (int)n_points, out_stride,
                            reinterpret_cast<float2*>(out), as_stream(stream));
}

int psa_disp_moments(const float* pos, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                     int64_t n_sel, double* out2, void* stream) {
  PSA_REQUIRE(pos && mean && out2, "psa_disp_moments: null pointer");
  PSA_REQUIRE(idx != nullptr || n_sel == n_a, "psa_disp_moments: n_sel must equal n_a when idx is NULL");
  DeviceGuard guard(out2);
  return launch_disp_moments(pos, mean, idx, n_t, n_a, n_sel, out2, as_stream(stream));
}

int psa_scale_intensity(float* x, int64_t n, int mode, void* stream) {
  PSA_REQUIRE(n >= 0 && (n == 0 || x), "psa_scale_intensity: bad arguments");
  DeviceGuard guard(x);
  return launch_scale_intensity(x, n, mode, as_stream(stream));
}

int psa_minmax(const float* x, int64_t n, void* out3, void* stream) {
  PSA_REQUIRE(n >= 0 && out3 && (n == 0 || x), "psa_minmax: bad arguments");
  DeviceGuard guard(out3);
  return launch_minmax(x, n, out3, as_stream(stream));
}

int psa_select_pass(const float* x, int64_t n, int done_bits, const uint32_t* prefix, int m, uint32_t* hist, void* stream) {
  PSA_REQUIRE(n >= 0 && prefix && hist && (n == 0 || x), "psa_select_pass: bad arguments");
  DeviceGuard guard(hist);
  return launch_select_pass(x, n, done_bits, prefix, m, hist, as_stream(stream));
}

int psa_absmax(const float* x, int64_t n, float* out, void* stream) {
  PSA_REQUIRE(out && (n == 0 || x), "psa_absmax: null pointer");
  DeviceGuard guard(out);
  return launch_absmax(x, n, out, as_stream(stream));
}

int psa_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                  void* stream) {
  if (width == 0 || height == 0) return PSA_OK;
  PSA_REQUIRE(dst && src, "psa_copy_rows: null pointer");
  PSA_REQUIRE(width > 0 && height > 0 && dst_pitch >= width && src_pitch >= width,
              "psa_copy_rows: bad extent (width=%lld height=%lld pitches %lld / %lld)", (long long)width,
              (long long)height, (long long)dst_pitch, (long long)src_pitch);
  DeviceGuard guard_src(src), guard_dst(dst);     // whichever of the two is device memory decides
  PSA_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)height,
                             cudaMemcpyDefault, as_stream(stream)));
  return PSA_OK;
}

}  // extern "C"
