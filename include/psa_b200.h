/* psa_b200 - C ABI of the B200-native SED hot path.
 *
 * Every function is `extern "C"`, takes plain pointers and sizes, returns an int status
 * (PSA_OK == 0) and never throws or aborts.  On failure the message is available from
 * psa_last_error() (thread-local).  All data pointers are DEVICE pointers unless the
 * parameter name ends in `_host`; the library never frees or retains caller memory.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 * asynchronous with respect to the host unless stated otherwise, re-entrant, and may be
 * issued from any host thread (the reference's GUI calls `calculate` from a worker thread:
 * src/psa/gui/psa_gui.py:1015); bind via ctypes.CDLL so the GIL is released.
 *
 * The reference (h-walk/PSA) is pure Python, so there is no existing FFI to mirror; each
 * entry point replaces one NumPy call site of src/psa/core/sed_calculator.py, cited below.
 * INTEGRATION.md shows the ctypes binding a maintainer would add to the reference.
 */
#ifndef PSA_B200_H
#define PSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSA_OK 0
#define PSA_ERR_BAD_ARG (-1)   /* maps to ValueError in the Python shim   */
#define PSA_ERR_CUDA (-2)      /* maps to RuntimeError                     */
#define PSA_ERR_UNSUPPORTED (-3)

/* psa_fft_sed() epilogues */
#define PSA_MODE_COHERENT 0    /* complex64 out[n_f][n_k_total][3]                        */
#define PSA_MODE_INCOHERENT 1  /* float32  out[n_f][n_k_total] = sum_groups sum_pol |S|^2 */

/* psa_project() kernels */
#define PSA_PROJECT_TENSOR 0   /* tcgen05 int8 tensor-core kernel (the product path)      */
#define PSA_PROJECT_SIMT 1     /* dp4a CUDA-core kernel, same exact result (bring-up/validation) */
#define PSA_PROJECT_TENSOR_PAIR 2 /* tcgen05 cta_group::2 variant of the tensor-core kernel (same result) */

int psa_version(void);
const char* psa_last_error(void);

/* 0 when `device` is a Blackwell sm_100 part this library was built for. */
int psa_device_check(int device);

/* Row pitch (bytes == atoms) the digit planes must use for n_sel selected atoms. */
int64_t psa_pitch(int64_t n_sel);

/* Time-averaged positions, float32, bit-identical to np.mean(positions, axis=0, dtype=float32)
 * (reference: sed_calculator.py:205, 384): sequential float32 sum over frames, then / n_t.
 *   pos  [n_t][n_a][3] float32      mean [n_a][3] float32 */
int psa_mean_positions(const float* pos, int64_t n_t, int64_t n_a, float* mean, void* stream);

/* Select + split the projected time series into int8 digit planes (one-time per trajectory/group).
 * Replaces the fancy-index copy velocities[:, idx, :] / positions[:, idx, :] - mean
 * (reference: sed_calculator.py:69-72).
 *   data [n_t][n_a][3] float32 (velocities, or positions when `mean` != NULL -> displacements)
 *   mean [n_a][3] float32 or NULL      idx [n_sel] int32 atom indices or NULL (= all atoms, n_sel == n_a)
 *   dig  [3 pol][4 slice][n_t][pitch] int8     expo [3][n_t] int32 (row exponent e: |x| < 2^e) */
int psa_digitize(const float* data, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                 int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, void* stream);

/* The two ingest steps on a contiguous range of frames, for a trajectory whose frames are spread over several
 * GPUs (each rank uploads 1/N of the frames over its own PCIe link):
 *   psa_mean_accumulate: out[c] = (acc_in ? acc_in[c] : 0) (+) pos[0][c] (+) ... (+) pos[n_rows-1][c] in float32, in
 *     row order, then / (float)divide_by when divide_by > 0.  Chained over the ranks in frame order this is the
 *     same sequence of additions as psa_mean_positions, i.e. np.mean(..., axis=0, dtype=float32) bit for bit
 *     (reference: sed_calculator.py:205).  acc_in may alias out.
 *   psa_digitize_rows: psa_digitize for rows [t0, t0 + n_rows) of an n_t_total-frame trajectory; `data` points at
 *     row t0, dig / expo are the full-size outputs. */
int psa_mean_accumulate(const float* pos, int64_t n_rows, int64_t n_atoms, const float* acc_in, int64_t divide_by,
                        float* out, void* stream);
int psa_digitize_rows(const float* data, const float* mean, const int32_t* idx, int64_t n_rows, int64_t n_atoms,
                      int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, int64_t n_t_total, int64_t t0,
                      void* stream);

/* Phase table exp(+i k.r) as digit planes.  theta = fma(k2,r2,fma(k1,r1,k0*r0)) in float32, then
 * correctly rounded float32 cos/sin (reference: sed_calculator.py:78, np.exp(1j*np.dot(k, r.T))).
 *   kvecs [n_k][3] float32   mean [n_a][3] float32   idx [n_sel] or NULL
 *   adig  [4 slice][rows_alloc][pitch] int8, row 2k = cos, row 2k+1 = sin, rows_alloc >= 2 n_k */
int psa_phase_digits(const float* kvecs, int64_t n_k, const float* mean, const int32_t* idx,
                     int64_t n_sel, int64_t pitch, int64_t rows_alloc, int8_t* adig, void* stream);

/* Projection  P[row][pol][t] = sum_atoms phase[row][atom] * data[t][atom][pol]  (exact integer
 * contraction of the digit planes, one float32 rounding at the end).  Replaces the einsum/cgemm loop
 * (reference: sed_calculator.py:80-81).
 *   P [rows][3][ldp] float32, ldp >= n_t and ldp % 4 == 0;  impl = PSA_PROJECT_* */
int psa_project(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                const int32_t* expo, int64_t n_t, int64_t n_sel, int64_t pitch, float* P, int64_t ldp,
                int impl, void* stream);

/* FFT plan for n_t frames: twiddles, plus (when n_t is not a power of two) the Bluestein chirp and
 * its spectrum.  The caller owns the buffer: allocate psa_fft_plan_bytes(n_t) device bytes (-1 = n_t not
 * supported: 2 <= n_t <= 2^19), fill it once with psa_fft_plan_init, reuse it for every psa_fft_sed. */
int64_t psa_fft_plan_bytes(int64_t n_t);
int psa_fft_plan_init(int64_t n_t, void* plan, void* stream);

/* Scratch bytes psa_fft_sed needs for one call (0 for power-of-two n_t; the chirped spectra otherwise). */
int64_t psa_fft_workspace_bytes(int64_t n_t, int64_t n_k, int64_t n_groups);

/* Time FFT of every (k, pol) column of P, scaled 1/n_t, fused with the assembly epilogue.
 * Replaces np.fft.fft(axis=0)/n_t and the coherent / incoherent assembly
 * (reference: sed_calculator.py:83-84, 296-327).  Any n_t in [2, 2^19]: powers of two >= 32 run as one
 * in-shared-memory transform per column, other lengths through Bluestein's chirp-z identity.
 *   P   [n_groups][2 n_k][3][ldp] float32 (group g at P + g*group_stride floats)
 *   plan, workspace: see above
 *   out coherent: complex64 [n_t][n_k_total][3], this call fills k in [k_offset, k_offset+n_k)
 *       incoherent: float32 [n_t][n_k_total] */
int psa_fft_sed(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t,
                int64_t ldp, const void* plan, void* workspace, int64_t workspace_bytes, int mode, void* out,
                int64_t n_k_total, int64_t k_offset, void* stream);

/* Chiral phase of two complex components (reference: sed_calculator.py:338-371).
 *   z1, z2 complex64 with element strides stride1/stride2 (in complex elements), n elements
 *   opt 'C' (folded angle difference), 'A' (acos), 'B' (asin);  out float32 [n] */
int psa_chiral_phase(const float* z1, const float* z2, int64_t n, int64_t stride1, int64_t stride2,
                     int opt, float* out, void* stream);

/* intensity[r] = sum_pol |sed[r][pol]|^2 (reference: sed.py:22-24). sed complex64 [n_rows][n_pol]. */
int psa_intensity(const float* sed, int64_t n_rows, int n_pol, float* out, void* stream);

/* Inverse projection (reference: sed_calculator.py:494-499, 533):
 *   out[f][a][p] = (add_mean ? mean[a][p] : 0) + scale * Re( amp[a][p] * exp(i (2 pi f / n_frames - k_act * (mean[a] . khat))) )
 *   amp  [n_a][3][2] float64 (re, im), zero for atoms that are not reconstructed
 *   khat [3] float32 (device)          out [n_frames][n_a][3] float32
 * add_mean = 0 returns the bare oscillation (used to find max |wiggle| for the 'auto' rescale). */
int psa_ised_frames(const float* mean, const double* amp, const float* khat, float k_act, double scale,
                    int add_mean, int64_t n_a, int64_t n_frames, float* out, void* stream);

/* sum and sum of squares of (pos - mean) over the selected atoms and all frames, float64
 * (for the 'auto' rescale of iSED, reference: sed_calculator.py:506-507).  out2 [2] float64. */
int psa_disp_moments(const float* pos, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                     int64_t n_sel, double* out2, void* stream);

/* max |x| over n float32 values (device scalar out). */
int psa_absmax(const float* x, int64_t n, float* out, void* stream);

/* Strided row copy between any two of {device, pinned host}: `height` rows of `width` bytes, row pitches in
 * bytes.  Used to stream the spectra of one k-chunk into its column slice of the host result
 * full_sed_data[:, k0:k1, :] (reference: sed_calculator.py:310, 325) while the next chunk is computed. */
int psa_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSA_B200_H */
