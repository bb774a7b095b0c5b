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
#define PSA_PROJECT_TENSOR 0   /* tcgen05 cta_group::2 int8 tensor-core kernel (the product path) */
#define PSA_PROJECT_SIMT 1     /* dp4a CUDA-core kernel, same exact result (bring-up/validation)  */

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
 *   weight [n_a] float32 or NULL: per-atom factor applied in float32 after the mean subtraction - the README
 *          facade's mass weighting sqrt(m) (reference: README.md:83-101; the shipped source is unweighted = NULL)
 *   dig  [3 pol][4 slice][n_t][pitch] int8
 *   expo [3][n_t] int32 (row exponent e: |x| < 2^e).  A row holding NaN / Inf / |x| >= 2^100 gets the poison
 *        exponent 0x40000000 and zero digits; psa_project turns it into NaN projections (the reference's float
 *        arithmetic propagates non-finite samples into the spectrum as well). */
int psa_digitize(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_t,
                 int64_t n_a, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, void* stream);

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
int psa_digitize_rows(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                      int64_t n_atoms, int64_t n_sel, int64_t pitch, int8_t* dig, int32_t* expo, int64_t n_t_total,
                      int64_t t0, void* stream);

/* psa_digitize_rows writing every row into SEVERAL sets of digit planes at once: this GPU's and its peers' (device
 * pointers into other GPUs' memory obtained with psa_ipc_open).  In a multi-GPU sliced ingest each rank digitises
 * its own frames straight into every rank's planes through NVLink - the all-gather that would follow
 * (the one exchange step of the k-sharded path) is fused into the producing kernel.
 *   dig_all_host / expo_all_host: HOST arrays of n_dst (<= 8) device pointers, each laid out like dig / expo above.
 *   light != 0: small-footprint launch (128-thread CTAs, no shared-memory staging) that fits on the SMs next to a
 *   running psa_project - the ring steps of a pipelined exchange run under the first k-chunk's projection.
 * The caller orders the ranks around the call (nobody still reads the previous planes; everybody has finished
 * writing before anyone projects). */
int psa_digitize_rows_peers(const float* data, const float* mean, const float* weight, const int32_t* idx, int64_t n_rows,
                            int64_t n_atoms, int64_t n_sel, int64_t pitch, void* const* dig_all_host,
                            void* const* expo_all_host, int64_t n_dst, int64_t n_t_total, int64_t t0, int light,
                            void* stream);

/* CUDA IPC plumbing for the above (one process per GPU, same box):
 *   psa_ipc_export: 64-byte handle of the allocation that contains `ptr` + the offset of `ptr` inside it (host outputs)
 *   psa_ipc_open:   map a peer's allocation into this process (enables peer access); returns its base device pointer
 *   psa_ipc_close:  unmap it.  An allocation may be opened once per process at a time. */
int psa_ipc_export(const void* ptr, void* handle64_host, int64_t* offset_host);
int psa_ipc_open(const void* handle64_host, void** base_out_host);
int psa_ipc_close(void* base);

/* Phase table exp(+i k.r) as digit planes.  theta = fma(k2,r2,fma(k1,r1,k0*r0)) in float32, then
 * correctly rounded float32 cos/sin (reference: sed_calculator.py:78, np.exp(1j*np.dot(k, r.T))).
 *   kvecs [n_k][3] float32   mean [n_a][3] float32   idx [n_sel] or NULL
 *   adig  [4 slice][rows_alloc][pitch] int8, row 2k = cos, row 2k+1 = sin, rows_alloc >= 2 n_k */
int psa_phase_digits(const float* kvecs, int64_t n_k, const float* mean, const int32_t* idx,
                     int64_t n_sel, int64_t pitch, int64_t rows_alloc, int8_t* adig, void* stream);

/* Projection  P[row][pol][t] = sum_atoms phase[row][atom] * data[t][atom][pol]  (exact integer
 * contraction of the digit planes, one float32 rounding at the end).  Replaces the einsum/cgemm loop
 * (reference: sed_calculator.py:80-81).
 *   P [rows][3][ldp] float32, ldp >= n_t and ldp % 4 == 0;  impl = PSA_PROJECT_*
 * P may live on a PEER GPU (mapped with psa_ipc_open): a frame-sharded multi-GPU run projects its own frames for the
 * k-points another rank owns and the epilogue's 128-byte row stores go straight into that rank's buffer through
 * NVLink, at the column offset of this rank's frames (pass P + first_frame, ldp = the owner's row pitch) - the
 * frames-to-k transpose (an all-to-all) is fused into the projection.  The kernel runs on adig's device. */
int psa_project(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig,
                const int32_t* expo, int64_t n_t, int64_t n_sel, int64_t pitch, float* P, int64_t ldp,
                int impl, void* stream);

/* The same for the frames [t0, t0 + n_t_rows) only (all rows, all polarisations): bdig / expo / P are the full-size
 * buffers of an n_t-frame trajectory.  Every (row, frame) of P is an independent exact sum, so projecting a
 * trajectory range by range gives the bits of one call; a multi-GPU sliced ingest uses this to start projecting the
 * frames that have already arrived from a peer while the next peer's rows are still crossing NVLink. */
int psa_project_rows(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig, const int32_t* expo,
                     int64_t n_t, int64_t t0, int64_t n_t_rows, int64_t n_sel, int64_t pitch, float* P, int64_t ldp,
                     int impl, void* stream);

/* psa_project on the tensor cores with a destination per ROW RANGE: rows [row_begin[q], row_begin[q + 1]) of the
 * projection are stored as rows 0, 1, ... of dests[q] ([rows_q][3][ldp] float32; local or peer memory) instead of into
 * one P.  dests[n_dest] and row_begin[n_dest + 1] are HOST arrays, 1 <= n_dest <= 8, row_begin[0] = 0,
 * row_begin[n_dest] = rows.  A frame-sharded multi-GPU run projects a rank's frames for the k-points of ALL owners in one
 * launch (whole waves of tiles, full-width tiles across owner boundaries) - the all-to-all that turns "frames per rank"
 * into "k-points per rank" (the k-sharding of sed_calculator.py:287's chunk loop over ranks) is the epilogue's stores. */
int psa_project_routed(const int8_t* adig, int64_t rows, int64_t rows_alloc, const int8_t* bdig, const int32_t* expo,
                       int64_t n_t, int64_t n_sel, int64_t pitch, float* const* dests, const int64_t* row_begin,
                       int n_dest, int64_t ldp, void* stream);

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
 *   window [n_t] float32 or NULL: taper multiplied into every column before the transform (the README's
 *          "window"; the shipped source has none, sed_calculator.py:83 - NULL reproduces it exactly)
 *   out coherent: complex64 [n_t][n_k_total][3], this call fills k in [k_offset, k_offset+n_k)
 *       incoherent: float32 [n_t][n_k_total] */
int psa_fft_sed(const float* P, int64_t n_groups, int64_t group_stride, int64_t n_k, int64_t n_t,
                int64_t ldp, const void* plan, void* workspace, int64_t workspace_bytes, const float* window,
                int mode, void* out, int64_t n_k_total, int64_t k_offset, void* stream);

/* Chiral phase of two complex components (reference: sed_calculator.py:338-371).
 *   z1, z2 complex64 with element strides stride1/stride2 (in complex elements), n elements
 *   opt 'C' (folded angle difference), 'A' (acos), 'B' (asin);  out float32 [n] */
int psa_chiral_phase(const float* z1, const float* z2, int64_t n, int64_t stride1, int64_t stride2,
                     int opt, float* out, void* stream);

/* intensity[r] = sum_pol |sed[r][pol]|^2 (reference: sed.py:22-24). sed complex64 [n_rows][n_pol]. */
int psa_intensity(const float* sed, int64_t n_rows, int n_pol, float* out, void* stream);

/* Inverse projection, batched over (k, omega) points (reference: sed_calculator.py:440-441, 494-533).  For point p:
 *   w[f][a][pol]   = sum over the groups g of atom a, in order, of
 *                    Re( amp[p][g][pol] * exp(i (2 pi f / n_frames - k_act[p] * (mean[a] . khat))) )
 *                    (float64 terms, float32 running sum like the reference's `wiggles[...] += ...`)
 *   out[p][f][a][pol] = mean[a][pol] + (w / div[p]) * mul[p]          float32, the reference's order of operations:
 *                    'auto' rescale: div = max |w|, mul = n-weighted std of the displacements (:517-524);
 *                    numeric factor: div = 1, mul = factor (:527); div = mul = 1 leaves w untouched.
 *   mean [n_a][3] float32   khat [3] float32   k_act [n_points] float32   amp [n_points][n_groups][3] complex64
 *   member_off [n_a + 1], member_grp: CSR list of the groups each atom belongs to (ascending; empty = atom keeps
 *   its mean position)                     out [n_points][n_frames][n_a][3] float32, n_frames <= 65536
 * psa_ised_absmax: wmax[p] = max over frames, polarisations and the atoms of every group of |w after that group|
 *   (the reference's running max_wiggle_amp_all, :502-504); stores nothing else. */
int psa_ised_absmax(const float* mean, const float* khat, const float* k_act, const float* amp,
                    const int32_t* member_off, const int32_t* member_grp, int64_t n_groups, int64_t n_a,
                    int64_t n_frames, int64_t n_points, float* wmax, void* stream);
int psa_ised_frames(const float* mean, const float* khat, const float* k_act, const float* amp,
                    const int32_t* member_off, const int32_t* member_grp, int64_t n_groups, int64_t n_a,
                    int64_t n_frames, int64_t n_points, const float* div, const float* mul, float* out, void* stream);

/* out[p][0..2] = sed[w_idx[p]][k_idx[p]][0..2]: the complex amplitudes iSED reads from a group's SED
 * (reference: sed_calculator.py:494-496).  sed complex64 [n_f][n_k][3]; out complex64 rows of out_stride elements. */
int psa_gather_bins(const float* sed, int64_t n_k, const int32_t* w_idx, const int32_t* k_idx, int64_t n_points,
                    int64_t out_stride, float* out, void* stream);

/* sum and sum of squares of (pos - mean) over the selected atoms and all frames, float64
 * (for the 'auto' rescale of iSED, reference: sed_calculator.py:506-507).  out2 [2] float64. */
int psa_disp_moments(const float* pos, const float* mean, const int32_t* idx, int64_t n_t, int64_t n_a,
                     int64_t n_sel, double* out2, void* stream);

/* Device-side consumers of an intensity map, float32 [n] on the device (what the reference's plotter / GUI do on
 * the host with the whole array: sed_plotter.py:160-181, 211-215; psa_gui.py:2424-2441):
 *   psa_scale_intensity: in place; mode 0 linear, 1 log10(max(x, 1e-12)), 2 sqrt(max(x, 0)), 3 "dsqrt" sqrt(sqrt(max(x, 0)))
 *   psa_minmax: out3 (device, 16 bytes) = [key(nanmin) u32][key(nanmax) u32][number of finite values u64], where
 *               key(v) = bits(v) ^ (v < 0 ? 0xFFFFFFFF : 0x80000000) orders like the floats
 *   psa_select_pass: one byte of a most-significant-first radix select over the finite values - hist[j][b] counts
 *               the values whose key shares the top `done_bits` (0/8/16/24) bits of prefix[j] and whose next byte
 *               is b, for m <= 4 searches at once; four passes yield exact order statistics (np.percentile's inputs) */
int psa_scale_intensity(float* x, int64_t n, int mode, void* stream);
int psa_minmax(const float* x, int64_t n, void* out3, void* stream);
int psa_select_pass(const float* x, int64_t n, int done_bits, const uint32_t* prefix, int m, uint32_t* hist, void* stream);

/* max |x| over n float32 values (device scalar out). */
int psa_absmax(const float* x, int64_t n, float* out, void* stream);

/* Strided row copy between any two of {device, pinned host}: `height` rows of `width` bytes, row pitches in
 * bytes.  Used to stream the spectra of one k-chunk into its column slice of the host result
 * full_sed_data[:, k0:k1, :] (reference: sed_calculator.py:310, 325) while the next chunk is computed. */
int psa_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t height,
                  void* stream);

/* LAMMPS text dump of iSED frames, byte for byte what the reference's out_to_qdump writes (reference:
 * src/psa/io/writer.py:139-228; parsed back by src/psa/gui/psa_gui.py:1396-1455).  HOST pointers; frames are formatted
 * by `n_threads` threads (0 = all cores) and written in order.
 *   frames_host [n_frames][n_atoms][3] float32   types_host [n_atoms] int32   box9_host: the 3x3 box matrix, float32 */
int psa_write_dump(const char* path, const float* frames_host, const int32_t* types_host, int64_t n_frames,
                   int64_t n_atoms, const float* box9_host, int n_threads);

/* Page-lock / release caller-owned host memory so that psa_copy_rows can stream into it asynchronously.  Used for a
 * result array in POSIX shared memory that every rank of a box maps: each rank copies its k-slice of
 * full_sed_data[:, k0:k1] (reference: sed_calculator.py:310, 325) over its own PCIe link, no gather through rank 0. */
int psa_host_register(void* host_ptr, int64_t bytes);
int psa_host_unregister(void* host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* PSA_B200_H */
