/* trv_capi.h -- flat C entry points of libtrv_b200.so (host C++ API over the
 * device layer of trvb.h).
 *
 * libtrv_b200.so carries the reference's C++ surface (namespace trv::, headers
 * in triumvirate_b200/include/trv/ mirroring /root/reference/src/triumvirate/
 * include/ (the .hpp set)) AND the functions below, which play the role of the
 * reference's Cython layer for FFI callers that cannot link C++ symbols
 * (ctypes, cgo, JNI, ...): they marshal plain arrays into
 * trv::ParticleCatalogue / trv::ParameterSet / trv::Binning, call the
 * estimator and copy the result vectors back.
 *
 * Every function returns 0 on success; non-zero: 1 = C++ exception
 * (trv::sys::*Error), 2 = invalid argument/parameter, 3 = device error
 * (no CUDA device, CUDA/cuFFT failure).  trv_last_error() holds the message.
 * Unlike the reference's `compute_*` externs (no `except +`,
 * T/_threept.pyx:50-95) no exception escapes.  There is no CPU fallback: with
 * no usable GPU the estimators return 3.
 *
 * Citations: T/ = src/triumvirate/, S/ = src/triumvirate/src/ of the reference.
 */
#ifndef TRV_CAPI_H_INCLUDED_
#define TRV_CAPI_H_INCLUDED_

#ifdef __cplusplus
extern "C" {
#endif

const char* trv_last_error(void);

/* trv::sys::get_gpu_count (S/monitor.cpp:258-282), honours TRV_GPU_MODE and
 * TRV_GPU_MAXNUM. */
int trv_gpu_count(void);

/* Multi-GPU work split (no reference counterpart: trv::sys::currTask is the
 * constant 0, I/monitor.hpp:249-250).  Owner rank in [0, world) of each of the
 * *dim data-vector entries of a `form` ("diag" | "off-diag" | "row" | "full";
 * "full" with ell1 == ell2 is the upper triangle, S/parameters.cpp:829-849)
 * over num_bins bins; `owner` holds >= num_bins^2 ints.  Shares are compact in
 * the (row bin, column bin) matrix and equal in distinct shell fields (the
 * transforms a share costs), not in entries. */
int trv_partition_owners(const char* form, int ell1, int ell2, int idx_bin, int num_bins,
                         int world, int* owner, int* dim);

/* trv::sys::count_fft / count_ifft / gbytesMaxMemGPU (I/monitor.hpp:252-266). */
void trv_counters(int* count_fft, int* count_ifft, double* gib_gpu_max);

/* Three-point estimators: replaces T/_threept.pyx:128-174 (_compute_bispec /
 * _compute_3pcf -> trv::compute_bispec / compute_3pcf, S/threept.cpp:248,1014)
 * and T/_threept.pyx:226-248 (_compute_*_in_gpp_box -> S/threept.cpp:1473,2190).
 *   stat            "bispec" | "3pcf"
 *   catalogue_type  "sim" (periodic box, global plane-parallel) | "survey"
 *   per catalogue   n, x/y/z/nz/ws/wc float64 columns of length n (nz, ws, wc
 *                   may be NULL: 0, 1, 1 -- T/threept.py:1482-1485), LOS as a
 *                   contiguous (n, 3) array or NULL (I/dataobjs.hpp:186-188)
 *   parameters      the members of trv::ParameterSet that the path reads
 *                   (T/parameters.pxd:15-79); validate() is applied
 *                   (S/parameters.cpp:466-1270), so e.g. `interlace` is forced
 *                   off and form="full" with ell1==ell2 becomes "triu"
 *   custom_edges    num_bins+1 edges for binning="custom", else NULL
 *   deterministic   1: bit-reproducible assignment and reductions
 *   part_rank/part_count  multi-GPU work split: this call computes the entries
 *                   that trv_partition_owners() gives to part_rank (compact
 *                   blocks of the bin-pair matrix, so that a rank transforms
 *                   only the shells it pairs) and leaves zeros elsewhere (sum
 *                   over ranks = the full result).  Bispectrum with two or
 *                   more ranks: the last rank computes the shot noise of every
 *                   entry and no pairs; the pairs are dealt to the others
 * Outputs (capacity >= max(num_bins^2, num_bins) entries): *dim = dv_dim;
 * bin centres, effective coordinates, nmodes/npairs, raw and shot statistics as
 * interleaved (re, im) already multiplied by norm_factor
 * (I/dataobjs.hpp:249-282). */
int trv_threept(
  const char* stat, const char* catalogue_type,
  int nd, const double* xd, const double* yd, const double* zd,
  const double* nzd, const double* wsd, const double* wcd, const double* los_d,
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment, const char* interlace,
  int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  const double* custom_edges,
  double norm_factor, int verbose, int deterministic, int part_rank, int part_count,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot, double* elapsed_s);

/* Periodic-box estimators from three coordinate arrays in host memory
 * (on_device == 0) or CUDA device memory (on_device != 0); unit weights.
 * Same results as trv_threept("...", "sim", ...) without the AoS staging copy of
 * trv::ParticleCatalogue::load_particle_data (S/particles.cpp:512-563). */
int trv_threept_box_arrays(
  const char* stat, long long n, const double* x, const double* y, const double* z,
  int on_device,
  const double* boxsize, const int* ngrid, const char* assignment,
  int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  double norm_factor, int verbose, int deterministic, int part_rank, int part_count,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot);

/* Two-point estimators: replaces T/_twopt.pyx:164-380 (_compute_powspec,
 * _compute_corrfunc, _compute_powspec_in_gpp_box, _compute_corrfunc_in_gpp_box,
 * _compute_corrfunc_window -> S/twopt.cpp:388-901).  stat = "powspec" | "2pcf" |
 * "2pcf-win"; catalogue_type = "sim" | "survey" | "random" (for "2pcf-win": the
 * random catalogue goes in the `r` slot, used with `alpha`, T/twopt.py:1538).
 * `interlace` = "true" | "false" is honoured (two-point statistics are the only
 * ones for which validate() keeps it, S/parameters.cpp:1240-1249).  ELL is the
 * multipole degree.  Outputs hold >= num_bins entries: bin centres, effective
 * coordinates, nmodes/npairs, raw statistic and (power spectrum only) shot noise
 * as interleaved (re, im), multiplied by norm_factor. */
int trv_twopt(
  const char* stat, const char* catalogue_type,
  int nd, const double* xd, const double* yd, const double* zd,
  const double* nzd, const double* wsd, const double* wcd, const double* los_d,
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment, const char* interlace,
  int ELL, const char* binning, double bin_min, double bin_max, int num_bins,
  const double* custom_edges, double alpha, double norm_factor, int verbose, int deterministic,
  int* dim, double* c_bin, double* c_eff, int* count, double* raw, double* shot,
  double* elapsed_s);

/* 3PCF window function from a random catalogue: replaces T/_threept.pyx:276-314
 * (_compute_3pcf_window -> trv::compute_3pcf_window, S/threept.cpp:2621-3077).
 * The catalogue is used with the given alpha contrast (the Python front end
 * passes alpha = 1, T/threept.py:2052-2056); wide_angle != 0 multiplies G_LM(x)
 * by |x|^(-i_wa-j_wa) (S/field.cpp:1727-1762).  Other arguments and outputs as
 * trv_threept. */
int trv_threept_window(
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment,
  int ell1, int ell2, int ELL, int i_wa, int j_wa, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  const double* custom_edges,
  double alpha, double norm_factor, int wide_angle,
  int verbose, int deterministic, int part_rank, int part_count,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot);

/* One process per GPU (torchrun, mpirun, ...): attach an NCCL communicator to this
 * process.  Rank 0 obtains the 128-byte id with trv_comm_unique_id and ships it to
 * the other ranks (any channel: torch.distributed, MPI, a file); every rank then
 * calls trv_comm_init(nranks, rank, id) -- collective.  From then on an estimator
 * call with part_count == nranks ends with one all-reduce of its result vectors, so
 * every rank returns the complete measurement.  trv_allreduce sums a host buffer of
 * n doubles over the ranks (no-op without a communicator).  Status 4: NCCL absent.
 * Single-process callers need none of this: with several usable GPUs (TRV_GPU_MAXNUM,
 * S/monitor.cpp:258-324) the estimators spread over them by themselves. */
int trv_comm_unique_id(char id[128]);
int trv_comm_init(int nranks, int rank, const char id[128]);
int trv_comm_size(void);
int trv_allreduce(double* buf, long long n);
void trv_comm_finalize(void);
/* GPUs a single-process estimator call on this mesh would spread over. */
int trv_multi_device_count(const int* ngrid);
/* Estimator calls of this process that ran the distributed mesh phase (trvb_dmesh_*). */
long long trv_dmesh_call_count(void);
/* Estimator calls of this process whose mesh phase ran with the fused x pass
 * (trvb_box_fields_fused: one GPU, box bispectrum, throughput mode). */
long long trv_fused_mesh_call_count(void);

/* cudaStream_t of the most recently used estimator context (NULL before the
 * first call). */
void* trv_last_stream(void);
/* Drop the cached contexts (cuFFT plans, tables). */
void trv_release_contexts(void);

/* Phase timer of the estimator pipeline (bench aid); report is a JSON object
 * {"phase": seconds, ...}. */
void trv_profile_enable(int on);
int trv_profile_report(char* buf, int cap);

/* trv::calc_bispec_normalisation_from_particles (S/threept.cpp:96-136; host
 * only) or, from_mesh != 0, _from_mesh (S/threept.cpp:138-149). */
int trv_norm(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment, double* norm);

/* trv::calc_powspec_normalisation_from_particles (S/twopt.cpp:56-96; host only)
 * or, from_mesh != 0, _from_mesh (S/twopt.cpp:98-109). */
int trv_norm_powspec(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment, double* norm);

/* trv::MeshField pipeline for intermediate checks (S/field.cpp:569-1112,
 * 1496-1720, 1764-1785).  stage: 0 assignment, 1 + fourier_transform,
 * 2 + apply_assignment_compensation, 3 + inv_fourier_transform; 4 = assignment followed by
 * apply_wide_angle_pow_law_kernel at order (i_wa, j_wa) = (1, 2) (S/field.cpp:1727-1762).  Weights:
 * complex per particle (w_re, w_im) or NULL (unit).  field_out: 2*nmesh doubles. */
int trv_mesh(
  int stage, int subtract_mean, int interlace, int deterministic,
  int n, const double* x, const double* y, const double* z,
  const double* w_re, const double* w_im,
  const double* boxsize, const int* ngrid, const char* assignment,
  double* field_out, double* elapsed_assign_s);

/* trv::maths (S/maths.cpp:167-375): reduced spherical harmonics at n positions
 * (pos is (n, 3); out interleaved), spline-evaluated and exact j_l, Wigner 3-j. */
void trv_ylm(int ell, int m, const double* pos, int n, double* out);
/* SphericalHarmonicCalculator::store_reduced_spherical_harmonic_in_fourier_space
 * (fourier != 0) / _in_config_space (S/maths.cpp:222-302): y_lm at every mesh cell,
 * out holds 2 * n_x n_y n_z doubles (host tables; the estimators do not use them). */
int trv_ylm_mesh(int fourier, int ell, int m, const double* boxsize, const int* ngrid,
                 double* out);
void trv_sjl(int ell, const double* x, int n, double* out);
double trv_sjl_exact(int ell, double x);
double trv_w3j(int j1, int j2, int j3, int m1, int m2, int m3);
/* trv::calc_coupling_coeff_3pt (S/threept.cpp:65-89). */
double trv_coupling(int l1, int l2, int L, int m1, int m2, int M);

/* trv::Binning(params).set_bins() (S/dataobjs.cpp:134-249). */
int trv_binning(
  const char* space, const char* scheme, double bmin, double bmax, int nb,
  const double* boxsize, const int* ngrid,
  double* edges, double* centres, double* widths);

/* trv::ParameterSet::validate (S/parameters.cpp:466-1270): derived members. */
int trv_validate(
  const char* catalogue_type, const char* statistic_type,
  const char* assignment, const char* interlace, const char* form,
  int ell1, int ell2, int ELL, int num_bins, int idx_bin,
  double bin_min, double bin_max,
  char* shape_out, char* interlace_out, char* npoint_out, char* space_out,
  int* assignment_order);

#ifdef __cplusplus
}
#endif

#endif  /* TRV_CAPI_H_INCLUDED_ */
