/* trvb.h -- C-ABI device layer of triumvirate_b200 (libtrvb.so).
 *
 * The reference (MikeSWang/Triumvirate) has no C ABI: its bindings link the
 * C++ headers directly.  This header is the layer the
 * B200 build inserts UNDER the reference's C++ surface: every function takes
 * plain pointers and sizes only, returns an int status (0 = ok) and records a
 * message retrievable with trvb_last_error().  The host-side C++ classes in
 * triumvirate_b200/include/trv/ (trv::MeshField, trv::FieldStats,
 * trv::compute_bispec*, trv::compute_3pcf*) are thin callers of these.
 *
 * Each entry point cites the reference code it replaces (paths relative to
 * /root/reference/src/triumvirate/, S/ = src/, I/ = include/).
 *
 * Conventions
 *  - All arithmetic is IEEE fp64.  Complex values are interleaved (re, im).
 *  - A "mesh" is row-major [n0][n1][n2], x slowest (S/field.cpp:533-538).
 *  - trvb_mesh describes one device buffer:
 *      TRVB_REAL     n0*n1*n2 doubles                 (configuration space)
 *      TRVB_COMPLEX  n0*n1*n2 complex                 (either space)
 *      TRVB_HALF     n0*n1*(n2/2+1) complex           (Fourier space of a
 *                    real field; the missing half is implied by Hermitian
 *                    symmetry and reconstructed on access)
 *  - Device pointers are ordinary CUDA device addresses (cudaMalloc'ed by
 *    trvb_malloc, or e.g. a torch tensor's data_ptr()).
 *  - Work is enqueued on the context's stream; functions that return results
 *    to host memory synchronise that stream before returning.
 */
#ifndef TRVB_H_INCLUDED_
#define TRVB_H_INCLUDED_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trvb_ctx trvb_ctx;   /* grid + box + plans + tables, one GPU */
typedef struct trvb_cat trvb_cat;   /* device-resident particle catalogue   */

enum { TRVB_REAL = 0, TRVB_COMPLEX = 1, TRVB_HALF = 2 };

typedef struct {
  void* data;      /* device pointer */
  int layout;      /* TRVB_REAL | TRVB_COMPLEX | TRVB_HALF */
  double k0_add;   /* real value added to the k = 0 element whenever the mesh is
                      READ as a Fourier-space source; lets N_00(k) = dn_00(k) +
                      N delta_k0 (S/threept.cpp:1554-1558) share dn_00's buffer.
                      Ignored for configuration-space meshes and destinations. */
} trvb_mesh;

/* Particle weight kinds for assignment / catalogue sums. */
enum {
  TRVB_W_UNIT = 0,      /* 1               S/field.cpp:1208-1227            */
  TRVB_W_W = 1,         /* w               S/field.cpp:2033-2036            */
  TRVB_W_YLM_W = 2,     /* y_LM(los) w     S/field.cpp:1259-1269            */
  TRVB_W_CYLM_W2 = 3,   /* conj(y_LM) w^2  S/field.cpp:1379-1391            */
  TRVB_W_YLM_W3 = 4,    /* y_LM(los) w^3   S/threept.cpp:156-236 (sums only)*/
  TRVB_W_CUSTOM = 5     /* caller-supplied complex weights
                           (fftw_complex* weights, S/field.cpp:569-571)     */
};

/* ---- status ---------------------------------------------------------- */
const char* trvb_last_error(void);
const char* trvb_version(void);
/* Number of visible CUDA devices (0 when none / no driver); replaces the
 * probe of S/monitor.cpp:258-282. */
int trvb_device_count(void);
/* The calling thread's current CUDA device (cudaGetDevice / cudaSetDevice): used by
 * the host layer to pick the device when no environment variable names one, and to
 * restore the caller's device after an estimator call.  -1 / non-zero on failure. */
int trvb_current_device(void);
int trvb_set_current_device(int device);
/* Number of hand-written kernels launched by this library since the last
 * reset, and (separately) the number of cuFFT executions. */
long long trvb_launch_count(void);
long long trvb_fft_exec_count(void);
void trvb_launch_count_reset(void);
/* Number of cudaMalloc calls made by the caching device arena so far (steady
 * state: constant from one estimator call to the next). */
long long trvb_arena_malloc_count(void);

/* ---- context: replaces MeshField/FieldStats ctor state ----------------
 * (S/field.cpp:45-365: dr, dk, vol, vol_cell, FFT plans; S/field.cpp:2076).
 * assignment_order: 1 ngp, 2 cic, 3 tsc, 4 pcs (S/parameters.cpp:659-683). */
int trvb_ctx_create(trvb_ctx** ctx, int device, const int ngrid[3],
                    const double boxsize[3], int assignment_order);
void trvb_ctx_destroy(trvb_ctx* ctx);
/* on != 0: every reduction of this context avoids floating-point atomics so
 * that results are bit-reproducible from run to run (slower shot-noise
 * reduction); assignment determinism is selected per trvb_assign call. */
int trvb_ctx_set_deterministic(trvb_ctx* ctx, int on);
int trvb_ctx_sync(trvb_ctx* ctx);
void* trvb_ctx_stream(trvb_ctx* ctx);          /* cudaStream_t */
long long trvb_ctx_nmesh(const trvb_ctx* ctx);
/* Bytes of a mesh buffer in the given layout. */
size_t trvb_mesh_bytes(const trvb_ctx* ctx, int layout);

/* ---- memory (replaces S/arrayops.cpp:273-343 H2D/D2H helpers) ---------- */
/* Free/total device memory as seen by this library: the driver's free memory
 * measured once (at context creation, after an arena trim or a failed
 * allocation) and kept current with the arena's own allocations, plus the blocks
 * the arena holds for reuse.  trvb_mem_info_invalidate forces the next call to
 * measure again (S/monitor.cpp:117-156 gbytesMem* accounting plays this role in
 * the reference). */
int trvb_mem_info(trvb_ctx* ctx, size_t* free_bytes, size_t* total_bytes);
void trvb_mem_info_invalidate(int device);
/* Return the blocks the caching arena of `device` holds for reuse to the driver
 * (synchronises the device). */
void trvb_arena_release(int device);
int trvb_malloc(trvb_ctx* ctx, void** dptr, size_t bytes);
int trvb_free(trvb_ctx* ctx, void* dptr);
int trvb_memset0(trvb_ctx* ctx, void* dptr, size_t bytes);
int trvb_h2d(trvb_ctx* ctx, void* dptr, const void* hptr, size_t bytes);
int trvb_d2h(trvb_ctx* ctx, void* hptr, const void* dptr, size_t bytes);
int trvb_d2d(trvb_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- catalogue (replaces I/particles.hpp:63-90 on the device) ----------
 * Host (or device, if src_on_device) arrays of length n; `w` may be NULL
 * (unit weights), `los` is n x 3 row-major unit vectors or NULL
 * (I/dataobjs.hpp:186-188).  Positions must already be aligned in the box.
 * src_on_device == 2 BORROWS the device arrays x, y, z instead of copying them:
 * the caller keeps them alive and unchanged until trvb_cat_destroy.  Device arrays
 * that live on ANOTHER GPU than the context's (single-process multi-GPU runs) are
 * copied over with cudaMemcpyPeer instead of borrowed. */
int trvb_cat_create(trvb_ctx* ctx, trvb_cat** cat, long long n,
                    const double* x, const double* y, const double* z,
                    const double* w, const double* los, int src_on_device);
/* Same from the reference's AoS layout: 7 doubles per particle
 * {x, y, z, nz, ws, wc, w} (I/particles.hpp:63-69). */
/* Device arrays of the positions held by a catalogue (valid until it is destroyed). */
int trvb_cat_positions(const trvb_cat* cat, const double** x, const double** y,
                       const double** z);
int trvb_cat_create_aos(trvb_ctx* ctx, trvb_cat** cat, long long n,
                        const double* pdata, const double* los);
/* Periodic-box fast path for HOST coordinate arrays (pinned memory preferred): the
 * catalogue is uploaded in chunks on a copy stream and every chunk is tile-sorted and
 * spread onto `mesh` (zero-filled here; unit weights times `scale`, no density units)
 * while the next chunk is on the wire, so the counting sort and the assignment hide
 * behind the PCIe transfer.  Returns the catalogue (chunk-wise sorted: valid for the
 * throughput assignment of later calls).  Equivalent to trvb_cat_create followed by
 * trvb_assign(kind = TRVB_W_UNIT) up to summation order. */
int trvb_cat_create_assign(trvb_ctx* ctx, trvb_cat** out, long long n,
                           const double* x, const double* y, const double* z,
                           double scale, trvb_mesh mesh);
/* Attach caller-supplied complex weights (n interleaved (re, im) pairs, host
 * memory) for TRVB_W_CUSTOM. */
int trvb_cat_set_custom_weights(trvb_ctx* ctx, trvb_cat* cat,
                                const double* weights);
/* Forget the cached cell-sorted order, so that the next trvb_assign sorts the
 * catalogue again (measurement aid: the reference scatters in catalogue order,
 * S/field.cpp:618-1112, and has no sort to cache). */
void trvb_cat_invalidate_sort(trvb_cat* cat);
void trvb_cat_destroy(trvb_cat* cat);
long long trvb_cat_size(const trvb_cat* cat);
/* sum_i weight(kind, L, M)_i, complex, e.g. Sbar_LM (S/threept.cpp:156-236). */
int trvb_cat_sum(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M,
                 double out[2]);

/* ---- mesh assignment (replaces S/field.cpp:569-1112) -------------------
 * Scatter `scale * weight(kind, L, M)_i * Wx Wy Wz` (times 1/vol_cell when
 * density_units != 0, exactly `inv_vol_cell*w*Wx*Wy*Wz` as S/field.cpp:1044)
 * into `mesh` (TRVB_REAL allowed only for real-valued weight kinds, else
 * TRVB_COMPLEX).  accumulate == 0 zero-fills first (reset_density_field).
 * shifted != 0 assigns on the half-cell-shifted shadow mesh
 * (S/field.cpp:1056-1111).  mode: 0 = throughput (cell-sorted, atomics),
 * 1 = deterministic (per-cell ascending particle order, no FMA; bit-exact
 * against the reference run single-threaded). */
int trvb_assign(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M,
                double scale, int density_units, int accumulate, int shifted,
                int mode, trvb_mesh mesh);

/* mesh.re += c   (S/field.cpp:1235-1243 with c = -N/V). */
int trvb_mesh_add_const(trvb_ctx* ctx, trvb_mesh mesh, double c);
/* dst = a*dst + b*src, elementwise over same-layout meshes
 * (S/field.cpp:1309-1312, 1433-1436, 1358-1361). */
int trvb_mesh_axpby(trvb_ctx* ctx, trvb_mesh dst, double a, trvb_mesh src,
                    double b);
/* mesh(x) *= |x|^(-power) where |x| >= 1e-6, x = signed cell offset vector
 * (MeshField::apply_wide_angle_pow_law_kernel with power = i_wa + j_wa,
 * S/field.cpp:1727-1762). */
int trvb_mesh_pow_law(trvb_ctx* ctx, trvb_mesh mesh, int power);
/* sum_x Re(mesh)^order, order >= 2 (S/field.cpp:2056-2058); _pow3 is order 3. */
int trvb_mesh_sum_pow(trvb_ctx* ctx, trvb_mesh mesh, int order, double* out);
int trvb_mesh_sum_pow3(trvb_ctx* ctx, trvb_mesh mesh, double* out);

/* ---- transforms (replace S/field.cpp:1496-1720) ------------------------
 * Forward: dst(k) = FFT[prescale * src(x)] (sign -1, unnormalised).
 *   src TRVB_REAL  -> dst TRVB_HALF (out of place)
 *   src TRVB_COMPLEX -> dst TRVB_COMPLEX (in place when dst.data==src.data)
 * Inverse: dst(x) = IFFT[src(k)] (sign +1, unnormalised), HALF -> REAL or
 * COMPLEX -> COMPLEX.  The HALF source is overwritten by cuFFT's Z2D. */
int trvb_fft_forward(trvb_ctx* ctx, trvb_mesh src, trvb_mesh dst,
                     double prescale);
int trvb_fft_inverse(trvb_ctx* ctx, trvb_mesh src, trvb_mesh dst);
/* field(k) += add at the k = 0 mode (mean subtraction done in Fourier
 * space: FFT of a constant only populates k = 0). */
int trvb_kmesh_add_zero_mode(trvb_ctx* ctx, trvb_mesh kmesh, double add_re);
/* f = (f + exp(+i pi (mx+my+mz)) f_s)/2  (S/field.cpp:1618-1653). */
int trvb_interlace_combine(trvb_ctx* ctx, trvb_mesh kmesh, trvb_mesh kmesh_s);
/* f /= W(k), W = prod_i sinc(pi m_i/n_i)^order (S/field.cpp:1114-1201,
 * 1764-1785). */
int trvb_compensate(trvb_ctx* ctx, trvb_mesh kmesh);

/* ---- sub-grid ("shell grid") ------------------------------------------
 * The shell fields F_b(x) and the band-limited part of G(x) that enter
 * sum_x F_a F_b G (S/threept.cpp:1708-1717) only involve Fourier modes with
 * |m_i| <= mcut; the sum over the n^3 mesh equals (n^3/ns^3) times the sum
 * over an ns^3 mesh whenever ns > 4*mcut (no aliasing of the triple product),
 * or ns = n.  A sub-grid context shares the parent's box, tables and stream; it
 * is cached in and owned by the parent (repeated calls with the same extents
 * return the same handle, trvb_ctx_destroy on it is a no-op). */
int trvb_subgrid_create(trvb_ctx* parent, trvb_ctx** sub, const int nsub[3]);
/* With TRV_OVERLAP=1 a sub-grid context enqueues on its own stream.  trvb_ctx_fork makes
 * that stream wait for everything enqueued so far on the parent's (the Fourier meshes
 * the sub-grid kernels read); trvb_ctx_join makes the parent's stream wait for the
 * sub-grid's.  Between the two, work on `sub` and work on `parent` may run concurrently.
 * No-ops when both share a stream (the default). */
int trvb_ctx_fork(trvb_ctx* parent, trvb_ctx* sub);
int trvb_ctx_join(trvb_ctx* parent, trvb_ctx* sub);

/* Number of modes and sum of |k| per shell [edges[b], edges[b+1]), all bins
 * in one pass (the k_eff / nmodes side of S/field.cpp:1815-1847).  `fine`
 * != 0 applies the two-stage rule of S/field.cpp:2619,2674-2676 instead
 * (mode -> fine bin int(|k|/1e-5); fine bin q in shell iff edge_lo <= q*1e-5
 * < edge_hi). */
int trvb_shell_stats(trvb_ctx* ctx, const double* edges, int nbins, int fine,
                     long long* nmodes, double* ksum);

/* dst(x) = IFFT on `sub`'s grid of
 *   amp * 1{klo <= |k| < khi} * y_lm(khat) * src(k) / W(k)
 * (S/field.cpp:1792-1906 with amp = 1/nmodes folded in; src lives on the
 * parent grid `ctx`, dst is a TRVB_COMPLEX mesh of `sub`).  klo < 0 and
 * khi < 0 disable the shell test (all modes representable on `sub`):
 * with l = m = 0 and amp = 1/V this is G(x) (S/field.cpp:1764-1785,1657).
 * dst may be TRVB_REAL when src is TRVB_HALF, m = 0 and l is even (the filtered
 * spectrum is then Hermitian). */
int trvb_shell_ifft(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell,
                    int m, double klo, double khi, double amp, trvb_mesh dst);

/* Batched trvb_shell_ifft: the `nbins` shells [klo[q], khi[q]) of one (l, m),
 * each scaled by amp[q], written as nbins CONSECUTIVE meshes of `sub` starting
 * at device address `dst`, in one sparse pass over the low-|k| modes and one
 * batched inverse FFT.  dst_layout: TRVB_COMPLEX, or TRVB_REAL when the filtered
 * spectrum is Hermitian (src TRVB_HALF, m = 0, even l), which halves the work. */
int trvb_shell_ifft_batch(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell,
                          int m, const double* klo, const double* khi,
                          const double* amp, int nbins, void* dst, int dst_layout);

/* Slab form of trvb_shell_ifft_batch for the multi-GPU pair phase: the same `nbins`
 * REAL shell fields (src TRVB_HALF, m = 0, even l; klo[q] < 0 and khi[q] < 0 disable the
 * shell test, as for G), but only on the x-planes [x0, x0 + nx) of `sub`'s grid, written
 * as nbins consecutive REAL blocks [nx][ns1][ns2] at device address `dst`.  The inverse
 * transform is pruned: a direct x-DFT of the non-zero low-|k| modes for the nx planes
 * wanted (each mode belongs to one shell: no dense spectrum is ever built), then batched
 * 1-D transforms along y and z on those planes only -- so R ranks that split the x-planes
 * share the transform work of every shell instead of each transforming whole shells.
 * `sub` must be a true sub-grid of `ctx`. */
int trvb_shell_slab_batch(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell, int m,
                          const double* klo, const double* khi, const double* amp,
                          int nbins, int x0, int nx, void* dst);
/* 1 when the last (z) pass of trvb_shell_slab_batch runs as the hand-written pruned-input
 * complex-to-real kernel for a sub-grid of z-extent n2 (csrc/trvb_zpass.cuh), 0 when it
 * falls back to zero-padded lines + cuFFT.  The estimators prefer such extents when they
 * size the sub-grid in throughput mode. */
int trvb_shell_zpass_supported(int n2);

/* dst(x) = IFFT[ amp * j_l(|k| r) * y_lm(khat) * src(k) / W(k) ]
 * (S/field.cpp:1908-2010, amp = 1/V), spline table from trvb_sjl_table. */
int trvb_sjl_ifft(trvb_ctx* ctx, trvb_mesh src, int ell, int m, double r,
                  double amp, trvb_mesh dst);
/* The same for `nbins` radii r[q] of one (l, m), written as nbins CONSECUTIVE
 * TRVB_COMPLEX meshes starting at device address `dst`: y_lm src / W and |k| are
 * evaluated once, then each pass over them emits four weighted spectra. */
int trvb_sjl_ifft_batch(trvb_ctx* ctx, trvb_mesh src, int ell, int m,
                        const double* r, double amp, int nbins, void* dst);
/* Upload the natural-cubic-spline table of j_ell (I/maths.hpp:305-306,
 * S/maths.cpp:309-375): knots x_i = step*i, values y[i], coefficients c[i],
 * i < nsample; beyond split = step*(nsample-1) j_ell is evaluated directly. */
int trvb_sjl_table(trvb_ctx* ctx, int ell, const double* y, const double* c,
                   int nsample, double step);

/* ---- reductions -------------------------------------------------------
 * out[p] = sum_x A[ia[p]](x) * B[ib[p]](x) * G(x), complex, for p < npairs
 * (S/threept.cpp:1708-1717 and clones, all pairs in one pass over x).
 * A, B: arrays of na / nb device pointers to meshes of `ctx` in G's layout:
 * all TRVB_COMPLEX, or all TRVB_REAL (real fields: a quarter of the flops).
 * conj_b != 0 uses conj(B[ib[p]]) (complex meshes only): the fields of the mirror
 * harmonic (l, -m) of a real density are (-1)^(l+m) conj of those of (l, m), so one set
 * of inverse transforms serves both sides of an (m, -m) term. */
int trvb_gram_reduce(trvb_ctx* ctx, const void* const* A, int na,
                     const void* const* B, int nb, trvb_mesh G,
                     const int* ia, const int* ib, int npairs, int conj_b, double* out);

/* Binned pseudo-2pt statistics in Fourier space with the fine-bin rule
 * (S/field.cpp:2511-2703): per bin nmodes, mean |k|, mean
 * y_lm fa conj(fb)/C1 and mean y_lm S (C1/C1).  Empty bins follow
 * S/field.cpp:2688-2692 (k = centre, pk = sn = 0).  interlaced != 0 selects the
 * branch of interlaced meshes (two-point statistics only, S/field.cpp:2543-2552):
 * division by the product of the assignment windows, and the isotropic
 * shot-noise aliasing function of S/field.cpp:3504-3527. */
int trvb_twopt_fourier(trvb_ctx* ctx, trvb_mesh fa, trvb_mesh fb,
                       const double S[2], int ell, int m, int interlaced,
                       const double* edges, const double* centres, int nbins,
                       long long* nmodes, double* k, double* pk, double* sn);

/* xi(x) = IFFT[ (fa conj(fb)/C1 - S C1/C1) / V ]  (S/field.cpp:3273-3345,
 * 3018-3090); dst TRVB_COMPLEX on the same grid, or TRVB_REAL when fa and fb
 * are both TRVB_HALF (spectra of real fields) and Im S = 0: the product is then
 * Hermitian, xi is real and half the traffic suffices. */
int trvb_shot_xi(trvb_ctx* ctx, trvb_mesh fa, trvb_mesh fb, const double S[2],
                 int interlaced, trvb_mesh dst);

/* out[p] = vol_cell * sum_x j_la(ka[p] |x|) j_lb(kb[p] |x|) y_la,ma(xhat)
 * y_lb,mb(xhat) xi(x)   (S/field.cpp:3362-3393), all pairs in one pass.
 * xi: TRVB_COMPLEX or TRVB_REAL (see trvb_shot_xi). */
int trvb_shot_bispec_reduce(trvb_ctx* ctx, trvb_mesh xi, int la, int ma,
                            int lb, int mb, const double* ka, const double* kb,
                            int npairs, double* out);

/* Radially binned y_a y_b xi(x) with the fine-bin rule, dr_sample = 1
 * (S/field.cpp:3092-3189): per bin npairs, mean |x| and the doubly
 * normalised xi.  parity = (-1)^(l1+l2) (S/field.cpp:3181). */
int trvb_shot_3pcf_bin(trvb_ctx* ctx, trvb_mesh xi, int la, int ma, int lb,
                       int mb, const double* edges, const double* centres,
                       int nbins, double parity, long long* npairs, double* r,
                       double* xi_out);

/* Radially binned y_lm xi(x) of the two-point correlation function with the
 * fine-bin rule, dr_sample = 0.1, n_sample = 1e6 (S/field.cpp:2846-2931): per
 * bin npairs, mean |x| and mean y_lm xi; empty bins: r = centre, xi = 0. */
int trvb_twopt_config_bin(trvb_ctx* ctx, trvb_mesh xi, int ell, int m,
                          const double* edges, const double* centres, int nbins,
                          long long* npairs, double* r, double* xi_out);

/* ---- multi-GPU exchange (`trvb_allreduce`) ---------------
 * The mesh state is replicated on every GPU, the data-vector entries are dealt to
 * the ranks (trv::partition_owners) and every entry is produced by exactly one of
 * them; one all-reduce(sum) of the result vectors (<= 26 KB) over NCCL completes
 * the call.  The reference's own multi-GPU mode is single-process cuFFT-Xt
 * (S/field.cpp:212-235); it has no counterpart of these functions.
 * NCCL is bound at run time (dlopen libnccl.so.2): status 4 when it is absent.
 *   trvb_comm_unique_id  rank 0 creates the 128-byte id and ships it to the others
 *                        (ncclGetUniqueId)
 *   trvb_comm_create     collective over the `nranks` processes/threads, one GPU
 *                        each (ncclCommInitRank on `device`)
 *   trvb_allreduce       in-place sum of n doubles in HOST memory, staged through
 *                        the device; returns when the sum is back
 *   trvb_allreduce_device  the same for a DEVICE buffer, enqueued on the context's
 *                        stream without host synchronisation */
typedef struct trvb_comm trvb_comm;
int trvb_nccl_version(void);   /* 0 when NCCL cannot be loaded */
int trvb_comm_unique_id(char id[128]);
int trvb_comm_create(trvb_comm** comm, int device, int nranks, int rank, const char id[128]);
void trvb_comm_destroy(trvb_comm* comm);
int trvb_comm_size(const trvb_comm* comm);
int trvb_comm_rank(const trvb_comm* comm);
int trvb_allreduce(trvb_ctx* ctx, trvb_comm* comm, double* host_buf, long long n);
int trvb_allreduce_device(trvb_ctx* ctx, trvb_comm* comm, double* dev_buf, long long n);
/* Device of a context. */
int trvb_ctx_device(const trvb_ctx* ctx);
/* Exchanges on the context's stream (device buffers, no host synchronisation):
 *   trvb_comm_alltoall        block q (`n` doubles) of `send` goes to rank q, block q of
 *                             `recv` comes from rank q (grouped ncclSend/ncclRecv)
 *   trvb_comm_bcast_segments  segment s of `buf` (offset[s], count[s] doubles) is sent by
 *                             rank root[s] to all others (one group of ncclSend/ncclRecv, or of
 *                             ncclBroadcast above 64 MB) */
int trvb_comm_alltoall(trvb_ctx* ctx, trvb_comm* comm, const double* send, double* recv,
                       long long n);
int trvb_comm_bcast_segments(trvb_ctx* ctx, trvb_comm* comm, double* buf, int nseg,
                             const int* root, const long long* offset, const long long* count);

/* ---- distributed mesh phase of the periodic-box estimators ------------------
 * The part of a box estimator call that a replicated grid cannot shrink -- particle
 * assignment (S/field.cpp:987-1112) and the two full-grid FFTs (S/field.cpp:1496-1720) --
 * spread over the ranks of a communicator: x-slabs of the configuration-space mesh,
 * k_y-slabs of the Fourier-space mesh, one all-to-all per transform.  The counterpart of
 * the reference's cuFFT-Xt mode (S/field.cpp:212-235), across processes.  Needs
 * n0 % R == 0 and n1 % R == 0 (trvb_dmesh_supported).  All calls below except
 * _supported, _planes and _forget_lowk are collective over the communicator.
 *   trvb_dmesh_density      delta n(k) of n unit-weight particles (device arrays holding
 *                           the whole catalogue on every rank); k0_add is added at k = 0
 *   trvb_dmesh_gather_lowk  the modes the grid of `sub` represents -> HALF mesh `dst` of
 *                           sub's extents on every rank; from then on a trvb_mesh at that
 *                           address is read as a spectrum of the BIG grid by every function
 *                           that takes a Fourier-space source (shell transforms, binned
 *                           two-point statistics), until _forget_lowk
 *   trvb_dmesh_shot_xi      xi(r) of trvb_shot_xi on this rank's planes
 *                           (REAL, [nx][n1][n2]) for fa = dn + add_a delta_k0,
 *                           fb = dn + add_b delta_k0
 *   trvb_shot_bispec_reduce_slab  trvb_shot_bispec_reduce from those planes: the radial
 *                           histogram is summed over the ranks, every rank gets all pairs */
typedef struct trvb_dmesh trvb_dmesh;
int trvb_dmesh_supported(const trvb_ctx* ctx, int nranks);
long long trvb_dmesh_call_count(void);   /* trvb_dmesh_density calls of this process */
int trvb_dmesh_create(trvb_ctx* ctx, trvb_comm* comm, trvb_dmesh** out);
/* The context's own distributed-mesh state for `comm` (plans and the assignment window are
 * built once per context; destroyed with it). */
int trvb_dmesh_get(trvb_ctx* ctx, trvb_comm* comm, trvb_dmesh** out);
void trvb_dmesh_destroy(trvb_dmesh* dm);
int trvb_dmesh_planes(const trvb_dmesh* dm, int* x0, int* nx);
int trvb_dmesh_density(trvb_dmesh* dm, long long n, const double* x, const double* y,
                       const double* z, double k0_add);
int trvb_dmesh_gather_lowk(trvb_dmesh* dm, trvb_ctx* sub, trvb_mesh dst);
void trvb_dmesh_forget_lowk(trvb_dmesh* dm);
int trvb_dmesh_shot_xi(trvb_dmesh* dm, double add_a, double add_b, const double S[2],
                       double* xi_planes);
int trvb_shot_bispec_reduce_slab(trvb_ctx* ctx, const double* xi_planes, int x0, int nx,
                                 trvb_comm* comm, int la, int ma, int lb, int mb,
                                 const double* ka, const double* kb, int npairs, double* out);

/* ---- box mesh phase on one GPU, x passes fused (csrc/trvb_xpass.cu) -----------
 * For the periodic-box bispectrum the full-grid spectrum delta n(k) is read twice and
 * never again: for the low-|k| modes the pair phase works on (the grid of `sub`) and for
 * xi(r) = IFFT[(fa conj(fb) / C1 - S) / V] of the shot noise (S/threept.cpp:1554-1558,
 * S/field.cpp:1496-1655, 3273-3345).  This call does both from the REAL mesh `x`:
 * 2-D D2Z of the x-planes (cuFFT), ONE hand-written pass over the columns along x
 * (forward FFT, low-|k| modes stored to `lowk`, spectrum, inverse FFT, in place), 2-D Z2D
 * of the planes into `xi`.  fa = FFT[x] + add_a delta_k0, fb = FFT[x] + add_b delta_k0;
 * `lowk` (HALF, sub's extents) receives fa and is registered as the context's low-|k|
 * view exactly as trvb_dmesh_gather_lowk does, until trvb_ctx_forget_lowk.  Throughput
 * mode only (the deterministic mode keeps the 3-D cuFFT transforms); n0 must be a power
 * of two in 32 .. 2048 (trvb_box_fields_fused_supported). */
int trvb_box_fields_fused_supported(const trvb_ctx* ctx);
int trvb_box_fields_fused(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh x, double add_a, double add_b,
                          const double S[2], trvb_mesh lowk, trvb_mesh xi);
void trvb_ctx_forget_lowk(trvb_ctx* ctx);
long long trvb_box_fields_fused_call_count(void);   /* calls of this process */
/* With TRV_XPASS_TRACE=1 in the environment: CUDA-event times (ms) of the last call's 2-D
 * D2Z, x pass (k_xpass_fused) and 2-D Z2D. */
void trvb_box_fields_fused_last_ms(double out[3]);

#ifdef __cplusplus
}
#endif

#endif  /* TRVB_H_INCLUDED_ */
