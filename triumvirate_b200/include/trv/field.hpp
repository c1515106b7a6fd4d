// field.hpp -- trv::MeshField and trv::FieldStats on the device.
//
// Same public surface as the reference (I/field.hpp:141-426, 584-774), with
// the mesh resident in HBM: every method enqueues hand-written kernels (and
// cuFFT) through the C-ABI layer include/trvb.h; nothing falls back to the
// host.  Differences from the reference, all deliberate:
//   * `field` is a HOST MIRROR, refreshed by sync_host() / operator[] after
//     device-side changes (the reference's `field` is the working buffer).
//   * y_lm is given by its (ell, m) orders and evaluated inside the kernels;
//     the N^3 y_lm tables of S/maths.cpp:222-302 are never materialised.
//   * CUDA / cuFFT failures throw trv::sys::DeviceError.
#ifndef TRV_B200_FIELD_HPP_
#define TRV_B200_FIELD_HPP_

#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "dataobjs.hpp"
#include "maths.hpp"
#include "monitor.hpp"
#include "parameters.hpp"
#include "particles.hpp"
#include "trvb.h"

// Callers of the reference API release FFTW's global state at exit (fftw_cleanup[_threads],
// T/main/triumvirate.cpp:905-911).  The device build holds none: no-ops keep them compiling.
inline void fftw_cleanup() {}
inline void fftw_cleanup_threads() {}

namespace trv {

namespace dev {

/// Throw trv::sys::DeviceError carrying trvb_last_error() if `status` != 0.
void check(int status, const char* what);

/// Optional phase timer (off by default): when enabled, mark() synchronises
/// the context stream and charges the time since the previous mark to `phase`.
void profile_enable(bool on);
bool profile_enabled();
void profile_reset();
void profile_mark(trvb_ctx* ctx, const char* phase);
std::string profile_report();

/// GPU used by this process's estimator calls: TRV_GPU_DEVICE, else LOCAL_RANK, else
/// the calling thread's current CUDA device; out-of-range ids throw DeviceError.
int select_device();
/// Restores the calling thread's current CUDA device when it goes out of scope (the
/// device layer makes the context's GPU current while it works).
class DeviceScope {
 public:
  DeviceScope();
  ~DeviceScope();
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
 private:
  int prev_;
};
/// Communicator of a one-process-per-GPU run (trv_comm_init / trv::dev::comm_init), or
/// null.  When it spans `part_count` ranks the estimators finish with one all-reduce of
/// their result vectors over NCCL, so every rank returns the complete measurement.
trvb_comm* process_comm();
/// Rank 0 creates the 128-byte id (ncclGetUniqueId) and ships it to the other ranks by
/// whatever channel launched them; every rank then calls comm_init (collective).
void comm_unique_id(char id[128]);
void comm_init(int nranks, int rank, const char id[128]);
void comm_finalize();
/// In-place sum over the ranks of the process communicator (no-op without one).
void allreduce(trvb_ctx* ctx, double* buf, long long n);
/// Number of GPUs a SINGLE-process estimator call spreads over: every usable device
/// (trv::sys::get_gpu_count(), i.e. capped by TRV_GPU_MAXNUM, S/monitor.cpp:258-324)
/// unless the process is pinned to one (TRV_GPU_DEVICE, LOCAL_RANK), a process
/// communicator exists, TRV_GPU_MULTI=0, or the mesh is small (< 256^3 cells).
int multi_device_count(const trv::ParameterSet& params);
/// Shared device context for one (device, ngrid, boxsize, assignment order).
std::shared_ptr<trvb_ctx> acquire_context(const trv::ParameterSet& params);
/// Most recently acquired context (null if none is alive); lets callers put
/// CUDA events on the stream the estimators enqueue on.
trvb_ctx* last_context();
/// Drop the contexts kept alive between calls (plans, tables, scratch).
void release_contexts();

/// Device catalogue (positions, w, optional LOS) with RAII ownership.
class Catalogue {
 public:
  Catalogue(std::shared_ptr<trvb_ctx> ctx, ParticleCatalogue& particles,
            LineOfSight* los, bool need_weights);
  /// From separate coordinate arrays in host or device memory (`w`, `los`
  /// may be null); no AoS staging copy.
  Catalogue(std::shared_ptr<trvb_ctx> ctx, long long n, const double* x,
            const double* y, const double* z, const double* w,
            const double* los, bool on_device);
  /// Adopts a device catalogue built by the device layer (trvb_cat_create_assign).
  Catalogue(std::shared_ptr<trvb_ctx> ctx, trvb_cat* adopted) : ctx_(ctx), cat_(adopted) {}
  ~Catalogue();
  Catalogue(const Catalogue&) = delete;
  Catalogue& operator=(const Catalogue&) = delete;
  trvb_cat* get() { return cat_; }
  /// Attach arbitrary complex per-particle weights (interleaved re, im).
  void set_custom_weights(const double* weights);
 private:
  std::shared_ptr<trvb_ctx> ctx_;
  trvb_cat* cat_ = nullptr;
};

/// One device mesh buffer with RAII ownership.
class Mesh {
 public:
  Mesh() = default;
  Mesh(std::shared_ptr<trvb_ctx> owner, trvb_ctx* grid, int layout);
  ~Mesh();
  Mesh(Mesh&& other) noexcept;
  Mesh& operator=(Mesh&& other) noexcept;
  Mesh(const Mesh&) = delete;
  Mesh& operator=(const Mesh&) = delete;
  trvb_mesh view() const {
    trvb_mesh m; m.data = data_; m.layout = layout_; m.k0_add = 0.; return m;
  }
  /// Fourier-space view whose k = 0 element reads as stored + `add`.
  trvb_mesh view_k0(double add) const { trvb_mesh m = view(); m.k0_add = add; return m; }
  void* data() const { return data_; }
  int layout() const { return layout_; }
  bool empty() const { return data_ == nullptr; }
  void release();
 private:
  std::shared_ptr<trvb_ctx> owner_;
  void* data_ = nullptr;
  int layout_ = TRVB_COMPLEX;
  size_t bytes_ = 0;
};

}  // namespace dev

class FieldStats;

class MeshField {
 public:
  trv::ParameterSet params;
  std::string name;
  double (*field)[2] = nullptr;   ///< host mirror of the complex mesh
  double dr[3];
  double dk[3];
  double vol;
  double vol_cell;

  explicit MeshField(trv::ParameterSet& params, bool plan_ini = true,
                     const std::string& name = "mesh-field");
  ~MeshField();
  MeshField(const MeshField&) = delete;
  MeshField& operator=(const MeshField&) = delete;

  void reset_density_field();
  /// Element access through the host mirror (downloads when stale).
  const double (&operator[](long long gid))[2];
  /// Refresh / push the host mirror explicitly.
  void sync_host();
  void sync_device();

  // -- Mesh assignment (S/field.cpp:569-1112) --
  void assign_weighted_field_to_mesh(ParticleCatalogue& particles, double (*weights)[2]);

  // -- Field computations (S/field.cpp:1208-1489) --
  void compute_unweighted_field(ParticleCatalogue& particles);
  void compute_unweighted_field_fluctuations_insitu(ParticleCatalogue& particles);
  void compute_ylm_wgtd_field(
    ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
    LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m);
  void compute_ylm_wgtd_field(
    ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m);
  void compute_ylm_wgtd_quad_field(
    ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
    LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m);
  void compute_ylm_wgtd_quad_field(
    ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m);

  // -- Transforms (S/field.cpp:1496-1720) --
  void fourier_transform();
  void inv_fourier_transform();

  // -- Field operations (S/field.cpp:1764-1785) --
  /// field(x) *= |x|^(-i_wa - j_wa) where |x| >= 1e-6 (params.i_wa, params.j_wa), x the
  /// signed cell offset vector (S/field.cpp:1727-1762).
  void apply_wide_angle_pow_law_kernel();
  void apply_assignment_compensation();

  // -- One-point statistics (S/field.cpp:1792-2010); y_lm by orders --
  void inv_fourier_transform_ylm_wgtd_field_band_limited(
    MeshField& field_fourier, int ell, int m,
    double k_lower, double k_upper, double& k_eff, int& nmodes);
  void inv_fourier_transform_sjl_ylm_wgtd_field(
    MeshField& field_fourier, int ell, int m,
    trv::maths::SphericalBesselCalculator& sjl, double r);

  // -- Misc (S/field.cpp:2017-2065) --
  double calc_grid_based_powlaw_norm(ParticleCatalogue& particles, int order);

  /// Device view of the primary mesh (COMPLEX layout).
  trvb_mesh device_view() const { return mesh_.view(); }
  trvb_ctx* context() const { return ctx_.get(); }

 private:
  friend class FieldStats;
  std::shared_ptr<trvb_ctx> ctx_;
  dev::Mesh mesh_;
  dev::Mesh mesh_s_;          // shadow mesh when interlacing
  bool host_stale_ = true;    // device has newer data than the mirror
  void assign_kind(ParticleCatalogue& particles, LineOfSight* los, int kind,
                   int ell, int m, double scale, bool accumulate);
};

class FieldStats {
 public:
  std::vector<int> nmodes;
  std::vector<int> npairs;
  std::vector<double> k;
  std::vector<double> r;
  std::vector< std::complex<double> > sn;
  std::vector< std::complex<double> > pk;
  std::vector< std::complex<double> > xi;

  explicit FieldStats(trv::ParameterSet& params, bool plan_ini = true);
  ~FieldStats() = default;

  void reset_stats();

  /// Mesh vectors (wavevectors for binning.space "fourier", separation vectors for
  /// "config") listed bin by bin, optionally written to `save_file`
  /// (S/field.cpp:2317-2509).  Host-side; bins ascending, row-major cell order within a bin.
  trv::BinnedVectors record_binned_vectors(trv::Binning& binning, const std::string& save_file = {});

  /// S/field.cpp:2511-2703.
  void compute_ylm_wgtd_2pt_stats_in_fourier(
    MeshField& field_a, MeshField& field_b, std::complex<double> shotnoise_amp,
    int ell, int m, trv::Binning& kbinning);

  /// S/field.cpp:2705-2945.
  void compute_ylm_wgtd_2pt_stats_in_config(
    MeshField& field_a, MeshField& field_b, std::complex<double> shotnoise_amp,
    int ell, int m, trv::Binning& rbinning);

  /// S/field.cpp:2947-3196; y_lm tables replaced by their orders.
  void compute_uncoupled_shotnoise_for_3pcf(
    MeshField& field_a, MeshField& field_b,
    int ell_a, int m_a, int ell_b, int m_b,
    std::complex<double> shotnoise_amp, trv::Binning& rbinning);

  /// S/field.cpp:3198-3396; y_lm tables replaced by their orders.
  std::complex<double> compute_uncoupled_shotnoise_for_bispec_per_bin(
    MeshField& field_a, MeshField& field_b,
    int ell_a, int m_a, int ell_b, int m_b,
    trv::maths::SphericalBesselCalculator& sj_a,
    trv::maths::SphericalBesselCalculator& sj_b,
    std::complex<double> shotnoise_amp, double k_a, double k_b);

 private:
  trv::ParameterSet params;
  std::shared_ptr<trvb_ctx> ctx_;
  void resize_stats(int num_bins);
  bool if_fields_compatible(MeshField& field_a, MeshField& field_b);
  int interlaced_() const { return this->params.interlace == "true" ? 1 : 0; }
};

}  // namespace trv

#endif  // TRV_B200_FIELD_HPP_
