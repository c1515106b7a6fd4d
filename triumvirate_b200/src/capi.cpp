// capi.cpp -- flat C entry points over the trv:: C++ API (libtrv_b200.so),
// the binding surface for Python (ctypes) and other FFI callers.  It plays
// the role of the reference's Cython layer (T/_threept.pyx:128-248,
// T/_particles.pyx:17-34, T/parameters.pyx:395-470): marshal arrays into
// trv::ParticleCatalogue / trv::ParameterSet / trv::Binning, call the
// estimator, copy the result vectors back.  Unlike the reference's `compute_*`
// externs (no `except +`), every C++ exception is caught and reported.
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "trv/dataobjs.hpp"
#include "trv/field.hpp"
#include "trv/maths.hpp"
#include "trv/monitor.hpp"
#include "trv/parameters.hpp"
#include "trv/particles.hpp"
#include "trv/threept.hpp"
#include "trv/twopt.hpp"

namespace {

thread_local std::string g_capi_err;

template <class F>
int guarded(F f) {
  try {
    trv::dev::DeviceScope restore_callers_device;
    f();
    return 0;
  } catch (const trv::sys::DeviceError& e) {
    g_capi_err = e.what(); return 3;
  } catch (const std::invalid_argument& e) {
    g_capi_err = e.what(); return 2;
  } catch (const std::exception& e) {
    g_capi_err = e.what(); return 1;
  }
}

void set_params(
  trv::ParameterSet& p, const char* catalogue_type, const char* statistic_type,
  const double* boxsize, const int* ngrid, const char* assignment,
  const char* interlace, int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  int verbose, int deterministic, int part_rank, int part_count
) {
  p.catalogue_type = catalogue_type;
  p.statistic_type = statistic_type;
  for (int ax = 0; ax < 3; ax++) { p.boxsize[ax] = boxsize[ax]; p.ngrid[ax] = ngrid[ax]; }
  p.assignment = assignment;
  p.interlace = interlace;
  p.ell1 = ell1; p.ell2 = ell2; p.ELL = ELL;
  p.form = form; p.idx_bin = idx_bin;
  p.binning = binning; p.bin_min = bin_min; p.bin_max = bin_max; p.num_bins = num_bins;
  p.verbose = verbose; p.progbar = "false";
  p.deterministic = deterministic;
  p.part_rank = part_rank; p.part_count = part_count;
  p.validate(false);
}

}  // namespace

extern "C" {

const char* trv_last_error() { return g_capi_err.c_str(); }

int trv_gpu_count() { return trv::sys::get_gpu_count(); }

int trv_partition_owners(const char* form, int ell1, int ell2, int idx_bin, int num_bins,
                         int world, int* owner, int* dim) {
  try {
    trv::ParameterSet p;
    p.form = form; p.ell1 = ell1; p.ell2 = ell2; p.idx_bin = idx_bin;
    // Form -> shape as validate() derives it (S/parameters.cpp:829-849).
    p.shape = (p.form == "full" && ell1 == ell2) ? "triu" : p.form;
    const std::vector<int> own = trv::partition_owners(p, num_bins, world);
    *dim = static_cast<int>(own.size());
    for (size_t i = 0; i < own.size(); i++) owner[i] = own[i];
    return 0;
  } catch (const std::exception& e) {
    g_capi_err = e.what();
    return 1;
  }
}

void trv_counters(int* count_fft, int* count_ifft, double* gib_gpu_max) {
  *count_fft = trv::sys::count_fft; *count_ifft = trv::sys::count_ifft;
  *gib_gpu_max = trv::sys::gbytesMaxMemGPU;
}

/// Three-point estimators.  Arrays as in the reference's Cython bindings:
/// six float64 columns per catalogue (nz/ws/wc may be null: 0/1/1), LOS as
/// contiguous (n, 3).  `stat`: "bispec" | "3pcf"; `catalogue_type`: "sim"
/// (trv::compute_*_in_gpp_box) | "survey" (trv::compute_bispec / compute_3pcf).
/// Outputs hold dv_dim entries (raw/shot: interleaved re, im).
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
  int* n1, int* n2, double* raw, double* shot, double* elapsed_s
) {
  return guarded([&]() {
    trv::ParameterSet params;
    set_params(params, catalogue_type, stat, boxsize, ngrid, assignment, interlace,
               ell1, ell2, ELL, form, idx_bin, binning, bin_min, bin_max, num_bins,
               verbose, deterministic, part_rank, part_count);
    trv::Binning bins(params);
    if (custom_edges != nullptr) {
      bins.set_bins(std::vector<double>(custom_edges, custom_edges + num_bins + 1));
    } else {
      bins.set_bins();
    }
    const bool survey = std::string(catalogue_type) == "survey";
    trv::ParticleCatalogue data(verbose), rand(verbose);
    data.load_particle_arrays(nd, xd, yd, zd, nzd, wsd, wcd);
    if (survey) rand.load_particle_arrays(nr, xr, yr, zr, nzr, wsr, wcr);
    trv::LineOfSight* ld = (trv::LineOfSight*)los_d;
    trv::LineOfSight* lr = (trv::LineOfSight*)los_r;

    auto t0 = std::chrono::steady_clock::now();
    if (std::string(stat) == "bispec") {
      trv::BispecMeasurements out = survey
        ? trv::compute_bispec(data, rand, ld, lr, params, bins, norm_factor)
        : trv::compute_bispec_in_gpp_box(data, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c1_bin[i] = out.k1_bin[i]; c2_bin[i] = out.k2_bin[i];
        c1_eff[i] = out.k1_eff[i]; c2_eff[i] = out.k2_eff[i];
        n1[i] = out.nmodes_1[i]; n2[i] = out.nmodes_2[i];
        raw[2*i] = out.bk_raw[i].real(); raw[2*i+1] = out.bk_raw[i].imag();
        shot[2*i] = out.bk_shot[i].real(); shot[2*i+1] = out.bk_shot[i].imag();
      }
    } else {
      trv::ThreePCFMeasurements out = survey
        ? trv::compute_3pcf(data, rand, ld, lr, params, bins, norm_factor)
        : trv::compute_3pcf_in_gpp_box(data, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c1_bin[i] = out.r1_bin[i]; c2_bin[i] = out.r2_bin[i];
        c1_eff[i] = out.r1_eff[i]; c2_eff[i] = out.r2_eff[i];
        n1[i] = out.npairs_1[i]; n2[i] = out.npairs_2[i];
        raw[2*i] = out.zeta_raw[i].real(); raw[2*i+1] = out.zeta_raw[i].imag();
        shot[2*i] = out.zeta_shot[i].real(); shot[2*i+1] = out.zeta_shot[i].imag();
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  });
}

/// Two-point estimators (trv::compute_powspec / compute_corrfunc and their
/// `_in_gpp_box` forms, trv::compute_corrfunc_window; S/twopt.cpp:388-901; bindings
/// T/_twopt.pyx).  `stat` = "powspec" | "2pcf" | "2pcf-win"; for "2pcf-win" the
/// catalogue is passed in the `r` slot with `alpha`.
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
  double* elapsed_s
) {
  return guarded([&]() {
    const std::string st(stat);
    trv::ParameterSet params;
    set_params(params, catalogue_type, stat, boxsize, ngrid, assignment, interlace,
               ELL, 0, ELL, "diag", 0, binning, bin_min, bin_max, num_bins,
               verbose, deterministic, 0, 1);
    trv::Binning bins(params);
    if (custom_edges != nullptr) {
      bins.set_bins(std::vector<double>(custom_edges, custom_edges + num_bins + 1));
    } else {
      bins.set_bins();
    }
    const bool survey = std::string(catalogue_type) == "survey";
    const bool window = st == "2pcf-win";
    trv::ParticleCatalogue data(verbose), rand(verbose);
    if (!window) data.load_particle_arrays(nd, xd, yd, zd, nzd, wsd, wcd);
    if (survey || window) rand.load_particle_arrays(nr, xr, yr, zr, nzr, wsr, wcr);
    trv::LineOfSight* ld = (trv::LineOfSight*)los_d;
    trv::LineOfSight* lr = (trv::LineOfSight*)los_r;

    auto t0 = std::chrono::steady_clock::now();
    if (st == "powspec") {
      trv::PowspecMeasurements out = survey
        ? trv::compute_powspec(data, rand, ld, lr, params, bins, norm_factor)
        : trv::compute_powspec_in_gpp_box(data, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c_bin[i] = out.kbin[i]; c_eff[i] = out.keff[i]; count[i] = out.nmodes[i];
        raw[2*i] = out.pk_raw[i].real(); raw[2*i+1] = out.pk_raw[i].imag();
        shot[2*i] = out.pk_shot[i].real(); shot[2*i+1] = out.pk_shot[i].imag();
      }
    } else if (window) {
      trv::TwoPCFWindowMeasurements out = trv::compute_corrfunc_window(
        rand, lr, params, bins, alpha, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c_bin[i] = out.rbin[i]; c_eff[i] = out.reff[i]; count[i] = out.npairs[i];
        raw[2*i] = out.xi[i].real(); raw[2*i+1] = out.xi[i].imag();
        shot[2*i] = 0.; shot[2*i+1] = 0.;
      }
    } else {
      trv::TwoPCFMeasurements out = survey
        ? trv::compute_corrfunc(data, rand, ld, lr, params, bins, norm_factor)
        : trv::compute_corrfunc_in_gpp_box(data, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c_bin[i] = out.rbin[i]; c_eff[i] = out.reff[i]; count[i] = out.npairs[i];
        raw[2*i] = out.xi[i].real(); raw[2*i+1] = out.xi[i].imag();
        shot[2*i] = 0.; shot[2*i+1] = 0.;
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
  });
}

/// 3PCF window function of a random catalogue (trv::compute_3pcf_window,
/// S/threept.cpp:2621-3077; binding T/_threept.pyx:276-314).  `wide_angle` != 0
/// applies the r^{-i_wa-j_wa} kernel to G_LM.
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
  int* n1, int* n2, double* raw, double* shot
) {
  return guarded([&]() {
    trv::ParameterSet params;
    params.i_wa = i_wa; params.j_wa = j_wa;
    set_params(params, "random", wide_angle ? "3pcf-win-wa" : "3pcf-win", boxsize, ngrid,
               assignment, "false", ell1, ell2, ELL, form, idx_bin, binning, bin_min, bin_max,
               num_bins, verbose, deterministic, part_rank, part_count);
    trv::Binning bins(params);
    if (custom_edges != nullptr) {
      bins.set_bins(std::vector<double>(custom_edges, custom_edges + num_bins + 1));
    } else {
      bins.set_bins();
    }
    trv::ParticleCatalogue rand(verbose);
    rand.load_particle_arrays(nr, xr, yr, zr, nzr, wsr, wcr);
    trv::ThreePCFWindowMeasurements out = trv::compute_3pcf_window(
      rand, (trv::LineOfSight*)los_r, params, bins, alpha, norm_factor, wide_angle != 0);
    *dim = out.dim;
    for (int i = 0; i < out.dim; i++) {
      c1_bin[i] = out.r1_bin[i]; c2_bin[i] = out.r2_bin[i];
      c1_eff[i] = out.r1_eff[i]; c2_eff[i] = out.r2_eff[i];
      n1[i] = out.npairs_1[i]; n2[i] = out.npairs_2[i];
      raw[2*i] = out.zeta_raw[i].real(); raw[2*i+1] = out.zeta_raw[i].imag();
      shot[2*i] = out.zeta_shot[i].real(); shot[2*i+1] = out.zeta_shot[i].imag();
    }
  });
}

/// Periodic-box estimators from coordinate arrays that may already live in
/// device memory (`on_device` != 0: x, y, z are CUDA device pointers, e.g. a
/// torch tensor's data_ptr()); unit weights.  No AoS staging copy.
int trv_threept_box_arrays(
  const char* stat, long long n, const double* x, const double* y, const double* z,
  int on_device,
  const double* boxsize, const int* ngrid, const char* assignment,
  int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  double norm_factor, int verbose, int deterministic, int part_rank, int part_count,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot
) {
  return guarded([&]() {
    trv::ParameterSet params;
    set_params(params, "sim", stat, boxsize, ngrid, assignment, "false",
               ell1, ell2, ELL, form, idx_bin, binning, bin_min, bin_max, num_bins,
               verbose, deterministic, part_rank, part_count);
    trv::Binning bins(params);
    bins.set_bins();
    if (std::string(stat) == "bispec") {
      trv::BispecMeasurements out = trv::compute_bispec_in_gpp_box(
        n, x, y, z, on_device != 0, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c1_bin[i] = out.k1_bin[i]; c2_bin[i] = out.k2_bin[i];
        c1_eff[i] = out.k1_eff[i]; c2_eff[i] = out.k2_eff[i];
        n1[i] = out.nmodes_1[i]; n2[i] = out.nmodes_2[i];
        raw[2*i] = out.bk_raw[i].real(); raw[2*i+1] = out.bk_raw[i].imag();
        shot[2*i] = out.bk_shot[i].real(); shot[2*i+1] = out.bk_shot[i].imag();
      }
    } else {
      trv::ThreePCFMeasurements out = trv::compute_3pcf_in_gpp_box(
        n, x, y, z, on_device != 0, params, bins, norm_factor);
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c1_bin[i] = out.r1_bin[i]; c2_bin[i] = out.r2_bin[i];
        c1_eff[i] = out.r1_eff[i]; c2_eff[i] = out.r2_eff[i];
        n1[i] = out.npairs_1[i]; n2[i] = out.npairs_2[i];
        raw[2*i] = out.zeta_raw[i].real(); raw[2*i+1] = out.zeta_raw[i].imag();
        shot[2*i] = out.zeta_shot[i].real(); shot[2*i+1] = out.zeta_shot[i].imag();
      }
    }
  });
}

/// cudaStream_t of the most recently used estimator context (null before the
/// first call), so callers can bracket calls with CUDA events on it.
void* trv_last_stream() {
  trvb_ctx* c = trv::dev::last_context();
  return c ? trvb_ctx_stream(c) : nullptr;
}

void trv_release_contexts() { trv::dev::release_contexts(); }

int trv_comm_unique_id(char id[128]) {
  return guarded([&]() { trv::dev::comm_unique_id(id); });
}

int trv_comm_init(int nranks, int rank, const char id[128]) {
  return guarded([&]() { trv::dev::comm_init(nranks, rank, id); });
}

int trv_comm_size(void) {
  trvb_comm* c = trv::dev::process_comm();
  return c ? trvb_comm_size(c) : 1;
}

int trv_allreduce(double* buf, long long n) {
  return guarded([&]() {
    if (trv::dev::process_comm() == nullptr) return;
    trv::ParameterSet p;
    for (int ax = 0; ax < 3; ax++) { p.boxsize[ax] = 1.; p.ngrid[ax] = 4; }
    p.assignment_order = 1;
    trvb_ctx* ctx = trv::dev::last_context();
    std::shared_ptr<trvb_ctx> hold;
    if (ctx == nullptr) { hold = trv::dev::acquire_context(p); ctx = hold.get(); }
    trv::dev::allreduce(ctx, buf, n);
  });
}

void trv_comm_finalize(void) { trv::dev::comm_finalize(); }

long long trv_dmesh_call_count(void) { return trvb_dmesh_call_count(); }
long long trv_fused_mesh_call_count(void) { return trvb_box_fields_fused_call_count(); }

int trv_multi_device_count(const int* ngrid) {
  trv::ParameterSet p;
  for (int ax = 0; ax < 3; ax++) { p.boxsize[ax] = 1.; p.ngrid[ax] = ngrid[ax]; }
  p.nmesh = (long long)ngrid[0] * ngrid[1] * ngrid[2];
  return trv::dev::multi_device_count(p);
}

/// Phase timer of the estimator pipeline (development / bench aid).
void trv_profile_enable(int on) { trv::dev::profile_enable(on != 0); }

int trv_profile_report(char* buf, int cap) {
  const std::string rep = trv::dev::profile_report();
  std::strncpy(buf, rep.c_str(), cap - 1);
  buf[cap - 1] = '\0';
  return static_cast<int>(rep.size());
}

/// Normalisation factors (trv::calc_bispec_normalisation_from_particles /
/// _from_mesh).
int trv_norm(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment, double* norm
) {
  return guarded([&]() {
    trv::ParticleCatalogue cat(60);
    cat.load_particle_arrays(n, x, y, z, nz, ws, wc);
    if (from_mesh) {
      trv::ParameterSet params;
      set_params(params, "sim", "bispec", boxsize, ngrid, assignment, "false",
                 0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, 60, 0, 0, 1);
      *norm = trv::calc_bispec_normalisation_from_mesh(cat, params, alpha);
    } else {
      *norm = trv::calc_bispec_normalisation_from_particles(cat, alpha);
    }
  });
}

int trv_norm_powspec(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment, double* norm
) {
  return guarded([&]() {
    trv::ParticleCatalogue cat(60);
    cat.load_particle_arrays(n, x, y, z, nz, ws, wc);
    if (from_mesh) {
      trv::ParameterSet params;
      set_params(params, "sim", "powspec", boxsize, ngrid, assignment, "false",
                 0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, 60, 0, 0, 1);
      *norm = trv::calc_powspec_normalisation_from_mesh(cat, params, alpha);
    } else {
      *norm = trv::calc_powspec_normalisation_from_particles(cat, alpha);
    }
  });
}

/// trv::MeshField pipeline for intermediate checks.  `stage`: 0 assignment,
/// 1 + fourier_transform, 2 + apply_assignment_compensation,
/// 3 + inv_fourier_transform.  Weights: complex per particle or null (unit).
/// `field_out`: 2 * nmesh doubles.
int trv_mesh(
  int stage, int subtract_mean, int interlace, int deterministic,
  int n, const double* x, const double* y, const double* z,
  const double* w_re, const double* w_im,
  const double* boxsize, const int* ngrid, const char* assignment,
  double* field_out, double* elapsed_assign_s
) {
  return guarded([&]() {
    trv::ParameterSet params;
    set_params(params, "sim", "bispec", boxsize, ngrid, assignment, "false",
               0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, 60, deterministic, 0, 1);
    if (interlace) params.interlace = "true";   // after validate(): SURVEY.md F2
    trv::ParticleCatalogue cat(60);
    cat.load_particle_arrays(n, x, y, z, nullptr, nullptr, nullptr);
    trv::MeshField mesh(params, true, "`capi_mesh`");
    std::vector<double> weights(2 * (size_t)n);
    for (int i = 0; i < n; i++) {
      weights[2 * i] = w_re ? w_re[i] : 1.;
      weights[2 * i + 1] = w_im ? w_im[i] : 0.;
    }
    auto t0 = std::chrono::steady_clock::now();
    mesh.assign_weighted_field_to_mesh(cat, reinterpret_cast<double (*)[2]>(weights.data()));
    trvb_ctx_sync(mesh.context());
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_assign_s) *elapsed_assign_s = std::chrono::duration<double>(t1 - t0).count();
    if (subtract_mean) {
      const double nbar = double(cat.ntotal) / mesh.vol;
      trv::dev::check(trvb_mesh_add_const(mesh.context(), mesh.device_view(), -nbar),
                      "trvb_mesh_add_const");
    }
    if (stage == 4) {
      // assignment followed by MeshField::apply_wide_angle_pow_law_kernel at order (1, 2)
      mesh.params.i_wa = 1; mesh.params.j_wa = 2;
      mesh.apply_wide_angle_pow_law_kernel();
    } else {
      if (stage >= 1) mesh.fourier_transform();
      if (stage >= 2) mesh.apply_assignment_compensation();
      if (stage >= 3) mesh.inv_fourier_transform();
    }
    mesh.sync_host();
    std::memcpy(field_out, mesh.field, 2 * sizeof(double) * (size_t)params.nmesh);
  });
}

void trv_ylm(int ell, int m, const double* pos, int n, double* out) {
  for (int i = 0; i < n; i++) {
    double p[3] = {pos[3*i], pos[3*i+1], pos[3*i+2]};
    std::complex<double> y = trv::maths::SphericalHarmonicCalculator::
      calc_reduced_spherical_harmonic(ell, m, p);
    out[2*i] = y.real(); out[2*i+1] = y.imag();
  }
}

int trv_ylm_mesh(int fourier, int ell, int m, const double* boxsize, const int* ngrid,
                 double* out) {
  return guarded([&]() {
    const long long nmesh = (long long)ngrid[0] * ngrid[1] * ngrid[2];
    std::vector< std::complex<double> > tab(nmesh);
    typedef trv::maths::SphericalHarmonicCalculator SHC;
    if (fourier) SHC::store_reduced_spherical_harmonic_in_fourier_space(ell, m, boxsize, ngrid, tab);
    else SHC::store_reduced_spherical_harmonic_in_config_space(ell, m, boxsize, ngrid, tab);
    std::memcpy(out, tab.data(), sizeof(double) * 2 * nmesh);
  });
}

void trv_sjl(int ell, const double* x, int n, double* out) {
  trv::maths::SphericalBesselCalculator sj(ell);
  for (int i = 0; i < n; i++) out[i] = sj.eval(x[i]);
}

double trv_sjl_exact(int ell, double x) { return trv::maths::sph_bessel_jl(ell, x); }

double trv_w3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  return trv::maths::wigner_3j(j1, j2, j3, m1, m2, m3);
}

double trv_coupling(int l1, int l2, int L, int m1, int m2, int M) {
  return trv::calc_coupling_coeff_3pt(l1, l2, L, m1, m2, M);
}

int trv_binning(
  const char* space, const char* scheme, double bmin, double bmax, int nb,
  const double* boxsize, const int* ngrid,
  double* edges, double* centres, double* widths
) {
  return guarded([&]() {
    trv::ParameterSet params;
    for (int ax = 0; ax < 3; ax++) { params.boxsize[ax] = boxsize[ax]; params.ngrid[ax] = ngrid[ax]; }
    params.space = space; params.binning = scheme;
    params.bin_min = bmin; params.bin_max = bmax; params.num_bins = nb;
    trv::Binning b(params);
    b.set_bins();
    for (int i = 0; i <= nb; i++) edges[i] = b.bin_edges[i];
    for (int i = 0; i < nb; i++) { centres[i] = b.bin_centres[i]; widths[i] = b.bin_widths[i]; }
  });
}

int trv_validate(
  const char* catalogue_type, const char* statistic_type,
  const char* assignment, const char* interlace, const char* form,
  int ell1, int ell2, int ELL, int num_bins, int idx_bin,
  double bin_min, double bin_max,
  char* shape_out, char* interlace_out, char* npoint_out, char* space_out,
  int* assignment_order
) {
  return guarded([&]() {
    trv::ParameterSet params;
    const double boxsize[3] = {1000., 1000., 1000.};
    const int ngrid[3] = {64, 64, 64};
    set_params(params, catalogue_type, statistic_type, boxsize, ngrid, assignment,
               interlace, ell1, ell2, ELL, form, idx_bin, "lin", bin_min, bin_max,
               num_bins, 60, 0, 0, 1);
    std::strcpy(shape_out, params.shape.c_str());
    std::strcpy(interlace_out, params.interlace.c_str());
    std::strcpy(npoint_out, params.npoint.c_str());
    std::strcpy(space_out, params.space.c_str());
    *assignment_order = params.assignment_order;
  });
}

}  // extern "C"
