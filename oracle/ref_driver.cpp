// TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled from
// /root/reference/src/triumvirate/src by oracle/Makefile into
// oracle/_ref/libtrv_ref.so).  It lets the Python tests and bench.py's
// cpu_baseline / --impl reference legs call the reference's own
// implementation of the hot path:
//   trv::compute_bispec / compute_3pcf / compute_bispec_in_gpp_box /
//   compute_3pcf_in_gpp_box / compute_3pcf_window
//                                           (S/threept.cpp:248,1014,1473,2190,2621)
//   trv::MeshField assignment / FFT / compensation (S/field.cpp:569-1785)
//   trv::calc_bispec_normalisation_from_{particles,mesh} (S/threept.cpp:96,138)
//   trv::compute_powspec / compute_corrfunc / *_in_gpp_box / compute_corrfunc_window
//                                           (S/twopt.cpp:388,498,609,703,795)
//   trv::maths calculators                  (S/maths.cpp:167-375)
// Nothing here re-implements reference logic: it only marshals arrays.

#include <chrono>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "dataobjs.hpp"
#include "field.hpp"
#include "maths.hpp"
#include "monitor.hpp"
#include "parameters.hpp"
#include "particles.hpp"
#include "threept.hpp"
#include "twopt.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

void load_catalogue(
  trv::ParticleCatalogue& cat, int n,
  const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc
) {
  std::vector<double> vx(x, x + n), vy(y, y + n), vz(z, z + n);
  std::vector<double> vnz(n, 0.), vws(n, 1.), vwc(n, 1.);
  if (nz) vnz.assign(nz, nz + n);
  if (ws) vws.assign(ws, ws + n);
  if (wc) vwc.assign(wc, wc + n);
  cat.load_particle_data(vx, vy, vz, vnz, vws, vwc);
}

void fill_params(
  trv::ParameterSet& params,
  const char* catalogue_type, const char* statistic_type,
  const double boxsize[3], const int ngrid[3], const char* assignment,
  int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  int interlace_after_validate, int verbose
) {
  params.catalogue_type = catalogue_type;
  params.statistic_type = statistic_type;
  for (int i = 0; i < 3; i++) {
    params.boxsize[i] = boxsize[i];
    params.ngrid[i] = ngrid[i];
  }
  params.alignment = "centre";
  params.padscale = "box";
  params.assignment = assignment;
  params.interlace = "false";
  params.ell1 = ell1; params.ell2 = ell2; params.ELL = ELL;
  params.form = form;
  params.idx_bin = idx_bin;
  params.norm_convention = "particle";
  params.binning = binning;
  params.bin_min = bin_min; params.bin_max = bin_max;
  params.num_bins = num_bins;
  params.fftw_scheme = "estimate";
  params.use_fftw_wisdom = "false";
  params.verbose = verbose;
  params.progbar = "false";
  params.validate(false);
  // Interlacing is forced off by validate() for three-point statistics
  // (S/parameters.cpp:1240-1249); the interlaced MeshField paths are only
  // reachable by poking the public member afterwards.
  if (interlace_after_validate) params.interlace = "true";
}

}  // namespace

extern "C" {

const char* trvref_last_error() { return g_err.c_str(); }

int trvref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void trvref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// Three-point estimator.  `stat` = "bispec" | "3pcf"; `catalogue_type` =
// "sim" (periodic box, global plane-parallel) | "survey" (data + randoms,
// local plane-parallel; `los_*` are n x 3 row-major unit vectors).
// Output arrays must hold at least dv_dim entries (raw/shot: 2 * dv_dim).
int trvref_threept(
  const char* stat, const char* catalogue_type,
  int nd, const double* xd, const double* yd, const double* zd,
  const double* nzd, const double* wsd, const double* wcd, const double* los_d,
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment,
  int ell1, int ell2, int ELL, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  int interlace_after_validate, double norm_factor, int verbose,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot, double* elapsed_s
) {
  try {
    trv::ParameterSet params;
    fill_params(
      params, catalogue_type, stat, boxsize, ngrid, assignment,
      ell1, ell2, ELL, form, idx_bin, binning, bin_min, bin_max, num_bins,
      interlace_after_validate, verbose
    );
    trv::Binning bins(params);
    bins.set_bins();

    trv::ParticleCatalogue data(verbose), rand(verbose);
    load_catalogue(data, nd, xd, yd, zd, nzd, wsd, wcd);
    const bool survey = std::string(catalogue_type) == "survey";
    if (survey) load_catalogue(rand, nr, xr, yr, zr, nzr, wsr, wcr);

    const bool is_bispec = std::string(stat) == "bispec";
    auto t0 = std::chrono::steady_clock::now();
    if (is_bispec) {
      trv::BispecMeasurements out = survey
        ? trv::compute_bispec(
            data, rand, (trv::LineOfSight*)los_d, (trv::LineOfSight*)los_r,
            params, bins, norm_factor)
        : trv::compute_bispec_in_gpp_box(data, params, bins, norm_factor);
      auto t1 = std::chrono::steady_clock::now();
      if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
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
        ? trv::compute_3pcf(
            data, rand, (trv::LineOfSight*)los_d, (trv::LineOfSight*)los_r,
            params, bins, norm_factor)
        : trv::compute_3pcf_in_gpp_box(data, params, bins, norm_factor);
      auto t1 = std::chrono::steady_clock::now();
      if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
      *dim = out.dim;
      for (int i = 0; i < out.dim; i++) {
        c1_bin[i] = out.r1_bin[i]; c2_bin[i] = out.r2_bin[i];
        c1_eff[i] = out.r1_eff[i]; c2_eff[i] = out.r2_eff[i];
        n1[i] = out.npairs_1[i]; n2[i] = out.npairs_2[i];
        raw[2*i] = out.zeta_raw[i].real(); raw[2*i+1] = out.zeta_raw[i].imag();
        shot[2*i] = out.zeta_shot[i].real(); shot[2*i+1] = out.zeta_shot[i].imag();
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// 3PCF window function (trv::compute_3pcf_window, S/threept.cpp:2621-3077).
int trvref_threept_window(
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment,
  int ell1, int ell2, int ELL, int i_wa, int j_wa, const char* form, int idx_bin,
  const char* binning, double bin_min, double bin_max, int num_bins,
  double alpha, double norm_factor, int wide_angle, int verbose,
  int* dim, double* c1_bin, double* c2_bin, double* c1_eff, double* c2_eff,
  int* n1, int* n2, double* raw, double* shot
) {
  try {
    trv::ParameterSet params;
    params.i_wa = i_wa; params.j_wa = j_wa;
    fill_params(
      params, "random", wide_angle ? "3pcf-win-wa" : "3pcf-win", boxsize, ngrid, assignment,
      ell1, ell2, ELL, form, idx_bin, binning, bin_min, bin_max, num_bins, 0, verbose
    );
    trv::Binning bins(params);
    bins.set_bins();
    trv::ParticleCatalogue rand(verbose);
    load_catalogue(rand, nr, xr, yr, zr, nzr, wsr, wcr);
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
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ---------------------------------------------------------------------
// Bounded-sample timing of the reference's periodic-box bispectrum.
//
// A full `form = full` run at 512^3 takes of order an hour on a CPU, so the
// bench times the reference in two parts that together are its whole loop
// (S/threept.cpp:1543-1669 and :1675-2140):
//   setup   dn_00, N_L0, G_00, the four y_lm tables, Bessel calculators
//           -- done once per (m1, m2, M) term, independent of the bin pair;
//   pair    what the reference repeats for EVERY bin pair: two band-limited
//           inverse transforms, the triple-product sum and the per-bin
//           shot-noise transform + reduction (3 IFFTs + 5 mesh passes).
// time(full run) = setup + (number of pairs) x pair, which the caller forms.
// ---------------------------------------------------------------------

struct RefBispecState {
  trv::ParameterSet params;
  trv::ParticleCatalogue* cat = nullptr;
  trv::Binning* bins = nullptr;
  trv::MeshField* dn_00 = nullptr;
  trv::MeshField* N_L0 = nullptr;
  trv::MeshField* G_00 = nullptr;
  trv::MeshField* F_a = nullptr;
  trv::MeshField* F_b = nullptr;
  trv::FieldStats* stats = nullptr;
  trv::maths::SphericalBesselCalculator* sj_a = nullptr;
  trv::maths::SphericalBesselCalculator* sj_b = nullptr;
  std::vector< std::complex<double> > ylm_k_a, ylm_k_b, ylm_r_a, ylm_r_b;
  double last_keff[2] = {0., 0.};
  int last_nmodes[2] = {0, 0};
};
static RefBispecState* g_state = nullptr;

void trvref_bispec_teardown() {
  if (!g_state) return;
  delete g_state->F_a; delete g_state->F_b; delete g_state->G_00;
  delete g_state->N_L0; delete g_state->dn_00; delete g_state->stats;
  delete g_state->sj_a; delete g_state->sj_b; delete g_state->bins; delete g_state->cat;
  delete g_state; g_state = nullptr;
}

int trvref_bispec_setup(
  int n, const double* x, const double* y, const double* z,
  const double* boxsize, const int* ngrid, const char* assignment,
  double bin_min, double bin_max, int num_bins, double* elapsed_s
) {
  try {
    trvref_bispec_teardown();
    g_state = new RefBispecState();
    RefBispecState& s = *g_state;
    fill_params(s.params, "sim", "bispec", boxsize, ngrid, assignment, 0, 0, 0, "full", 0,
                "lin", bin_min, bin_max, num_bins, 0, 60);
    s.bins = new trv::Binning(s.params);
    s.bins->set_bins();
    s.cat = new trv::ParticleCatalogue(60);
    load_catalogue(*s.cat, n, x, y, z, nullptr, nullptr, nullptr);
    auto t0 = std::chrono::steady_clock::now();
    // S/threept.cpp:1543-1569.
    s.dn_00 = new trv::MeshField(s.params, true, "`dn_00`");
    s.dn_00->compute_unweighted_field_fluctuations_insitu(*s.cat);
    s.dn_00->fourier_transform();
    s.N_L0 = new trv::MeshField(s.params, true, "`N_L0`");
    s.N_L0->compute_unweighted_field(*s.cat);
    s.N_L0->fourier_transform();
    s.sj_a = new trv::maths::SphericalBesselCalculator(0);
    s.sj_b = new trv::maths::SphericalBesselCalculator(0);
    s.stats = new trv::FieldStats(s.params);
    s.ylm_k_a.resize(s.params.nmesh); s.ylm_k_b.resize(s.params.nmesh);
    s.ylm_r_a.resize(s.params.nmesh); s.ylm_r_b.resize(s.params.nmesh);
    // S/threept.cpp:1618-1633.
    typedef trv::maths::SphericalHarmonicCalculator SHC;
    SHC::store_reduced_spherical_harmonic_in_fourier_space(0, 0, s.params.boxsize, s.params.ngrid, s.ylm_k_a);
    SHC::store_reduced_spherical_harmonic_in_fourier_space(0, 0, s.params.boxsize, s.params.ngrid, s.ylm_k_b);
    SHC::store_reduced_spherical_harmonic_in_config_space(0, 0, s.params.boxsize, s.params.ngrid, s.ylm_r_a);
    SHC::store_reduced_spherical_harmonic_in_config_space(0, 0, s.params.boxsize, s.params.ngrid, s.ylm_r_b);
    // S/threept.cpp:1665-1672.
    s.G_00 = new trv::MeshField(s.params, true, "`G_00`");
    s.G_00->compute_unweighted_field_fluctuations_insitu(*s.cat);
    s.G_00->fourier_transform();
    s.G_00->apply_assignment_compensation();
    s.G_00->inv_fourier_transform();
    s.F_a = new trv::MeshField(s.params, true, "`F_lm_a`");
    s.F_b = new trv::MeshField(s.params, true, "`F_lm_b`");
    auto t1 = std::chrono::steady_clock::now();
    *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// One bin pair (idx_row, idx_col) exactly as the reference's `triu` branch
// does it (S/threept.cpp:1900-1968) plus its per-pair shot-noise term
// (S/threept.cpp:2126-2140).  out = {Re bk, Im bk, Re S, Im S}.
int trvref_bispec_pair(int idx_row, int idx_col, double* out, double* elapsed_s) {
  try {
    if (!g_state) { g_err = "trvref_bispec_setup not called"; return 1; }
    RefBispecState& s = *g_state;
    auto t0 = std::chrono::steady_clock::now();
    double k_eff_a, k_eff_b; int nmodes_a, nmodes_b;
    s.F_a->inv_fourier_transform_ylm_wgtd_field_band_limited(
      *s.dn_00, s.ylm_k_a, s.bins->bin_edges[idx_row], s.bins->bin_edges[idx_row + 1],
      k_eff_a, nmodes_a);
    s.F_b->inv_fourier_transform_ylm_wgtd_field_band_limited(
      *s.dn_00, s.ylm_k_b, s.bins->bin_edges[idx_col], s.bins->bin_edges[idx_col + 1],
      k_eff_b, nmodes_b);
    double bk_re = 0., bk_im = 0.;
    trv::MeshField& Fa = *s.F_a; trv::MeshField& Fb = *s.F_b; trv::MeshField& G = *s.G_00;
#pragma omp parallel for reduction(+:bk_re, bk_im)
    for (long long gid = 0; gid < s.params.nmesh; gid++) {
      std::complex<double> fa(Fa[gid][0], Fa[gid][1]);
      std::complex<double> fb(Fb[gid][0], Fb[gid][1]);
      std::complex<double> g(G[gid][0], G[gid][1]);
      std::complex<double> v = fa * fb * g;
      bk_re += v.real(); bk_im += v.imag();
    }
    std::complex<double> S = s.stats->compute_uncoupled_shotnoise_for_bispec_per_bin(
      *s.dn_00, *s.N_L0, s.ylm_r_a, s.ylm_r_b, *s.sj_a, *s.sj_b,
      double(s.cat->ntotal), k_eff_a, k_eff_b);
    auto t1 = std::chrono::steady_clock::now();
    *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    out[0] = bk_re * s.dn_00->vol_cell; out[1] = bk_im * s.dn_00->vol_cell;
    out[2] = S.real(); out[3] = S.imag();
    s.last_keff[0] = k_eff_a; s.last_keff[1] = k_eff_b;
    s.last_nmodes[0] = nmodes_a; s.last_nmodes[1] = nmodes_b;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Shell statistics (k_eff, nmodes) of the two bins of the last trvref_bispec_pair call
// (what S/threept.cpp:1912-1919 stores in k1eff_dv/k2eff_dv/nmodes1_dv/nmodes2_dv).
int trvref_bispec_pair_shells(double* keff, int* nmodes) {
  if (!g_state) { g_err = "trvref_bispec_setup not called"; return 1; }
  keff[0] = g_state->last_keff[0]; keff[1] = g_state->last_keff[1];
  nmodes[0] = g_state->last_nmodes[0]; nmodes[1] = g_state->last_nmodes[1];
  return 0;
}

// The binned two-point statistics behind the S|{i != j = k} and S|{j != i = k}
// shot-noise terms (S/threept.cpp:1988-1992, 2057-2061):
// FieldStats::compute_ylm_wgtd_2pt_stats_in_fourier(dn_00, N_L0, Sbar, 0, 0, kbinning).
// pk_sn holds num_bins x {Re pk, Im pk, Re sn, Im sn}.
int trvref_bispec_twopt(double* pk_sn) {
  try {
    if (!g_state) { g_err = "trvref_bispec_setup not called"; return 1; }
    RefBispecState& s = *g_state;
    std::complex<double> Sbar = double(s.cat->ntotal);
    s.stats->compute_ylm_wgtd_2pt_stats_in_fourier(*s.dn_00, *s.N_L0, Sbar, 0, 0, *s.bins);
    for (int i = 0; i < s.bins->num_bins; i++) {
      pk_sn[4*i] = s.stats->pk[i].real(); pk_sn[4*i + 1] = s.stats->pk[i].imag();
      pk_sn[4*i + 2] = s.stats->sn[i].real(); pk_sn[4*i + 3] = s.stats->sn[i].imag();
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Calibration of the FFTW stand-in (oracle/shim/trvshim_fft.cpp): best-of-`reps`
// wall time of one in-place forward complex transform of an n^3 array through the
// same calls the reference makes (fftw_plan_dft_3d + fftw_execute,
// S/field.cpp:246,1552), to be quoted beside scipy.fft.fftn(workers=cores) on the
// same host.
int trvref_fft_time(int n, int reps, double* seconds) {
  try {
    size_t nmesh = size_t(n) * n * n;
    fftw_complex* a = fftw_alloc_complex(nmesh);
    if (!a) { g_err = "fftw_alloc_complex failed"; return 1; }
#if defined(TRV_USE_OMP) && defined(TRV_USE_FFTWOMP)
    fftw_init_threads();
    fftw_plan_with_nthreads(omp_get_max_threads());
#endif
    fftw_plan plan = fftw_plan_dft_3d(n, n, n, a, a, FFTW_FORWARD, FFTW_ESTIMATE);
    double best = 1.e300;
    for (int r = 0; r < reps; r++) {
#pragma omp parallel for
      for (long long i = 0; i < (long long)nmesh; i++) {
        a[i][0] = double((i * 2654435761ULL) % 1024) / 1024. - 0.5; a[i][1] = 0.;
      }
      auto t0 = std::chrono::steady_clock::now();
      fftw_execute(plan);
      auto t1 = std::chrono::steady_clock::now();
      double t = std::chrono::duration<double>(t1 - t0).count();
      if (t < best) best = t;
    }
    fftw_destroy_plan(plan);
    fftw_free(a);
    *seconds = best;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Normalisation factors (S/threept.cpp:96-149).
int trvref_norm(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment,
  double* norm
) {
  try {
    trv::ParticleCatalogue cat(60);
    load_catalogue(cat, n, x, y, z, nz, ws, wc);
    if (from_mesh) {
      trv::ParameterSet params;
      fill_params(params, "sim", "bispec", boxsize, ngrid, assignment,
                  0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, 0, 60);
      *norm = trv::calc_bispec_normalisation_from_mesh(cat, params, alpha);
    } else {
      *norm = trv::calc_bispec_normalisation_from_particles(cat, alpha);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Two-point estimators (S/twopt.cpp:388-901).  `stat` = "powspec" | "2pcf" |
// "2pcf-win"; interlacing is honoured by validate() for these statistics.
int trvref_twopt(
  const char* stat, const char* catalogue_type,
  int nd, const double* xd, const double* yd, const double* zd,
  const double* nzd, const double* wsd, const double* wcd, const double* los_d,
  int nr, const double* xr, const double* yr, const double* zr,
  const double* nzr, const double* wsr, const double* wcr, const double* los_r,
  const double* boxsize, const int* ngrid, const char* assignment, int interlace,
  int ELL, const char* binning, double bin_min, double bin_max, int num_bins,
  double alpha, double norm_factor, int verbose,
  int* dim, double* c_bin, double* c_eff, int* count, double* raw, double* shot,
  double* elapsed_s
) {
  try {
    const std::string st(stat);
    trv::ParameterSet params;
    params.catalogue_type = catalogue_type;
    params.statistic_type = stat;
    for (int i = 0; i < 3; i++) { params.boxsize[i] = boxsize[i]; params.ngrid[i] = ngrid[i]; }
    params.alignment = "centre";
    params.padscale = "box";
    params.assignment = assignment;
    params.interlace = interlace ? "true" : "false";
    params.ell1 = ELL; params.ell2 = 0; params.ELL = ELL;
    params.form = "diag";
    params.idx_bin = 0;
    params.norm_convention = "particle";
    params.binning = binning;
    params.bin_min = bin_min; params.bin_max = bin_max;
    params.num_bins = num_bins;
    params.fftw_scheme = "estimate";
    params.use_fftw_wisdom = "false";
    params.verbose = verbose;
    params.progbar = "false";
    params.validate(false);
    trv::Binning bins(params);
    bins.set_bins();

    const bool survey = std::string(catalogue_type) == "survey";
    const bool window = st == "2pcf-win";
    trv::ParticleCatalogue data(verbose), rand(verbose);
    if (!window) load_catalogue(data, nd, xd, yd, zd, nzd, wsd, wcd);
    if (survey || window) load_catalogue(rand, nr, xr, yr, zr, nzr, wsr, wcr);
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
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

int trvref_norm_powspec(
  int from_mesh, int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc, double alpha,
  const double* boxsize, const int* ngrid, const char* assignment,
  double* norm
) {
  try {
    trv::ParticleCatalogue cat(60);
    load_catalogue(cat, n, x, y, z, nz, ws, wc);
    if (from_mesh) {
      trv::ParameterSet params;
      fill_params(params, "sim", "powspec", boxsize, ngrid, assignment,
                  0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, 0, 60);
      *norm = trv::calc_powspec_normalisation_from_mesh(cat, params, alpha);
    } else {
      *norm = trv::calc_powspec_normalisation_from_particles(cat, alpha);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Mesh pipeline for intermediate parity checks.  `stage`:
//   0  assignment only (complex weights w_re/w_im; NULL -> unit weights)
//   1  + fourier_transform()                       (S/field.cpp:1496)
//   2  + apply_assignment_compensation()           (S/field.cpp:1764)
//   3  + inv_fourier_transform()                   (S/field.cpp:1657)
// `subtract_mean` applies compute_unweighted_field_fluctuations_insitu's
// `field.re -= N/V` after assignment (S/field.cpp:1235-1243).
// `field_out` receives 2*nmesh doubles (re, im interleaved).
int trvref_mesh(
  int stage, int subtract_mean, int interlace,
  int n, const double* x, const double* y, const double* z,
  const double* w_re, const double* w_im,
  const double* boxsize, const int* ngrid, const char* assignment,
  double* field_out, double* elapsed_assign_s
) {
  try {
    trv::ParameterSet params;
    fill_params(params, "sim", "bispec", boxsize, ngrid, assignment,
                0, 0, 0, "diag", 0, "lin", 0.005, 0.105, 4, interlace, 60);
    trv::ParticleCatalogue cat(60);
    load_catalogue(cat, n, x, y, z, nullptr, nullptr, nullptr);
    trv::MeshField mesh(params, true, "`oracle_mesh`");
    fftw_complex* weights = fftw_alloc_complex(n);
    for (int i = 0; i < n; i++) {
      weights[i][0] = w_re ? w_re[i] : 1.;
      weights[i][1] = w_im ? w_im[i] : 0.;
    }
    auto t0 = std::chrono::steady_clock::now();
    mesh.assign_weighted_field_to_mesh(cat, weights);
    auto t1 = std::chrono::steady_clock::now();
    if (elapsed_assign_s)
      *elapsed_assign_s = std::chrono::duration<double>(t1 - t0).count();
    fftw_free(weights);
    if (subtract_mean) {
      double nbar = double(cat.ntotal) / mesh.vol;
      for (long long g = 0; g < params.nmesh; g++) mesh.field[g][0] -= nbar;
    }
    if (stage == 4) {
      // assignment followed by MeshField::apply_wide_angle_pow_law_kernel at order (1, 2)
      mesh.params.i_wa = 1; mesh.params.j_wa = 2;
      mesh.apply_wide_angle_pow_law_kernel();
      std::memcpy(field_out, mesh.field, sizeof(fftw_complex) * params.nmesh);
      return 0;
    }
    if (stage >= 1) mesh.fourier_transform();
    if (stage >= 2) mesh.apply_assignment_compensation();
    if (stage >= 3) mesh.inv_fourier_transform();
    std::memcpy(field_out, mesh.field, sizeof(fftw_complex) * params.nmesh);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Calculators (S/maths.cpp).
void trvref_ylm(int ell, int m, const double* pos, int n, double* out) {
  for (int i = 0; i < n; i++) {
    double p[3] = {pos[3*i], pos[3*i+1], pos[3*i+2]};
    std::complex<double> y = trv::maths::SphericalHarmonicCalculator::
      calc_reduced_spherical_harmonic(ell, m, p);
    out[2*i] = y.real(); out[2*i+1] = y.imag();
  }
}

// SphericalHarmonicCalculator::store_reduced_spherical_harmonic_in_{fourier,config}_space
// (S/maths.cpp:222-302); out holds 2 * nmesh doubles.
int trvref_ylm_mesh(int fourier, int ell, int m, const double* boxsize, const int* ngrid,
                    double* out) {
  try {
    const long long nmesh = (long long)ngrid[0] * ngrid[1] * ngrid[2];
    std::vector< std::complex<double> > tab(nmesh);
    typedef trv::maths::SphericalHarmonicCalculator SHC;
    if (fourier) SHC::store_reduced_spherical_harmonic_in_fourier_space(ell, m, boxsize, ngrid, tab);
    else SHC::store_reduced_spherical_harmonic_in_config_space(ell, m, boxsize, ngrid, tab);
    std::memcpy(out, tab.data(), sizeof(double) * 2 * nmesh);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

void trvref_sjl(int ell, const double* x, int n, double* out) {
  trv::maths::SphericalBesselCalculator sj(ell);
  for (int i = 0; i < n; i++) out[i] = sj.eval(x[i]);
}

double trvref_w3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  return trv::maths::wigner_3j(j1, j2, j3, m1, m2, m3);
}

double trvref_coupling(int l1, int l2, int L, int m1, int m2, int M) {
  return trv::calc_coupling_coeff_3pt(l1, l2, L, m1, m2, M);
}

// Binning (S/dataobjs.cpp:134-249): edges[nb+1], centres[nb], widths[nb].
int trvref_binning(
  const char* space, const char* scheme, double bmin, double bmax, int nb,
  const double* boxsize, const int* ngrid,
  double* edges, double* centres, double* widths
) {
  try {
    trv::ParameterSet params;
    for (int i = 0; i < 3; i++) {
      params.boxsize[i] = boxsize[i]; params.ngrid[i] = ngrid[i];
    }
    params.space = space; params.binning = scheme;
    params.bin_min = bmin; params.bin_max = bmax; params.num_bins = nb;
    trv::Binning b(params);
    b.set_bins();
    for (int i = 0; i <= nb; i++) edges[i] = b.bin_edges[i];
    for (int i = 0; i < nb; i++) {
      centres[i] = b.bin_centres[i]; widths[i] = b.bin_widths[i];
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ParameterSet::validate derivations (S/parameters.cpp:466-1270).
int trvref_validate(
  const char* catalogue_type, const char* statistic_type,
  const char* assignment, const char* interlace, const char* form,
  int ell1, int ell2, int ELL, int num_bins, int idx_bin,
  double bin_min, double bin_max,
  char* shape_out, char* interlace_out, char* npoint_out, char* space_out,
  int* assignment_order
) {
  try {
    trv::ParameterSet params;
    params.catalogue_type = catalogue_type;
    params.statistic_type = statistic_type;
    params.boxsize[0] = params.boxsize[1] = params.boxsize[2] = 1000.;
    params.ngrid[0] = params.ngrid[1] = params.ngrid[2] = 64;
    params.assignment = assignment;
    params.interlace = interlace;
    params.form = form;
    params.ell1 = ell1; params.ell2 = ell2; params.ELL = ELL;
    params.num_bins = num_bins; params.idx_bin = idx_bin;
    params.bin_min = bin_min; params.bin_max = bin_max;
    params.verbose = 60;
    params.fftw_scheme = "estimate";
    params.validate(false);
    std::strcpy(shape_out, params.shape.c_str());
    std::strcpy(interlace_out, params.interlace.c_str());
    std::strcpy(npoint_out, params.npoint.c_str());
    std::strcpy(space_out, params.space.c_str());
    *assignment_order = params.assignment_order;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
