// twopt.cpp -- two-point estimators on the device (drop-in for S/twopt.cpp).
//
// The reference builds every delta n_LM with a fresh MeshField (catalogue
// scatter on the host, mesh copied to the GPU and back around each cuFFT call,
// S/field.cpp:1525-1536) and bins on the host.  Here the catalogues are uploaded
// once per call, meshes never leave the device, real fields (M = 0, no
// interlacing) use real-to-complex transforms and the binned statistics visit
// only the cells that can fall in a bin.  What is reproduced literally: the
// (m1, M) term loop and couplings (S/twopt.cpp:434-466), the fine-bin sampling
// rules (dk = 1e-5 / dr = 0.1, S/field.cpp:2571-2572, 2846-2847), the window /
// aliasing branches of interlaced and plain meshes (S/field.cpp:2543-2569,
// 3504-3527) and the normalisation conventions.

#include "trv/twopt.hpp"

#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "trv/maths.hpp"
#include "trv/monitor.hpp"
#include "trvb.h"

namespace trvs = trv::sys;
namespace trvm = trv::maths;

namespace trv {

namespace {

typedef std::complex<double> cdouble;

/// CAVEAT of the reference (S/twopt.cpp:32-33): discretionary tolerance.
const double eps_norm = 1.e-5;

cdouble cat_sum(trvb_ctx* ctx, dev::Catalogue& cat, int kind, int ell, int m) {
  double out[2];
  dev::check(trvb_cat_sum(ctx, cat.get(), kind, ell, m, out), "trvb_cat_sum");
  return cdouble(out[0], out[1]);
}

trv::ParameterSet small_context_params() {
  // A context is needed only for its device/stream; any small grid does.
  trv::ParameterSet p;
  for (int ax = 0; ax < 3; ax++) { p.boxsize[ax] = 1.; p.ngrid[ax] = 4; }
  p.assignment_order = 1;
  return p;
}

void require_particles(ParticleCatalogue& particles) {
  if (particles.pdata == nullptr) {
    if (trvs::currTask == 0) trvs::logger.error("Particle data are uninitialised.");
    throw trvs::InvalidDataError("Particle data are uninitialised.");
  }
}

/// Device state of one two-point estimator call.
class TwoPtEngine {
 public:
  /// Paired survey-type catalogues (`rand` non-null) or a periodic box.
  TwoPtEngine(trv::ParameterSet& params, ParticleCatalogue& data, ParticleCatalogue* rand,
              LineOfSight* los_data, LineOfSight* los_rand)
    : params_(params), survey_(rand != nullptr) {
    init();
    data_.reset(new dev::Catalogue(ctx_, data, los_data, survey_));
    if (survey_) rand_.reset(new dev::Catalogue(ctx_, *rand, los_rand, true));
    ndata_ = data.ntotal;
    if (survey_) alpha_ = data.wstotal / rand->wstotal;   // S/twopt.cpp:407
  }

  /// Window mode: one random catalogue with a given alpha (S/twopt.cpp:795-901).
  TwoPtEngine(trv::ParameterSet& params, ParticleCatalogue& rand, LineOfSight* los_rand,
              double alpha)
    : params_(params), survey_(true), window_(true) {
    init();
    data_.reset(new dev::Catalogue(ctx_, rand, los_rand, true));
    ndata_ = rand.ntotal;
    alpha_ = alpha;
  }

  double alpha() const { return alpha_; }

  /// delta n_LM(k) including the dV of the forward transform
  /// (S/field.cpp:1229-1362, 1496-1655); interlaced when the parameter set asks.
  dev::Mesh density_fluctuation(int L, int M) {
    // The interlaced combination is not Hermitian on the Nyquist planes (the
    // phase of m = -1/2 is its own mirror), so interlaced meshes stay complex.
    const bool real_field = (M == 0) && !interlaced_;
    dev::Mesh k = transformed(L, M, /*shifted=*/0, real_field);
    if (interlaced_) {
      dev::Mesh ks = transformed(L, M, /*shifted=*/1, real_field);
      dev::check(trvb_interlace_combine(c_, k.view(), ks.view()), "trvb_interlace_combine");
    }
    if (!survey_) {
      // Mean subtraction touches the k = 0 mode only: FFT[nbar dV] = N delta_k0.
      // The reference subtracts nbar from the primary mesh but NOT from the shadow
      // mesh (S/field.cpp:1228-1244 touches `field` only), so the interlaced
      // combination keeps (N - N + N) / 2 = N/2 at k = 0; reproduced for parity
      // (it offsets the interlaced box 2PCF by a constant).
      const double sub = interlaced_ ? 0.5 * double(ndata_) : double(ndata_);
      dev::check(trvb_kmesh_add_zero_mode(c_, k.view(), -sub), "trvb_kmesh_add_zero_mode");
    }
    return k;
  }

  /// \bar{N}_LM (S/twopt.cpp:298-380); box: N (S/twopt.cpp:659).
  cdouble shotnoise_amp(int L, int M) {
    if (!survey_) return cdouble(double(ndata_), 0.);
    // sum y_LM w^2 = conj(sum conj(y_LM) w^2): w is real.
    if (window_) return std::pow(alpha_, 2) * std::conj(cat_sum(c_, *data_, TRVB_W_CYLM_W2, L, M));
    return std::conj(cat_sum(c_, *data_, TRVB_W_CYLM_W2, L, M))
      + std::pow(alpha_, 2) * std::conj(cat_sum(c_, *rand_, TRVB_W_CYLM_W2, L, M));
  }

  /// FieldStats::compute_ylm_wgtd_2pt_stats_in_fourier (S/field.cpp:2511-2703).
  void stats_fourier(const dev::Mesh& fa, const dev::Mesh& fb, cdouble S, int ell, int m,
                     trv::Binning& kbinning, std::vector<int>& nmodes, std::vector<double>& k,
                     std::vector<cdouble>& pk, std::vector<cdouble>& sn) {
    const int nb = kbinning.num_bins;
    std::vector<long long> nm(nb);
    std::vector<double> pk2(2 * nb), sn2(2 * nb);
    k.assign(nb, 0.);
    const double Sv[2] = {S.real(), S.imag()};
    dev::check(trvb_twopt_fourier(c_, fa.view(), fb.view(), Sv, ell, m, interlaced_,
                                  kbinning.bin_edges.data(), kbinning.bin_centres.data(), nb,
                                  nm.data(), k.data(), pk2.data(), sn2.data()),
               "trvb_twopt_fourier");
    nmodes.resize(nb); pk.resize(nb); sn.resize(nb);
    for (int b = 0; b < nb; b++) {
      nmodes[b] = static_cast<int>(nm[b]);
      pk[b] = cdouble(pk2[2 * b], pk2[2 * b + 1]);
      sn[b] = cdouble(sn2[2 * b], sn2[2 * b + 1]);
    }
  }

  /// FieldStats::compute_ylm_wgtd_2pt_stats_in_config (S/field.cpp:2705-2945).
  void stats_config(const dev::Mesh& fa, const dev::Mesh& fb, cdouble S, int ell, int m,
                    trv::Binning& rbinning, std::vector<int>& npairs, std::vector<double>& r,
                    std::vector<cdouble>& xi) {
    const int nb = rbinning.num_bins;
    // Two half spectra and a real amplitude give a Hermitian product: real xi.
    const bool real_xi = fa.layout() == TRVB_HALF && fb.layout() == TRVB_HALF && S.imag() == 0.;
    dev::Mesh xi3d(ctx_, c_, real_xi ? TRVB_REAL : TRVB_COMPLEX);
    const double Sv[2] = {S.real(), S.imag()};
    dev::check(trvb_shot_xi(c_, fa.view(), fb.view(), Sv, interlaced_, xi3d.view()),
               "trvb_shot_xi");
    trvs::count_ifft += 1;
    std::vector<long long> np(nb);
    std::vector<double> xi2(2 * nb);
    r.assign(nb, 0.);
    dev::check(trvb_twopt_config_bin(c_, xi3d.view(), ell, m, rbinning.bin_edges.data(),
                                     rbinning.bin_centres.data(), nb, np.data(), r.data(),
                                     xi2.data()), "trvb_twopt_config_bin");
    npairs.resize(nb); xi.resize(nb);
    for (int b = 0; b < nb; b++) {
      npairs[b] = static_cast<int>(np[b]);
      xi[b] = cdouble(xi2[2 * b], xi2[2 * b + 1]);
    }
  }

 private:
  void init() {
    ctx_ = dev::acquire_context(params_);
    c_ = ctx_.get();
    interlaced_ = params_.interlace == "true" ? 1 : 0;
    mode_ = params_.deterministic ? 1 : 0;
    dev::check(trvb_ctx_set_deterministic(c_, mode_), "trvb_ctx_set_deterministic");
  }

  void assign(dev::Catalogue& cat, int kind, int L, int M, double scale, bool accumulate,
              int shifted, dev::Mesh& mesh) {
    dev::check(trvb_assign(c_, cat.get(), kind, L, M, scale, /*density_units=*/0,
                           accumulate ? 1 : 0, shifted, mode_, mesh.view()), "trvb_assign");
  }

  /// Assignment WITHOUT the 1/dV density factor followed by the forward transform
  /// WITHOUT its dV prescale: the two cancel (S/field.cpp:996, 1503-1510).
  dev::Mesh transformed(int L, int M, int shifted, bool real_field) {
    dev::Mesh x(ctx_, c_, real_field ? TRVB_REAL : TRVB_COMPLEX);
    if (!survey_) {
      assign(*data_, TRVB_W_UNIT, 0, 0, 1., false, shifted, x);
    } else if (window_) {
      // The reference applies alpha to the primary mesh only (S/field.cpp:1352-1360
      // scales `field`, never `field_s`): the shadow mesh of an interlaced window
      // measurement keeps unit scale.  Reproduced for parity; the Python front end
      // always passes alpha = 1 (T/twopt.py:1538), where it makes no difference.
      assign(*data_, TRVB_W_YLM_W, L, M, shifted ? 1. : alpha_, false, shifted, x);
    } else {
      assign(*data_, TRVB_W_YLM_W, L, M, 1., false, shifted, x);
      assign(*rand_, TRVB_W_YLM_W, L, M, -alpha_, true, shifted, x);
    }
    if (real_field) {
      dev::Mesh k(ctx_, c_, TRVB_HALF);
      dev::check(trvb_fft_forward(c_, x.view(), k.view(), 1.), "trvb_fft_forward");
      trvs::count_fft += 1;
      return k;
    }
    dev::check(trvb_fft_forward(c_, x.view(), x.view(), 1.), "trvb_fft_forward");
    trvs::count_fft += 1;
    return x;
  }

  trv::ParameterSet& params_;
  bool survey_;
  bool window_ = false;
  std::shared_ptr<trvb_ctx> ctx_;
  trvb_ctx* c_ = nullptr;
  std::unique_ptr<dev::Catalogue> data_, rand_;
  long long ndata_ = 0;
  double alpha_ = 1.;
  int interlaced_ = 0;
  int mode_ = 0;
};

/// Which space a multipole estimator bins in.
enum class Space { fourier, config };

struct TwoPtSums {
  std::vector<int> count;
  std::vector<double> coord;
  std::vector<cdouble> stat, shot;
};

/// The (M, m1) loop shared by compute_powspec, compute_corrfunc and
/// compute_corrfunc_window (S/twopt.cpp:427-470, 535-571, 834-866).
TwoPtSums multipole_terms(TwoPtEngine& eng, trv::ParameterSet& params, trv::Binning& binning,
                          Space space, const char* what) {
  const int nb = binning.num_bins;
  const int ell1 = params.ELL;
  TwoPtSums acc;
  acc.count.assign(nb, 0); acc.coord.assign(nb, 0.);
  acc.stat.assign(nb, cdouble(0., 0.)); acc.shot.assign(nb, cdouble(0., 0.));

  dev::Mesh dn_00 = eng.density_fluctuation(0, 0);   // delta n_00(k)
  for (int M_ = -params.ELL; M_ <= params.ELL; M_++) {
    dev::Mesh dn_LM_own;
    const bool reuse = (params.ELL == 0);   // delta n_LM is delta n_00 itself
    if (!reuse) dn_LM_own = eng.density_fluctuation(params.ELL, M_);
    const dev::Mesh& dn_LM = reuse ? dn_00 : dn_LM_own;
    const cdouble sn_amp = eng.shotnoise_amp(params.ELL, M_);   // \bar{N}_LM

    for (int m1 = -ell1; m1 <= ell1; m1++) {
      const double coupling = calc_coupling_coeff_2pt(ell1, params.ELL, m1, M_);
      if (std::fabs(coupling) < trvm::eps_coupling) continue;
      std::vector<int> count; std::vector<double> coord; std::vector<cdouble> stat, shot;
      if (space == Space::fourier) {
        eng.stats_fourier(dn_LM, dn_00, sn_amp, ell1, m1, binning, count, coord, stat, shot);
      } else {
        eng.stats_config(dn_LM, dn_00, sn_amp, ell1, m1, binning, count, coord, stat);
        shot.assign(nb, cdouble(0., 0.));
      }
      for (int b = 0; b < nb; b++) {
        acc.stat[b] += coupling * stat[b];
        acc.shot[b] += coupling * shot[b];
      }
      if (M_ == 0 && m1 == 0) { acc.count = count; acc.coord = coord; }
    }
    if (trvs::currTask == 0) trvs::logger.stat("%s term computed at order M = %d.", what, M_);
  }
  return acc;
}

void warn_box_norm(ParticleCatalogue& catalogue, trv::ParameterSet& params, double norm_factor,
                   const char* what) {
  const double norm = double(catalogue.ntotal) * double(catalogue.ntotal) / params.volume;
  if (std::fabs(1 - norm * norm_factor) > eps_norm && trvs::currTask == 0) {
    trvs::logger.warn(
      "%s normalisation input differs from expected value for an unweight field "
      "in a periodic box.", what);
  }
}

}  // namespace

// =====================================================================
// Couplings, normalisation, shot noise
// =====================================================================

double calc_coupling_coeff_2pt(int ell, int ELL, int m, int M) {
  return (2 * ell + 1) * (2 * ELL + 1)
    * trvm::wigner_3j(ell, 0, ELL, 0, 0, 0) * trvm::wigner_3j(ell, 0, ELL, m, 0, M);
}

double calc_powspec_normalisation_from_particles(ParticleCatalogue& particles, double alpha) {
  require_particles(particles);
  double norm = 0.;   // I_2
  for (int pid = 0; pid < particles.ntotal; pid++) {
    const ParticleData& p = particles[pid];
    norm += p.ws * p.nz * std::pow(p.wc, 2);
  }
  if (norm == 0.) {
    const char* msg =
      "Particle 'nz' values appear to be all zeros. "
      "Check the input catalogue contains valid 'nz' field.";
    if (trvs::currTask == 0) trvs::logger.error(msg);
    throw trvs::InvalidDataError(msg);
  }
  return 1. / (alpha * norm);
}

double calc_powspec_normalisation_from_mesh(
  ParticleCatalogue& particles, trv::ParameterSet& params, double alpha
) {
  MeshField catalogue_mesh(params, false, "`catalogue_mesh`");
  double norm_factor = catalogue_mesh.calc_grid_based_powlaw_norm(particles, 2);
  norm_factor /= std::pow(alpha, 2);
  return norm_factor;
}

double calc_powspec_normalisation_from_meshes(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  trv::ParameterSet& params, double alpha
) {
  // (Re-)align the particles in the box, restoring the positions afterwards.
  double pos_min_data[3], pos_min_rand[3];
  for (int ax = 0; ax < 3; ax++) {
    pos_min_data[ax] = particles_data.pos_min[ax];
    pos_min_rand[ax] = particles_rand.pos_min[ax];
  }
  if (params.alignment == "pad") {
    const double pad[3] = {params.padfactor, params.padfactor, params.padfactor};
    if (params.padscale == "grid") {
      ParticleCatalogue::pad_grids(particles_data, particles_rand, params.boxsize,
                                   params.ngrid, pad);
    } else if (params.padscale == "box") {
      ParticleCatalogue::pad_in_box(particles_data, particles_rand, params.boxsize, pad);
    }
  } else if (params.alignment == "centre") {
    ParticleCatalogue::centre_in_box(particles_data, particles_rand, params.boxsize);
  }
  double offset_data[3], offset_rand[3];
  for (int ax = 0; ax < 3; ax++) {
    offset_data[ax] = particles_data.pos_min[ax] - pos_min_data[ax];
    offset_rand[ax] = particles_rand.pos_min[ax] - pos_min_rand[ax];
  }

  // sum_x n_data(x) n_rand(x) on the device: the two weighted meshes are real, so
  // the sum is the k = 0 mode of ... nothing cleverer than a product reduction;
  // it is a normalisation, done once: host mirror of both meshes.
  double norm = 0.;
  {
    MeshField mesh_data(params, false, "`mesh_data`");
    MeshField mesh_rand(params, false, "`mesh_rand`");
    std::vector<double> wd(2 * (size_t)particles_data.ntotal), wr(2 * (size_t)particles_rand.ntotal);
    for (int pid = 0; pid < particles_data.ntotal; pid++) {
      wd[2 * (size_t)pid] = particles_data[pid].w; wd[2 * (size_t)pid + 1] = 0.;
    }
    for (int pid = 0; pid < particles_rand.ntotal; pid++) {
      wr[2 * (size_t)pid] = particles_rand[pid].w; wr[2 * (size_t)pid + 1] = 0.;
    }
    mesh_data.assign_weighted_field_to_mesh(
      particles_data, reinterpret_cast<double (*)[2]>(wd.data()));
    mesh_rand.assign_weighted_field_to_mesh(
      particles_rand, reinterpret_cast<double (*)[2]>(wr.data()));
    mesh_data.sync_host(); mesh_rand.sync_host();
    for (long long gid = 0; gid < params.nmesh; gid++) {
      norm += mesh_data.field[gid][0] * mesh_rand.field[gid][0];
    }
  }
  const double vol_cell = params.volume / double(params.nmesh);
  const double norm_factor = 1. / (alpha * vol_cell * norm);   // 1/I_2

  // Restore the particle positions (the reference adds the alignment offset,
  // S/twopt.cpp:223-224, as coded).
  particles_data.offset_coords(offset_data);
  particles_rand.offset_coords(offset_rand);
  return norm_factor;
}

double calc_powspec_normalisation_from_meshes(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  trv::ParameterSet& params, double alpha,
  double padding, double cellsize, const std::string& assignment
) {
  trv::ParameterSet params_norm(params);
  double boxsize_norm = (1. + padding) * std::max(
    *std::max_element(particles_data.pos_span, particles_data.pos_span + 3),
    *std::max_element(particles_rand.pos_span, particles_rand.pos_span + 3));
  int ngrid_norm = static_cast<int>(std::ceil(boxsize_norm / cellsize));
  ngrid_norm += ngrid_norm % 2;           // ensure even
  boxsize_norm = ngrid_norm * cellsize;   // enforce cell size
  for (int ax = 0; ax < 3; ax++) {
    params_norm.boxsize[ax] = boxsize_norm;
    params_norm.ngrid[ax] = ngrid_norm;
  }
  params_norm.assignment = assignment;
  params_norm.validate(true);
  return calc_powspec_normalisation_from_meshes(
    particles_data, particles_rand, params_norm, alpha);
}

double calc_powspec_shotnoise_from_particles(ParticleCatalogue& particles, double alpha) {
  (void)alpha;   // unused by the reference too (S/twopt.cpp:271-296)
  require_particles(particles);
  double shotnoise = 0.;
  for (int pid = 0; pid < particles.ntotal; pid++) {
    shotnoise += std::pow(particles[pid].ws, 2) * std::pow(particles[pid].wc, 2);
  }
  return shotnoise;
}

std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_powspec(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m
) {
  trv::ParameterSet p = small_context_params();
  auto ctx = dev::acquire_context(p);
  dev::Catalogue cd(ctx, particles_data, los_data, true);
  dev::Catalogue cr(ctx, particles_rand, los_rand, true);
  return std::conj(cat_sum(ctx.get(), cd, TRVB_W_CYLM_W2, ell, m))
    + std::pow(alpha, 2) * std::conj(cat_sum(ctx.get(), cr, TRVB_W_CYLM_W2, ell, m));
}

std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_powspec(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m
) {
  trv::ParameterSet p = small_context_params();
  auto ctx = dev::acquire_context(p);
  dev::Catalogue c(ctx, particles, los, true);
  return std::pow(alpha, 2) * std::conj(cat_sum(ctx.get(), c, TRVB_W_CYLM_W2, ell, m));
}

// =====================================================================
// Full statistics
// =====================================================================

trv::PowspecMeasurements compute_powspec(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& kbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat("Computing power spectrum from paired survey-type catalogues...");
  }
  TwoPtEngine eng(params, catalogue_data, &catalogue_rand, los_data, los_rand);
  const TwoPtSums acc = multipole_terms(eng, params, kbinning, Space::fourier, "Power spectrum");

  trv::PowspecMeasurements out;
  for (int b = 0; b < kbinning.num_bins; b++) {
    out.kbin.push_back(kbinning.bin_centres[b]);
    out.keff.push_back(acc.coord[b]);
    out.nmodes.push_back(acc.count[b]);
    out.pk_raw.push_back(norm_factor * acc.stat[b]);
    out.pk_shot.push_back(norm_factor * acc.shot[b]);
  }
  out.dim = kbinning.num_bins;
  if (trvs::currTask == 0) {
    trvs::logger.stat("... computed power spectrum from paired survey-type catalogues.");
  }
  return out;
}

trv::TwoPCFMeasurements compute_corrfunc(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing two-point correlation function from paired survey-type catalogues...");
  }
  TwoPtEngine eng(params, catalogue_data, &catalogue_rand, los_data, los_rand);
  const TwoPtSums acc = multipole_terms(eng, params, rbinning, Space::config,
                                        "Two-point correlation function");

  trv::TwoPCFMeasurements out;
  for (int b = 0; b < rbinning.num_bins; b++) {
    out.rbin.push_back(rbinning.bin_centres[b]);
    out.reff.push_back(acc.coord[b]);
    out.npairs.push_back(acc.count[b]);
    out.xi.push_back(norm_factor * acc.stat[b]);
  }
  out.dim = rbinning.num_bins;
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed two-point correlation function from paired survey-type catalogues.");
  }
  return out;
}

trv::PowspecMeasurements compute_powspec_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning kbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing power spectrum from a periodic-box simulation-type catalogue "
      "in the global plane-parallel approximation.");
  }
  warn_box_norm(catalogue_data, params, norm_factor, "Power spectrum");

  TwoPtEngine eng(params, catalogue_data, nullptr, nullptr, nullptr);
  dev::Mesh dn = eng.density_fluctuation(0, 0);           // delta n(k)
  const cdouble sn_amp = eng.shotnoise_amp(0, 0);         // \bar{N}
  std::vector<int> nmodes; std::vector<double> k; std::vector<cdouble> pk, sn;
  eng.stats_fourier(dn, dn, sn_amp, params.ELL, 0, kbinning, nmodes, k, pk, sn);

  trv::PowspecMeasurements out;
  for (int b = 0; b < kbinning.num_bins; b++) {
    out.kbin.push_back(kbinning.bin_centres[b]);
    out.keff.push_back(k[b]);
    out.nmodes.push_back(nmodes[b]);
    out.pk_raw.push_back(norm_factor * (double(2 * params.ELL + 1) * pk[b]));
    out.pk_shot.push_back(norm_factor * (double(2 * params.ELL + 1) * sn[b]));
  }
  out.dim = kbinning.num_bins;
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed power spectrum from a periodic-box simulation-type catalogue "
      "in the global plane-parallel approximation.");
  }
  return out;
}

trv::TwoPCFMeasurements compute_corrfunc_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing two-point correlation function from a periodic-box simulation-type "
      "catalogue in the global plane-parallel approximation.");
  }
  warn_box_norm(catalogue_data, params, norm_factor, "Two-point correlation function");

  TwoPtEngine eng(params, catalogue_data, nullptr, nullptr, nullptr);
  dev::Mesh dn = eng.density_fluctuation(0, 0);
  const cdouble sn_amp = eng.shotnoise_amp(0, 0);
  std::vector<int> npairs; std::vector<double> r; std::vector<cdouble> xi;
  eng.stats_config(dn, dn, sn_amp, params.ELL, 0, rbinning, npairs, r, xi);

  trv::TwoPCFMeasurements out;
  for (int b = 0; b < rbinning.num_bins; b++) {
    out.rbin.push_back(rbinning.bin_centres[b]);
    out.reff.push_back(r[b]);
    out.npairs.push_back(npairs[b]);
    out.xi.push_back(norm_factor * (double(2 * params.ELL + 1) * xi[b]));
  }
  out.dim = rbinning.num_bins;
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed two-point correlation function from a periodic-box simulation-type "
      "catalogue in the global plane-parallel approximation.");
  }
  return out;
}

trv::TwoPCFWindowMeasurements compute_corrfunc_window(
  ParticleCatalogue& catalogue_rand, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning rbinning, double alpha, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing two-point correlation function window from a random catalogue...");
  }
  TwoPtEngine eng(params, catalogue_rand, los_rand, alpha);
  const TwoPtSums acc = multipole_terms(eng, params, rbinning, Space::config,
                                        "Two-point correlation function window");

  trv::TwoPCFWindowMeasurements out;
  for (int b = 0; b < rbinning.num_bins; b++) {
    out.rbin.push_back(rbinning.bin_centres[b]);
    out.reff.push_back(acc.coord[b]);
    out.npairs.push_back(acc.count[b]);
    out.xi.push_back(norm_factor * acc.stat[b]);
  }
  out.dim = rbinning.num_bins;
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed two-point correlation function window from a random catalogue.");
  }
  return out;
}

}  // namespace trv
