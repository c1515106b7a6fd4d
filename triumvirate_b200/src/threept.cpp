// threept.cpp -- three-point estimators on the device.
//
// Same entry points, output layout and arithmetic conventions as the reference
// (S/threept.cpp:248-2619; formula sheet in SURVEY.md Appendix A), organised
// around three observations:
//   1. A shell field F_b^{lm}(x) depends on the bin b only, so each is built
//      ONCE per (l, m) and all (a, b) pairs are reduced in a single tiled pass
//      (trvb_gram_reduce) instead of 2 IFFTs + 1 reduction per pair.
//   2. sum_x F_a F_b G over the n^3 mesh only involves Fourier modes below
//      2 k_max; it equals (n^3/ns^3) x the sum over an ns^3 sub-grid with
//      ns > 4 k_max/dk, so shell fields live on the sub-grid whenever the bins
//      stay well below the Nyquist wavenumber (else ns = n).
//   3. The shot-noise mesh xi(x) does not depend on the bin pair: one IFFT per
//      term and one pass over xi for all pairs (trvb_shot_bispec_reduce).
#include "trv/threept.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <mutex>
#include <memory>
#include <set>
#include <thread>
#include <tuple>

namespace trvs = trv::sys;
namespace trvm = trv::maths;

namespace trv {

// =====================================================================
// Coupling coefficients, normalisation, shot-noise amplitude
// =====================================================================

double calc_coupling_coeff_3pt(int ell1, int ell2, int ELL, int m1, int m2, int M) {
  return double(2 * ell1 + 1) * double(2 * ell2 + 1) * double(2 * ELL + 1)
    * trvm::wigner_3j(ell1, ell2, ELL, 0, 0, 0)
    * trvm::wigner_3j(ell1, ell2, ELL, m1, m2, M);
}

void validate_multipole_coupling(trv::ParameterSet& params) {
  const double coupling = trvm::wigner_3j(params.ell1, params.ell2, params.ELL, 0, 0, 0);
  if (std::fabs(coupling) < trvm::eps_coupling) {
    const char* msg =
      "Specified three-point correlator multipole vanishes identically "
      "owing to zero-valued Wigner 3-j symbol.";
    if (trvs::currTask == 0) trvs::logger.error(msg);
    throw trvs::InvalidParameterError(msg);
  }
}

double calc_bispec_normalisation_from_particles(ParticleCatalogue& particles, double alpha) {
  if (particles.pdata == nullptr) {
    if (trvs::currTask == 0) trvs::logger.error("Particle data are uninitialised.");
    throw trvs::InvalidDataError("Particle data are uninitialised.");
  }
  double norm = 0.;
#pragma omp parallel for reduction(+:norm)
  for (int pid = 0; pid < particles.ntotal; pid++) {
    const ParticleData& p = particles.pdata[pid];
    norm += p.ws * (p.nz * p.nz) * (p.wc * p.wc * p.wc);
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

double calc_bispec_normalisation_from_mesh(
  ParticleCatalogue& particles, trv::ParameterSet& params, double alpha
) {
  MeshField catalogue_mesh(params, false, "`catalogue_mesh`");
  double norm_factor = catalogue_mesh.calc_grid_based_powlaw_norm(particles, 3);
  norm_factor /= std::pow(alpha, 3);
  return norm_factor;
}

namespace {

std::complex<double> cat_sum(trvb_ctx* ctx, dev::Catalogue& cat, int kind, int ell, int m) {
  double out[2];
  dev::check(trvb_cat_sum(ctx, cat.get(), kind, ell, m, out), "trvb_cat_sum");
  return std::complex<double>(out[0], out[1]);
}

trv::ParameterSet params_for_context(ParticleCatalogue& particles) {
  // A context is needed only for its device/stream here; any small grid does.
  (void)particles;
  trv::ParameterSet p;
  for (int ax = 0; ax < 3; ax++) { p.boxsize[ax] = 1.; p.ngrid[ax] = 4; }
  p.assignment_order = 1;
  return p;
}

}  // namespace

std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_bispec(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m
) {
  auto ctx = dev::acquire_context(params_for_context(particles_data));
  dev::Catalogue cd(ctx, particles_data, los_data, true);
  dev::Catalogue cr(ctx, particles_rand, los_rand, true);
  // Note the `+`: the code, not the header comment, is authoritative
  // (S/threept.cpp:207).
  return cat_sum(ctx.get(), cd, TRVB_W_YLM_W3, ell, m)
    + std::pow(alpha, 3) * cat_sum(ctx.get(), cr, TRVB_W_YLM_W3, ell, m);
}

std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_bispec(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m
) {
  auto ctx = dev::acquire_context(params_for_context(particles));
  dev::Catalogue c(ctx, particles, los, true);
  return std::pow(alpha, 3) * cat_sum(ctx.get(), c, TRVB_W_YLM_W3, ell, m);
}

// =====================================================================
// Shared estimator machinery
// =====================================================================

namespace {

typedef std::complex<double> cdouble;

/// Data-vector layout for a shape (S/threept.cpp:275-298 and 465-770).
struct DataVector {
  int dim = 0;
  std::vector<int> row, col;   // bin indices of each entry
};

DataVector make_data_vector(const trv::ParameterSet& params, int num_bins) {
  DataVector dv;
  const std::string& shape = params.shape;
  if (shape == "diag") {
    for (int b = 0; b < num_bins; b++) { dv.row.push_back(b); dv.col.push_back(b); }
  } else if (shape == "off-diag") {
    const int off = std::abs(params.idx_bin);
    for (int i = 0; i < num_bins - off; i++) {
      if (params.idx_bin >= 0) { dv.row.push_back(i); dv.col.push_back(i + off); }
      else { dv.row.push_back(i + off); dv.col.push_back(i); }
    }
  } else if (shape == "row") {
    for (int b = 0; b < num_bins; b++) { dv.row.push_back(params.idx_bin); dv.col.push_back(b); }
  } else if (shape == "full") {
    for (int a = 0; a < num_bins; a++)
      for (int b = 0; b < num_bins; b++) { dv.row.push_back(a); dv.col.push_back(b); }
  } else if (shape == "triu") {
    // idx = (2N - a + 1) a / 2 + (b - a), b >= a: row-major upper triangle.
    for (int a = 0; a < num_bins; a++)
      for (int b = a; b < num_bins; b++) { dv.row.push_back(a); dv.col.push_back(b); }
  } else {
    if (trvs::currTask == 0) {
      trvs::logger.error(
        "Three-point statistic form is not recognised: `form` = '%s'.", params.form.c_str());
    }
    throw trvs::InvalidParameterError(
      "Three-point statistic form is not recognised: `form` = '%s'.", params.form.c_str());
  }
  dv.dim = static_cast<int>(dv.row.size());
  return dv;
}

/// Measured cost of a batched fp64 3-D cuFFT transform (Z2D, out of place) on
/// B200, in ns per 1000 cells, for every 7-smooth length 32..735
/// (lab/fft_sweep.cu, cuFFT 11.4; profiles/r01_fft_sweep.txt).  Lengths that
/// factor as a^2 / a^3 of a small radix (100, 128, 144, 256, 343, 512, 729)
/// run at 9-12; most others at 14-24.
const struct { int n; double ns_per_kcell; } kFftCost[] = {
  {32, 31.2}, {35, 27.2}, {36, 21.0}, {40, 21.6}, {42, 20.1}, {45, 19.7}, {48, 16.8},
  {49, 14.2}, {50, 18.1}, {54, 14.6}, {56, 12.0}, {60, 14.2}, {63, 12.6}, {64, 12.7},
  {70, 12.7}, {72, 13.0}, {75, 21.9}, {80, 14.2}, {81, 9.9}, {84, 12.9}, {90, 10.9},
  {96, 16.9}, {98, 18.7}, {100, 9.6}, {105, 18.9}, {108, 11.2}, {112, 13.9}, {120, 11.9},
  {125, 11.6}, {126, 19.0}, {128, 9.9}, {135, 19.2}, {140, 16.8}, {144, 8.8}, {147, 19.2},
  {150, 14.4}, {160, 10.3}, {162, 20.0}, {168, 17.7}, {175, 14.9}, {180, 17.3}, {189, 20.6},
  {192, 10.5}, {196, 10.5}, {200, 18.5}, {210, 15.6}, {216, 10.5}, {224, 19.0}, {225, 16.9},
  {240, 17.4}, {243, 10.4}, {245, 18.6}, {250, 16.0}, {252, 14.8}, {256, 9.2}, {270, 13.1},
  {280, 19.4}, {288, 14.9}, {294, 23.6}, {300, 22.1}, {315, 15.8}, {320, 15.8}, {324, 14.4},
  {336, 18.9}, {343, 11.2}, {350, 22.2}, {360, 15.7}, {375, 19.3}, {378, 15.2}, {384, 16.4},
  {392, 17.0}, {400, 23.7}, {405, 23.1}, {420, 22.2}, {432, 20.0}, {441, 14.9}, {448, 19.3},
  {450, 17.4}, {480, 22.4}, {486, 30.2}, {490, 20.6}, {500, 16.5}, {504, 14.2}, {512, 10.5},
  {525, 24.4}, {540, 22.1}, {560, 20.4}, {567, 28.0}, {576, 22.1}, {588, 23.3}, {600, 23.9},
  {625, 14.4}, {630, 28.7}, {640, 18.5}, {648, 18.9}, {672, 24.2}, {675, 22.7}, {686, 20.0},
  {700, 21.9}, {720, 23.7}, {729, 12.5}, {735, 30.6}
};

/// Transform length >= n with the least measured cost n^3 * cost(n) among the
/// candidates up to 1.4 n; beyond the table, the next 5-smooth multiple of 16.
int next_fft_size(int n) {
  int best = 0; double best_cost = 0.;
  const int lo = std::max(n, 2), hi = std::max(lo + 16, (int)(1.4 * lo));
  for (const auto& e : kFftCost) {
    if (e.n < lo || e.n > hi) continue;
    const double cost = e.ns_per_kcell * double(e.n) * double(e.n) * double(e.n);
    if (best == 0 || cost < best_cost) { best = e.n; best_cost = cost; }
  }
  if (best) return best;
  for (int v = (lo + 15) / 16 * 16;; v += 16) {
    int r = v;
    for (int p : {2, 3, 5}) while (r % p == 0) r /= p;
    if (r == 1) return v;
  }
}

std::vector<int> distinct_sorted(const std::vector<int>& v) {
  std::set<int> s(v.begin(), v.end());
  return std::vector<int>(s.begin(), s.end());
}

/// One surviving (m1, m2, M) term with its coupling and mirror factors
/// (S/threept.cpp:356-419, 1580-1616).
struct Term {
  int m1, m2, M;
  double coupling;
  double factor_mirror;       // 0 for (0,0,0), else (-1)^ELL
  double factor_mirror_w3j;   // 0 for (0,0,0), else 1
};

std::vector<Term> enumerate_terms(const trv::ParameterSet& params, bool survey) {
  std::vector<Term> terms;
  std::vector<SphericalOrderTriplet> computed;
  for (int m1 = -params.ell1; m1 <= params.ell1; m1++) {
    for (int m2 = -params.ell2; m2 <= params.ell2; m2++) {
      const int Mlo = survey ? -params.ELL : 0;
      const int Mhi = survey ? params.ELL : 0;
      for (int M = Mlo; M <= Mhi; M++) {
        SphericalOrderTriplet cur = {m1, m2, M};
        bool redundant = false;
        for (const SphericalOrderTriplet& t : computed) {
          if (cur.is_inverse(t)) redundant = true;
        }
        if (redundant) continue;
        const double coupling = calc_coupling_coeff_3pt(
          params.ell1, params.ell2, params.ELL, m1, m2, M);
        if (std::fabs(coupling) < trvm::eps_coupling) continue;
        Term t;
        t.m1 = m1; t.m2 = m2; t.M = M; t.coupling = coupling;
        t.factor_mirror = cur.is_zeros() ? 0. : std::pow(-1., params.ELL);
        t.factor_mirror_w3j = cur.is_zeros() ? 0. : 1.;
        terms.push_back(t);
        computed.push_back(cur);
      }
    }
  }
  return terms;
}

/// Everything an estimator call keeps on the device.
class Engine {
 public:
  Engine(trv::ParameterSet& params, ParticleCatalogue& data, ParticleCatalogue* rand,
         LineOfSight* los_data, LineOfSight* los_rand)
    : params_(params), survey_(rand != nullptr) {
    ctx_ = dev::acquire_context(params);
    c_ = ctx_.get();
    dev::profile_reset();
    data_.reset(new dev::Catalogue(ctx_, data, los_data, survey_));
    if (survey_) rand_.reset(new dev::Catalogue(ctx_, *rand, los_rand, true));
    ndata_ = data.ntotal;
    if (survey_) alpha_ = data.wstotal / rand->wstotal;   // S/threept.cpp:269
    init_common();
    dev::profile_mark(c_, "upload");
  }

  /// Window mode: one (random) catalogue with lines of sight and a given
  /// alpha contrast (S/threept.cpp:2696-2702, S/field.cpp:1325-1362,1449-1489).
  Engine(trv::ParameterSet& params, ParticleCatalogue& rand, LineOfSight* los_rand,
         double alpha)
    : params_(params), survey_(true), window_(true) {
    ctx_ = dev::acquire_context(params);
    c_ = ctx_.get();
    dev::profile_reset();
    data_.reset(new dev::Catalogue(ctx_, rand, los_rand, true));
    ndata_ = rand.ntotal;
    alpha_ = alpha;
    init_common();
    dev::profile_mark(c_, "upload");
  }

  /// Periodic-box catalogue given as coordinate arrays (host or device);
  /// unit weights (S/threept.cpp:1543-1558).
  Engine(trv::ParameterSet& params, long long n, const double* x, const double* y,
         const double* z, bool on_device)
    : params_(params), survey_(false) {
    ctx_ = dev::acquire_context(params);
    c_ = ctx_.get();
    dev::profile_reset();
    ndata_ = n;
    init_common();
    if (!on_device && !params.deterministic) {
      // Host arrays, throughput mode: the upload is streamed together with the first
      // assignment (density_fluctuation), chunk by chunk.
      host_x_ = x; host_y_ = y; host_z_ = z;
    } else {
      data_.reset(new dev::Catalogue(ctx_, n, x, y, z, nullptr, nullptr, on_device));
    }
    dev::profile_mark(c_, "upload");
  }

  void init_common() {
    vol_ = params_.volume;
    vol_cell_ = vol_ / double(params_.nmesh);
    mode_ = params_.deterministic ? 1 : 0;
    dev::check(trvb_ctx_set_deterministic(c_, mode_), "trvb_ctx_set_deterministic");
  }

  trvb_ctx* ctx() { return c_; }
  std::shared_ptr<trvb_ctx> shared() { return ctx_; }
  double alpha() const { return alpha_; }
  double vol() const { return vol_; }
  double vol_cell() const { return vol_cell_; }
  bool survey() const { return survey_; }
  bool window() const { return window_; }
  long long ndata() const { return ndata_; }

  /// Fourier transform of the (L, M)-weighted number-density fluctuation,
  /// delta n_LM(k), including the dV factor of S/field.cpp:1503-1510:
  ///   box    : unit weights, mean subtracted (S/field.cpp:1229-1244)
  ///   survey : y_LM w for data minus alpha x randoms (S/field.cpp:1246-1323)
  /// Weights are assigned WITHOUT the 1/dV density factor, which cancels the
  /// dV of the forward transform exactly (up to rounding).
  dev::Mesh density_fluctuation(int L, int M) {
    const bool real_field = (M == 0);
    dev::Mesh x(ctx_, c_, real_field ? TRVB_REAL : TRVB_COMPLEX);
    if (!survey_ && !data_) {
      // Pending host arrays: chunked upload, each chunk tile-sorted and spread while the
      // next one is on the wire (trvb_cat_create_assign).
      trvb_cat* cat = nullptr;
      dev::check(trvb_cat_create_assign(c_, &cat, ndata_, host_x_, host_y_, host_z_, 1., x.view()),
                 "trvb_cat_create_assign");
      data_.reset(new dev::Catalogue(ctx_, cat));
    } else if (!survey_) {
      assign(*data_, TRVB_W_UNIT, 0, 0, 1., false, x);
    } else if (window_) {
      assign(*data_, TRVB_W_YLM_W, L, M, alpha_, false, x);
    } else {
      assign(*data_, TRVB_W_YLM_W, L, M, 1., false, x);
      assign(*rand_, TRVB_W_YLM_W, L, M, -alpha_, true, x);
    }
    dev::Mesh k = forward(x);
    if (!survey_) {
      // Mean subtraction touches the k = 0 mode only: FFT[nbar dV] = N delta_k0.
      dev::check(trvb_kmesh_add_zero_mode(c_, k.view(), -double(ndata_)),
                 "trvb_kmesh_add_zero_mode");
    }
    return k;
  }

  /// Box catalogue, multi-GPU runs with a distributed mesh phase (trvb_dmesh_*): the ranks of
  /// `comm` assign and transform slabs of the mesh; every rank gets the modes that the grid
  /// of `sub` represents as a HALF mesh of sub's extents, which the device layer reads as a
  /// low-|k| view of the full-grid spectrum.  Collective.
  dev::Mesh density_fluctuation_dist(trvb_comm* comm, trvb_ctx* sub) {
    if (!data_) {   // pending host arrays: the whole catalogue goes up first
      data_.reset(new dev::Catalogue(ctx_, ndata_, host_x_, host_y_, host_z_, nullptr, nullptr, false));
    }
    const double* x = nullptr; const double* y = nullptr; const double* z = nullptr;
    dev::check(trvb_cat_positions(data_->get(), &x, &y, &z), "trvb_cat_positions");
    dev::check(trvb_dmesh_get(c_, comm, &dmesh_), "trvb_dmesh_get");
    // Mean subtraction touches the k = 0 mode only: FFT[nbar dV] = N delta_k0.
    dev::check(trvb_dmesh_density(dmesh_, ndata_, x, y, z, -double(ndata_)), "trvb_dmesh_density");
    trvs::count_fft += 1;
    dev::Mesh k(ctx_, sub, TRVB_HALF);
    dev::check(trvb_dmesh_gather_lowk(dmesh_, sub, k.view()), "trvb_dmesh_gather_lowk");
    return k;
  }
  trvb_dmesh* dmesh() { return dmesh_; }
  ~Engine() {
    if (dmesh_) trvb_dmesh_forget_lowk(dmesh_);
    if (fused_lowk_) trvb_ctx_forget_lowk(c_);
  }

  /// Box catalogue on one GPU, throughput mode (trvb_box_fields_fused): delta n(k) is
  /// needed for two things only -- the modes that the grid of `sub` represents and the
  /// shot-noise xi(r) of S/threept.cpp:1977-2110 -- so the x passes of the forward and
  /// the inverse full-grid transform are fused around the spectrum arithmetic and the
  /// full-grid delta n(k) is never stored.  Returns the low-|k| modes as a HALF mesh of
  /// sub's extents (read as a view of the full-grid spectrum, like the distributed
  /// mesh's) and fills `xi` (REAL, full grid).
  dev::Mesh density_fluctuation_fused(trvb_ctx* sub, dev::Mesh& xi) {
    dev::Mesh x(ctx_, c_, TRVB_REAL);
    if (!data_) {
      trvb_cat* cat = nullptr;
      dev::check(trvb_cat_create_assign(c_, &cat, ndata_, host_x_, host_y_, host_z_, 1., x.view()),
                 "trvb_cat_create_assign");
      data_.reset(new dev::Catalogue(ctx_, cat));
    } else {
      assign(*data_, TRVB_W_UNIT, 0, 0, 1., false, x);
    }
    dev::Mesh k(ctx_, sub, TRVB_HALF);
    xi = dev::Mesh(ctx_, c_, TRVB_REAL);
    // fa = delta n (mean subtracted: FFT[nbar dV] = N delta_k0), fb = N_00 (not subtracted);
    // S = N (S/threept.cpp:1977-1979)
    const double S[2] = {double(ndata_), 0.};
    dev::check(trvb_box_fields_fused(c_, sub, x.view(), -double(ndata_), 0., S, k.view(), xi.view()),
               "trvb_box_fields_fused");
    fused_lowk_ = true;
    trvs::count_fft += 1;
    trvs::count_ifft += 1;
    return k;
  }

  /// N_LM(k): conj(y_LM) w^2 for data plus alpha^2 x randoms
  /// (S/field.cpp:1364-1447); box: unit weights, no mean subtraction.
  dev::Mesh quadratic_field(int L, int M) {
    const bool real_field = (M == 0);
    dev::Mesh x(ctx_, c_, real_field ? TRVB_REAL : TRVB_COMPLEX);
    if (!survey_) {
      assign(*data_, TRVB_W_UNIT, 0, 0, 1., false, x);
    } else if (window_) {
      assign(*data_, TRVB_W_CYLM_W2, L, M, std::pow(alpha_, 2), false, x);
    } else {
      assign(*data_, TRVB_W_CYLM_W2, L, M, 1., false, x);
      assign(*rand_, TRVB_W_CYLM_W2, L, M, std::pow(alpha_, 2), true, x);
    }
    return forward(x);
  }

  /// Box only: N_00(k) is delta n_00(k) with the k = 0 mode restored
  /// (S/threept.cpp:1554-1558 assigns and transforms the same catalogue a
  /// second time); expressed as a shifted view of the same buffer.
  trvb_mesh quadratic_view_of_fluctuation(const dev::Mesh& dn) const {
    return dn.view_k0(double(ndata_));
  }

  cdouble shotnoise_amp(int L, int M) {
    if (!survey_) return cdouble(double(ndata_), 0.);   // S/threept.cpp:1977-1979
    if (window_) {                                       // S/threept.cpp:210-236
      return std::pow(alpha_, 3) * cat_sum(c_, *data_, TRVB_W_YLM_W3, L, M);
    }
    return cat_sum(c_, *data_, TRVB_W_YLM_W3, L, M)
      + std::pow(alpha_, 3) * cat_sum(c_, *rand_, TRVB_W_YLM_W3, L, M);
  }

  void upload_sjl(trvm::SphericalBesselCalculator& sj) {
    dev::check(trvb_sjl_table(c_, sj.order, sj.y.data(), sj.c.data(),
                              static_cast<int>(sj.y.size()), sj.step), "trvb_sjl_table");
  }

  /// Number of additional meshes of `bytes` each that fit in free HBM.
  int mesh_capacity(size_t bytes, size_t reserve = 0) {
    size_t free_b = 0, total_b = 0;
    dev::check(trvb_mem_info(c_, &free_b, &total_b), "trvb_mem_info");
    const double usable = 0.92 * double(free_b) - double(reserve);
    return usable <= 0. ? 0 : static_cast<int>(usable / double(bytes));
  }

 private:
  void assign(dev::Catalogue& cat, int kind, int L, int M, double scale, bool accumulate,
              dev::Mesh& mesh) {
    dev::check(trvb_assign(c_, cat.get(), kind, L, M, scale, /*density_units=*/0,
                           accumulate ? 1 : 0, /*shifted=*/0, mode_, mesh.view()),
               "trvb_assign");
  }

  dev::Mesh forward(dev::Mesh& x) {
    if (x.layout() == TRVB_REAL) {
      dev::Mesh k(ctx_, c_, TRVB_HALF);
      dev::check(trvb_fft_forward(c_, x.view(), k.view(), 1.), "trvb_fft_forward");
      trvs::count_fft += 1;
      return k;
    }
    dev::check(trvb_fft_forward(c_, x.view(), x.view(), 1.), "trvb_fft_forward");
    trvs::count_fft += 1;
    return std::move(x);
  }

  trv::ParameterSet& params_;
  bool survey_;
  bool window_ = false;
  std::shared_ptr<trvb_ctx> ctx_;
  trvb_ctx* c_ = nullptr;
  std::unique_ptr<dev::Catalogue> data_, rand_;
  long long ndata_ = 0;
  double alpha_ = 1.;
  double vol_ = 0., vol_cell_ = 0.;
  int mode_ = 0;
  const double* host_x_ = nullptr;   // box arrays awaiting the streamed upload
  const double* host_y_ = nullptr;
  const double* host_z_ = nullptr;
  trvb_dmesh* dmesh_ = nullptr;      // owned by the context
  bool fused_lowk_ = false;          // the context's low-|k| view points at one of our meshes
};

/// `count` consecutive meshes of one grid in a single device allocation (the
/// unit the batched transforms write and the pair reduction reads).
class Slab {
 public:
  Slab(std::shared_ptr<trvb_ctx> owner, trvb_ctx* grid, int layout, int count)
    : owner_(owner), stride_(trvb_mesh_bytes(grid, layout)), count_(count) {
    dev::check(trvb_malloc(owner_.get(), &data_, stride_ * (size_t)std::max(count, 1)),
               "trvb_malloc");
    trvs::gbytesMemGPU += gib();
    trvs::update_maxmem(true);
  }
  ~Slab() {
    if (data_) { trvb_free(owner_.get(), data_); trvs::gbytesMemGPU -= gib(); }
  }
  Slab(const Slab&) = delete;
  Slab& operator=(const Slab&) = delete;
  void* data() const { return data_; }
  void* mesh(int i) const { return static_cast<char*>(data_) + stride_ * (size_t)i; }
  int count() const { return count_; }
 private:
  double gib() const { return double(stride_) * count_ / (1024. * 1024. * 1024.); }
  std::shared_ptr<trvb_ctx> owner_;
  void* data_ = nullptr;
  size_t stride_;
  int count_;
};

/// A plain device block with RAII ownership.
class PlaneBlock {
 public:
  PlaneBlock(std::shared_ptr<trvb_ctx> owner, size_t bytes) : owner_(owner) {
    dev::check(trvb_malloc(owner_.get(), &data_, bytes), "trvb_malloc");
  }
  ~PlaneBlock() { if (data_) trvb_free(owner_.get(), data_); }
  PlaneBlock(const PlaneBlock&) = delete;
  PlaneBlock& operator=(const PlaneBlock&) = delete;
  void* data() const { return data_; }
 private:
  std::shared_ptr<trvb_ctx> owner_;
  void* data_ = nullptr;
};

/// Blocked all-pairs reduction: out[idx] = sum_x A_{row(idx)} B_{col(idx)} G
/// for the entries of `dv` selected by `active`.  `get_a(bins)` and
/// `get_b(bins)` return a slab whose mesh i (layout of G, on `grid`) holds the
/// field of bin bins[i]; they may hand back a slab cached from an earlier term.
/// `conj_b`: the B fields are used complex-conjugated (mirror harmonics, see
/// mirror_harmonics()).
template <class MakeA, class MakeB>
void reduce_pairs(Engine& eng, trvb_ctx* grid, const DataVector& dv,
                  const std::vector<char>& active, bool same_fields,
                  trvb_mesh G, MakeA get_a, MakeB get_b,
                  std::vector<cdouble>& out, bool conj_b = false) {
  out.assign(dv.dim, cdouble(0., 0.));
  std::vector<int> rows_all, cols_all;
  for (int i = 0; i < dv.dim; i++) {
    if (active[i]) { rows_all.push_back(dv.row[i]); cols_all.push_back(dv.col[i]); }
  }
  if (rows_all.empty()) return;
  const std::vector<int> rows = distinct_sorted(rows_all);
  const std::vector<int> cols = distinct_sorted(cols_all);

  // Memory plan.  The batched shell transform keeps a transient half-spectrum slab
  // and cuFFT's work area, both bounded by its sub-batching (<= 6 GiB each, or the
  // whole batch when that is smaller); everything else can hold fields.
  const size_t bytes = trvb_mesh_bytes(grid, G.layout);
  std::vector<int> all_bins = rows;
  all_bins.insert(all_bins.end(), cols.begin(), cols.end());
  const int want = same_fields ? static_cast<int>(distinct_sorted(all_bins).size())
                               : static_cast<int>(rows.size() + cols.size());
  const size_t transient = std::min<size_t>(size_t(14) << 30, 2 * bytes * size_t(want));
  const int cap = eng.mesh_capacity(bytes, transient);
  dev::profile_mark(eng.ctx(), "pairs:capacity");
  if (cap < 2) {
    throw trvs::DeviceError(
      "Insufficient device memory: fewer than two %zu-byte shell meshes fit.", bytes);
  }
  int block_r = static_cast<int>(rows.size());
  int block_c = static_cast<int>(cols.size());
  if (want > cap) {
    const int half = std::max(1, cap / 2);
    block_r = std::min(block_r, half);
    block_c = std::min(block_c, half);
  }

  for (size_t r0 = 0; r0 < rows.size(); r0 += block_r) {
    const size_t r1 = std::min(rows.size(), r0 + block_r);
    const std::vector<int> bins_a(rows.begin() + r0, rows.begin() + r1);
    const std::shared_ptr<Slab> fa_holder = get_a(bins_a);
    const Slab& fa = *fa_holder;
    dev::profile_mark(eng.ctx(), "pairs:fields_a");
    std::map<int, int> ia_of;
    std::vector<const void*> pa;
    for (size_t i = 0; i < bins_a.size(); i++) { ia_of[bins_a[i]] = (int)i; pa.push_back(fa.mesh((int)i)); }
    for (size_t c0 = 0; c0 < cols.size(); c0 += block_c) {
      const size_t c1 = std::min(cols.size(), c0 + block_c);
      // Columns of this block that pair with one of its rows; those whose field
      // already exists among the rows share it.
      std::set<int> cols_used;
      for (int i = 0; i < dv.dim; i++) {
        if (!active[i] || !ia_of.count(dv.row[i])) continue;
        if (dv.col[i] >= cols[c0] && dv.col[i] <= cols[c1 - 1]) cols_used.insert(dv.col[i]);
      }
      if (cols_used.empty()) continue;
      std::vector<int> bins_b;
      for (int cb : cols_used) {
        if (!(same_fields && ia_of.count(cb))) bins_b.push_back(cb);
      }
      std::shared_ptr<Slab> fb;
      if (!bins_b.empty()) fb = get_b(bins_b);
      dev::profile_mark(eng.ctx(), "pairs:fields_b");
      std::map<int, int> ib_of;
      std::vector<const void*> pb;
      int next_own = 0;
      for (int cb : cols_used) {
        ib_of[cb] = (int)pb.size();
        if (same_fields && ia_of.count(cb)) pb.push_back(fa.mesh(ia_of[cb]));
        else pb.push_back(fb->mesh(next_own++));
      }
      // Pair list restricted to this block.
      std::vector<int> ia, ib, where;
      for (int i = 0; i < dv.dim; i++) {
        if (!active[i]) continue;
        auto fr = ia_of.find(dv.row[i]); auto fc = ib_of.find(dv.col[i]);
        if (fr == ia_of.end() || fc == ib_of.end()) continue;
        ia.push_back(fr->second); ib.push_back(fc->second); where.push_back(i);
      }
      if (ia.empty()) continue;
      std::vector<double> sums(2 * ia.size());
      dev::check(trvb_gram_reduce(grid, pa.data(), (int)pa.size(), pb.data(), (int)pb.size(),
                                  G, ia.data(), ib.data(), (int)ia.size(), conj_b ? 1 : 0,
                                  sums.data()), "trvb_gram_reduce");
      dev::profile_mark(eng.ctx(), "pairs:gram");
      for (size_t p = 0; p < ia.size(); p++) {
        out[where[p]] = cdouble(sums[2 * p], sums[2 * p + 1]);
      }
    }
  }
}

}  // namespace

/// Owner rank of every data-vector entry.  An entry (row bin, column bin) needs
/// the shell fields of both bins on its rank and those inverse transforms are
/// what the pair phase costs (two orders of magnitude more than a pair product), so
/// a rank should own a compact rectangle of the (row, column) matrix and the shares
/// should be equal in FIELDS rather than in entries: the entries are ordered by
/// (strip of h consecutive rows, column -- serpentine, so that a share crossing into
/// the next strip keeps its columns --, row) and that order is cut greedily under the
/// smallest cost cap (distinct fields + pairs / 100) that needs no more than `world`
/// shares; h is scanned around sqrt(entries per rank).  `same_fields`: row and column
/// bins index the same set of fields (equal degrees), so a bin on both sides of a
/// share counts once.  For `diag`, `off-diag` and `row` shapes this degenerates to
/// contiguous bin ranges.
std::vector<int> partition_owners(const std::vector<int>& row, const std::vector<int>& col,
                                  int world, bool same_fields, double handicap_last) {
  const int dim = static_cast<int>(row.size());
  std::vector<int> owner(dim, 0);
  if (world <= 1 || dim == 0) return owner;
  // The search below costs milliseconds (24 orderings x 40 bisection steps over all
  // entries) and every estimator call of a run asks for the same partition: remembered.
  typedef std::tuple<std::vector<int>, std::vector<int>, int, bool, double> Key;
  static std::mutex cache_mutex;
  static std::map<Key, std::vector<int> > cache;
  const Key key(row, col, world, same_fields, handicap_last);
  {
    std::lock_guard<std::mutex> lock(cache_mutex);
    auto hit = cache.find(key);
    if (hit != cache.end()) return hit->second;
  }
  const std::vector<int> rows = distinct_sorted(row), cols = distinct_sorted(col);
  auto rank_in = [](const std::vector<int>& sorted, int v) {
    return static_cast<int>(std::lower_bound(sorted.begin(), sorted.end(), v) - sorted.begin());
  };
  // Field ids: rows and columns share ids when they index the same fields.
  std::vector<int> all = rows;
  all.insert(all.end(), cols.begin(), cols.end());
  const std::vector<int> fields = distinct_sorted(all);
  const int nfield = same_fields ? static_cast<int>(fields.size())
                                 : static_cast<int>(rows.size() + cols.size());
  std::vector<int> fr(dim), fc(dim), rr(dim), cc(dim);
  for (int i = 0; i < dim; i++) {
    rr[i] = rank_in(rows, row[i]); cc[i] = rank_in(cols, col[i]);
    fr[i] = same_fields ? rank_in(fields, row[i]) : rr[i];
    fc[i] = same_fields ? rank_in(fields, col[i]) : static_cast<int>(rows.size()) + cc[i];
  }
  const double pair_weight = 0.01;

  // Greedy cut of `order` under a cost cap; returns the number of shares.
  std::vector<int> stamp(nfield, -1);
  // The last share may carry other work worth `handicap_last` fields (the shot-noise
  // branch of the bispectrum): its cap is lower by that much, possibly leaving it empty.
  auto cut = [&](const std::vector<int>& order, double cap, std::vector<int>* out) {
    int share = 0, nf = 0, np = 0;
    std::fill(stamp.begin(), stamp.end(), -1);
    for (int idx : order) {
      for (;;) {
        const int add = (stamp[fr[idx]] != share) + (fc[idx] != fr[idx] && stamp[fc[idx]] != share);
        const double cap_here = (share == world - 1) ? cap - handicap_last : cap;
        if (nf + add + pair_weight * (np + 1) <= cap_here) {
          stamp[fr[idx]] = share; stamp[fc[idx]] = share;
          nf += add; np++;
          if (out) (*out)[idx] = share;
          break;
        }
        if (np == 0 || share == world - 1) return world + 1;   // does not fit under this cap
        share++; nf = 0; np = 0;
      }
    }
    return share + 1;
  };

  const double per = double(dim) / double(world);
  const int h0 = std::max(1, static_cast<int>(std::floor(std::sqrt(per) + 0.5)));
  double best_cap = 0.; std::vector<int> best_order;
  const int h_lo = std::max(1, h0 / 2), h_hi = 2 * h0 + 1;
  const int h_step = std::max(1, (h_hi - h_lo + 23) / 24);   // at most 24 candidates
  for (int h = h_lo; h <= h_hi; h += h_step) {
    std::vector<std::array<int, 4> > keyed(dim);
    for (int i = 0; i < dim; i++) {
      const int strip = rr[i] / h;
      keyed[i] = {strip, (strip % 2 == 0) ? cc[i] : -cc[i], rr[i], i};
    }
    std::sort(keyed.begin(), keyed.end());
    std::vector<int> order(dim);
    for (int t = 0; t < dim; t++) order[t] = keyed[t][3];
    // one share always fits under hi
    double lo = 0., hi = nfield + pair_weight * dim + 1. + std::max(0., handicap_last);
    for (int it = 0; it < 40; it++) {
      const double mid = 0.5 * (lo + hi);
      if (cut(order, mid, nullptr) <= world) hi = mid; else lo = mid;
    }
    if (best_order.empty() || hi < best_cap - 1.e-9) { best_cap = hi; best_order = order; }
  }
  cut(best_order, best_cap, &owner);
  {
    std::lock_guard<std::mutex> lock(cache_mutex);
    if (cache.size() >= 64) cache.clear();
    cache[key] = owner;
  }
  return owner;
}

std::vector<int> partition_owners(const trv::ParameterSet& params, int num_bins, int world) {
  const DataVector dv = make_data_vector(params, num_bins);
  return partition_owners(dv.row, dv.col, world, params.ell1 == params.ell2);
}

namespace {

/// For a REAL source field (its spectrum is stored as a half spectrum) the shell fields of
/// the mirror harmonic satisfy F_{l,-m}(x) = (-1)^(l+m) conj F_{l,m}(x)
/// [y_{l,-m} = (-1)^m conj y_{lm}, y_lm(-k) = (-1)^l y_lm(k), delta n(-k) = conj delta n(k),
/// every other factor real and even], so a term with l1 = l2 and m2 = -m1 needs only the
/// (l1, m1) transforms: the pair reduction takes conj(B) and the sign is applied after.
/// The identity fails for modes ON a Nyquist plane (index n/2 stands for -n/2 only, it has
/// no partner), so it is used only when no shell can reach one: k_max <= (n_a/2) dk_a on
/// every axis.  The spherical-Bessel fields of the 3PCF weight ALL modes and never qualify.
bool mirror_harmonics(const trv::ParameterSet& params, const Term& t, const dev::Mesh& source,
                      double kmax) {
  if (!(params.ell1 == params.ell2 && t.m2 == -t.m1 && t.m1 != 0
        && source.layout() == TRVB_HALF && kmax > 0.)) return false;
  for (int ax = 0; ax < 3; ax++) {
    const double k_nyq = (params.ngrid[ax] / 2) * (2. * M_PI / params.boxsize[ax]);
    if (kmax > k_nyq) return false;
  }
  return true;
}

std::vector<char> active_entries(const trv::ParameterSet& params, const DataVector& dv) {
  std::vector<char> active(dv.dim, 0);
  const std::vector<int> owner = partition_owners(dv.row, dv.col, params.part_count,
                                                  params.ell1 == params.ell2);
  for (int i = 0; i < dv.dim; i++) active[i] = owner[i] == params.part_rank;
  return active;
}

/// Bispectrum work split.  The shot-noise branch (one full-grid inverse FFT and its
/// reductions) does not depend on the pair partition, so with two or more ranks it
/// goes to the LAST rank for every entry; the pair entries are dealt to all ranks with
/// the last one handicapped by the cost of that branch in units of one shell field
/// (measured: a full-grid shot-noise pass costs 1.7x a sub-grid shell transform per
/// cell).
struct BispecShare {
  std::vector<char> pairs;   // entries whose raw bispectrum this rank computes
  std::vector<char> shot;    // entries whose shot noise this rank computes
  bool any_pairs = false, any_shot = false;
};

BispecShare bispec_share(const trv::ParameterSet& params, const DataVector& dv,
                         double shot_cost_in_fields) {
  BispecShare sh;
  sh.pairs.assign(dv.dim, 0); sh.shot.assign(dv.dim, 0);
  const int world = params.part_count, rank = params.part_rank;
  if (world < 2) {
    sh.pairs.assign(dv.dim, 1); sh.shot.assign(dv.dim, 1);
  } else {
    // The last rank's share of the pairs is cut short by what the shot-noise branch costs
    // (none at all when that exceeds a fair share).
    const std::vector<int> owner = partition_owners(dv.row, dv.col, world,
                                                    params.ell1 == params.ell2,
                                                    shot_cost_in_fields);
    for (int i = 0; i < dv.dim; i++) {
      sh.pairs[i] = owner[i] == rank;
      sh.shot[i] = rank == world - 1;
    }
  }
  for (int i = 0; i < dv.dim; i++) { sh.any_pairs |= sh.pairs[i]; sh.any_shot |= sh.shot[i]; }
  return sh;
}

/// Sub-grid extents for shells reaching k_max (section 2 of the file header).
void choose_subgrid(const trv::ParameterSet& params, double kmax, int nsub[3]) {
  bool coarsen = true;
  const char* env_z = std::getenv("TRV_NO_ZPASS");
  const bool no_zpass = env_z != nullptr && env_z[0] == '1';
  for (int ax = 0; ax < 3; ax++) {
    const double dk = 2. * M_PI / params.boxsize[ax];
    const long long mcut = static_cast<long long>(std::floor(kmax / dk)) + 1;
    const long long need = 4 * mcut + 2;
    if (need >= params.ngrid[ax]) { coarsen = false; break; }
    nsub[ax] = next_fft_size(static_cast<int>(need));
    // Throughput mode builds real shell fields with the pruned per-axis transform, whose
    // last pass is hand-written for a set of extents (trvb_shell_zpass_supported): the
    // smallest of those within 25 % of the need beats the table above, which ranks dense
    // 3-D cuFFT transforms.  The deterministic mode keeps the dense transform and the table.
    if (!params.deterministic && !no_zpass) {
      for (long long v = need; 4 * v <= 5 * need; v++) {
        if (trvb_shell_zpass_supported(static_cast<int>(v))) { nsub[ax] = static_cast<int>(v); break; }
      }
    }
    if (nsub[ax] >= params.ngrid[ax]) { coarsen = false; break; }
  }
  const char* env = std::getenv("TRV_NO_SUBGRID");
  if (env != nullptr && std::string(env) == "1") coarsen = false;
  if (!coarsen) for (int ax = 0; ax < 3; ax++) nsub[ax] = params.ngrid[ax];
}

}  // namespace

// =====================================================================
// Bispectrum
// =====================================================================

namespace {

trv::BispecMeasurements bispec_impl(
  Engine& eng, trv::ParameterSet& params, trv::Binning& kbinning, double norm_factor
) {
  const bool survey = eng.survey();
  const cdouble factor_phase = std::pow(trvm::M_I, params.ell1 + params.ell2);
  const int nb = kbinning.num_bins;
  const DataVector dv = make_data_vector(params, nb);
  // Sub-grid for the shell fields (decided here: the work split weighs the shot-noise
  // branch on the full grid against shell transforms on the sub-grid).
  int nsub[3];
  choose_subgrid(params, kbinning.bin_edges.back(), nsub);
  const double shot_cost_in_fields = 1.7 * double(params.nmesh)
    / (double(nsub[0]) * double(nsub[1]) * double(nsub[2]));
  const std::vector<Term> terms = enumerate_terms(params, survey);

  // Slab mode (two or more ranks, throughput mode, real shell fields on a true sub-grid):
  // instead of dealing whole PAIRS -- which makes every rank transform ~2/sqrt(R) of the
  // shells -- the ranks split the x-PLANES of the sub-grid.  Each rank builds every shell
  // field on its own planes only (pruned transforms, trvb_shell_slab_batch), reduces ALL
  // pairs over those cells, and the sum over ranks completes each entry.  The last rank
  // still owns the shot-noise branch and takes a thinner slab (none when that branch
  // outweighs a fair share).
  bool slab_mode = params.part_count > 1 && !params.deterministic
    && nsub[0] != params.ngrid[0] && (params.ell1 % 2 == 0) && (params.ell2 % 2 == 0)
    && (!survey || params.ELL % 2 == 0);
  for (const Term& t : terms) slab_mode = slab_mode && t.m1 == 0 && t.m2 == 0 && t.M == 0;
  {
    const char* env = std::getenv("TRV_NO_SLAB");
    if (env != nullptr && env[0] == '1') slab_mode = false;
  }
  // Distributed mesh phase (box catalogues, NCCL communicator attached, grid extents that
  // split over the ranks): assignment and the full-grid transforms are shared as well --
  // x-slabs of the mesh, one all-to-all per transform (trvb_dmesh_*) -- every rank reads
  // the low-|k| modes it needs from a gathered copy, and the shot-noise branch is split
  // like the rest, so the planes of the sub-grid are dealt evenly.  TRV_NO_DIST_MESH=1:
  // replicated mesh.
  trvb_comm* comm = dev::process_comm();
  bool dist_mesh = slab_mode && !survey && comm != nullptr
    && trvb_comm_size(comm) == params.part_count && trvb_comm_rank(comm) == params.part_rank
    && trvb_dmesh_supported(eng.ctx(), params.part_count)
    && params.boxsize[0] * params.ngrid[1] == params.boxsize[1] * params.ngrid[0]
    && params.boxsize[0] * params.ngrid[2] == params.boxsize[2] * params.ngrid[0];
  {
    const char* env = std::getenv("TRV_NO_DIST_MESH");
    if (env != nullptr && env[0] == '1') dist_mesh = false;
  }
  int slab_x0 = 0, slab_nx = 0;
  if (slab_mode) {
    const int R = params.part_count, r = params.part_rank;
    const double W = 1.2 * nb * (params.ell1 == params.ell2 ? 1. : 2.) + 1.;   // fields + pair products
    const double f_last = dist_mesh ? 1. / R
      : std::max(0., (W - (R - 1) * shot_cost_in_fields) / (R * W));
    const int n_last = static_cast<int>(std::floor(f_last * nsub[0] + 0.5));
    const int rest = nsub[0] - n_last;
    auto first_plane = [&](int q) {   // ranks 0 .. R-2 share `rest` planes evenly
      return static_cast<int>((long long)rest * q / (R - 1));
    };
    if (r < R - 1) { slab_x0 = first_plane(r); slab_nx = first_plane(r + 1) - slab_x0; }
    else { slab_x0 = rest; slab_nx = n_last; }
  }

  BispecShare share;
  if (slab_mode) {
    // planes, not pairs, are dealt: every rank reduces all pairs over its planes.  The shot
    // noise stays with the last rank, or -- distributed mesh: every rank takes part in
    // xi(r) -- is dealt in blocks of entries.
    const int R = params.part_count, r = params.part_rank;
    share.pairs.assign(dv.dim, slab_nx > 0 ? 1 : 0);
    share.shot.assign(dv.dim, 0);
    // (compact blocks of the pair matrix: a rank evaluates j_l for the few wavenumbers of
    // its block only; the partition is remembered per binning)
    std::vector<int> shot_owner;
    if (dist_mesh) shot_owner = partition_owners(dv.row, dv.col, R, params.ell1 == params.ell2, 0.);
    for (int i = 0; i < dv.dim; i++) share.shot[i] = dist_mesh ? (shot_owner[i] == r) : (r == R - 1);
    share.any_pairs = slab_nx > 0;
    share.any_shot = dist_mesh || r == R - 1;
  } else {
    share = bispec_share(params, dv, shot_cost_in_fields);
  }
  const std::vector<char>& active = share.pairs;
  const std::vector<char>& shot_active = share.shot;

  trvb_ctx* c = eng.ctx();

  const bool coarse = nsub[0] != params.ngrid[0];
  std::shared_ptr<trvb_ctx> sub_holder;
  trvb_ctx* sub = c;
  if (coarse) {
    trvb_ctx* raw = nullptr;
    dev::check(trvb_subgrid_create(c, &raw, nsub), "trvb_subgrid_create");
    sub_holder.reset(raw, [](trvb_ctx* p) { trvb_ctx_destroy(p); });
    sub = raw;
  }

  // One GPU holding the whole job (box, throughput mode, real shell fields on a true
  // sub-grid): the mesh phase with the x passes of both full-grid transforms fused around
  // the shot-noise spectrum (trvb_box_fields_fused); xi(r) comes out of the same call.
  // TRV_NO_FUSED_X=1: 3-D cuFFT transforms + separate spectrum kernel.
  bool fused_x = !survey && !dist_mesh && coarse && !params.deterministic
    && share.any_pairs && share.any_shot && params.ell1 == 0 && params.ell2 == 0
    && params.ELL == 0 && trvb_box_fields_fused_supported(c);
  {
    const char* env = std::getenv("TRV_NO_FUSED_X");
    if (env != nullptr && env[0] == '1') fused_x = false;
  }
  dev::Mesh xi_fused;

  // Common fields: delta n_00(k) and N_00(k).
  dev::Mesh dn_00 = dist_mesh ? eng.density_fluctuation_dist(comm, sub)
    : fused_x ? eng.density_fluctuation_fused(sub, xi_fused)
    : eng.density_fluctuation(0, 0);
  dev::Mesh N_00_own;
  if (survey && share.any_shot) N_00_own = eng.quadratic_field(0, 0);   // shot noise only
  const trvb_mesh N_00 = (survey && share.any_shot)
    ? N_00_own.view() : eng.quadratic_view_of_fluctuation(dn_00);
  dev::profile_mark(c, "fields_00");

  trvm::SphericalBesselCalculator sj_a(params.ell1), sj_b(params.ell2);
  eng.upload_sjl(sj_a);
  if (params.ell2 != params.ell1) eng.upload_sjl(sj_b);

  // Shell statistics (k_eff, nmodes) for every bin in one pass; they do not
  // depend on (l, m) (S/field.cpp:1815-1847, 1905).
  // They depend on the grid and the bin edges only, not on the catalogue: remembered per
  // (box, mesh, edges), which also spares every later call a kernel and a host synchronisation
  // in the middle of its stream of launches.
  std::vector<long long> nmodes(nb);
  std::vector<double> ksum(nb), keff(nb);
  {
    typedef std::tuple<std::vector<double>, std::vector<double> > StatsKey;
    static std::mutex stats_mutex;
    static std::map<StatsKey, std::pair<std::vector<long long>, std::vector<double> > > stats_cache;
    std::vector<double> grid(params.boxsize, params.boxsize + 3);
    for (int ax = 0; ax < 3; ax++) grid.push_back(double(params.ngrid[ax]));
    const StatsKey key(grid, kbinning.bin_edges);
    bool hit = false;
    {
      std::lock_guard<std::mutex> lock(stats_mutex);
      auto it = stats_cache.find(key);
      if (it != stats_cache.end()) { nmodes = it->second.first; ksum = it->second.second; hit = true; }
    }
    if (!hit) {
      dev::check(trvb_shell_stats(c, kbinning.bin_edges.data(), nb, 0, nmodes.data(),
                                  ksum.data()), "trvb_shell_stats");
      std::lock_guard<std::mutex> lock(stats_mutex);
      if (stats_cache.size() >= 64) stats_cache.clear();
      stats_cache[key] = std::make_pair(nmodes, ksum);
    }
  }
  for (int b = 0; b < nb; b++) keff[b] = ksum[b] / double(nmodes[b]);
  dev::profile_mark(c, "shell_stats");

  const double nsub_mesh = double(nsub[0]) * nsub[1] * nsub[2];
  // vol_cell * sum over the n^3 mesh == (V / ns^3) * sum over the sub-grid.
  const double vol_cell_sub = eng.vol() / nsub_mesh;

  std::vector<cdouble> bk_dv(dv.dim, 0.), sn_dv(dv.dim, 0.);
  std::vector<double> k1eff(dv.dim), k2eff(dv.dim);
  for (int i = 0; i < dv.dim; i++) { k1eff[i] = keff[dv.row[i]]; k2eff[i] = keff[dv.col[i]]; }

  // Slab mode: a context whose "grid" is this rank's block of x-planes (allocation sizes
  // and the cell count of the pair reduction).
  trvb_ctx* slab_ctx = nullptr;
  if (slab_mode && slab_nx > 0) {
    const int dims[3] = {slab_nx, nsub[1], nsub[2]};
    dev::check(trvb_subgrid_create(c, &slab_ctx, dims), "trvb_subgrid_create (slab)");
  }
  trvb_ctx* pair_grid = slab_ctx ? slab_ctx : sub;
  // Real shell fields on a true sub-grid take the pruned per-axis transform (throughput
  // mode; the deterministic mode keeps the dense batched 3-D transform, whose arithmetic
  // per shell does not depend on how the shells are grouped).  TRV_NO_PRUNE=1: dense.
  bool pruned = coarse && !params.deterministic;
  {
    const char* env = std::getenv("TRV_NO_PRUNE");
    if (env != nullptr && env[0] == '1') pruned = false;
  }
  dev::Mesh xi;             // shot-noise mesh, reused across terms in a box
  std::unique_ptr<PlaneBlock> xi_planes;   // distributed mesh: this rank's planes of it
  const long long params_ndata = eng.ndata();
  dev::Mesh G;              // G_LM(x) on the sub-grid
  int G_M = 0; bool have_G = false, have_xi = false, G_on_slab = false;
  if (fused_x) { xi = std::move(xi_fused); have_xi = true; }
  dev::Mesh dn_LM;          // survey: delta n_LM(k) of the current M
  dev::Mesh N_LM;
  cdouble Sbar_LM = 0.;
  int cached_M = 0; bool have_LM = false;

  // All shells of one (ell, m) in a single sparse pass + one batched IFFT.
  auto shell_fields = [&](const dev::Mesh& src, int ell, int m, const std::vector<int>& bins,
                          int layout, Slab& dst) {
    std::vector<double> klo, khi, amp;
    for (int b : bins) {
      klo.push_back(kbinning.bin_edges[b]);
      khi.push_back(kbinning.bin_edges[b + 1]);
      amp.push_back(1. / double(nmodes[b]));   // F /= nmodes, S/field.cpp:1900-1905
    }
    if (slab_ctx) {
      dev::check(trvb_shell_slab_batch(c, sub, src.view(), ell, m, klo.data(), khi.data(),
                                       amp.data(), (int)bins.size(), slab_x0, slab_nx, dst.data()),
                 "trvb_shell_slab_batch");
    } else if (pruned && layout == TRVB_REAL) {
      // one GPU: the same pruned per-axis transform over all the planes
      dev::check(trvb_shell_slab_batch(c, sub, src.view(), ell, m, klo.data(), khi.data(),
                                       amp.data(), (int)bins.size(), 0, nsub[0], dst.data()),
                 "trvb_shell_slab_batch");
    } else {
      dev::check(trvb_shell_ifft_batch(c, sub, src.view(), ell, m, klo.data(), khi.data(),
                                       amp.data(), (int)bins.size(), dst.data(), layout),
                 "trvb_shell_ifft_batch");
    }
    trvs::count_ifft += (int)bins.size();
  };
  // Shell fields depend on (ell, m, bin) only: terms that share a side (e.g.
  // the (l2, m2) = (0, 0) side of every B_202 term) reuse its slab while
  // device memory is plentiful.
  typedef std::tuple<int, int, int, std::vector<int> > SlabKey;
  std::map<SlabKey, std::shared_ptr<Slab> > slab_cache;
  auto shell_slab = [&](int ell, int m, const std::vector<int>& bins, int layout) {
    const SlabKey key(ell, m, layout, bins);
    auto hit = slab_cache.find(key);
    if (hit != slab_cache.end()) return hit->second;
    auto slab = std::make_shared<Slab>(eng.shared(), pair_grid, layout, (int)bins.size());
    shell_fields(dn_00, ell, m, bins, layout, *slab);
    const size_t bytes = trvb_mesh_bytes(pair_grid, layout) * bins.size();
    if (terms.size() > 1 && eng.mesh_capacity(bytes) >= 6) slab_cache[key] = slab;
    return slab;
  };
  // A shell field is real when the filtered spectrum is Hermitian.
  auto shell_is_real = [&](const dev::Mesh& src, int ell, int m) {
    return src.layout() == TRVB_HALF && m == 0 && ell % 2 == 0;
  };

  // With TRV_OVERLAP=1 the pair branch lives on the sub-grid's own stream: xi(x)
  // (spectrum + full-grid inverse FFT, no host synchronisation) is then enqueued BEFORE
  // the pair branch and may run beside it.  On one stream the order is immaterial.
  const bool overlap = coarse && share.any_pairs && share.any_shot && !dev::profile_enabled();

  for (const Term& t : terms) {
    // ---- fields that depend on (L, M) only -----------------------------
    if (survey && !(have_LM && cached_M == t.M)) {
      dn_LM = eng.density_fluctuation(params.ELL, t.M);
      if (share.any_shot) {   // N_LM and the amplitude feed the shot-noise branch only
        N_LM = eng.quadratic_field(params.ELL, t.M);
        Sbar_LM = eng.shotnoise_amp(params.ELL, t.M);
      }
      cached_M = t.M; have_LM = true; have_G = false; have_xi = false;
      dev::profile_mark(c, "fields_LM");
    }
    if (!survey && !have_LM) {
      Sbar_LM = eng.shotnoise_amp(0, 0);
      have_LM = true;
    }
    const dev::Mesh& dn_LM_ref = survey ? dn_LM : dn_00;
    const trvb_mesh N_LM_ref = (survey && share.any_shot) ? N_LM.view() : N_00;

    // S|{i = j != k}: one xi mesh per (L, M), all pairs in one pass later.
    auto ensure_xi = [&]() {
      if (have_xi) return;
      const double S[2] = {Sbar_LM.real(), Sbar_LM.imag()};
      if (dist_mesh) {   // this rank's planes of xi(r) only
        int x0 = 0, nx = 0;
        dev::check(trvb_dmesh_planes(eng.dmesh(), &x0, &nx), "trvb_dmesh_planes");
        xi_planes.reset(new PlaneBlock(eng.shared(),
          sizeof(double) * (size_t)nx * (size_t)params.ngrid[1] * (size_t)params.ngrid[2]));
        dev::check(trvb_dmesh_shot_xi(eng.dmesh(), 0., double(params_ndata), S,
                                      static_cast<double*>(xi_planes->data())), "trvb_dmesh_shot_xi");
        trvs::count_ifft += 1;
        have_xi = true;
        dev::profile_mark(c, "shot_xi");
        return;
      }
      // Spectra of two real fields and a real amplitude: xi(x) is real.
      const bool real_xi = dn_LM_ref.layout() == TRVB_HALF && N_00.layout == TRVB_HALF
        && S[1] == 0.;
      xi = dev::Mesh(eng.shared(), c, real_xi ? TRVB_REAL : TRVB_COMPLEX);
      dev::check(trvb_shot_xi(c, dn_LM_ref.view(), N_00, S, /*interlaced=*/0, xi.view()), "trvb_shot_xi");
      trvs::count_ifft += 1;
      have_xi = true;
      dev::profile_mark(c, "shot_xi");
    };
    // The sub-grid stream may start once the Fourier meshes above are complete.
    dev::check(trvb_ctx_fork(c, sub), "trvb_ctx_fork");
    if (overlap) ensure_xi();

    if (share.any_pairs) {
      // ---- raw bispectrum --------------------------------------------------
      // Real arithmetic throughout when G and both shell fields are real.
      const bool real_path = dn_LM_ref.layout() == TRVB_HALF
        && shell_is_real(dn_00, params.ell1, t.m1) && shell_is_real(dn_00, params.ell2, t.m2);
      const int layout = real_path ? TRVB_REAL : TRVB_COMPLEX;
      if (!(have_G && G_M == t.M && G.layout() == layout)) {
        // G_LM(x) = IFFT[delta n_LM(k) / W(k)] / V (S/threept.cpp:452-459).
        // (slab mode too: G holds every mode the sub-grid represents -- twice the shells'
        // cut-off per axis -- so the pruned x-DFT does not pay for it; the whole G is
        // transformed once, 2 % of the job, and the slab reads its own planes of it)
        if (slab_ctx && layout == TRVB_REAL) {
          // x-slabs: only this rank's planes of G, through the same per-axis transform as
          // the shell fields (an "unbounded shell": the x pass runs on every column of the
          // sub-grid's spectrum, the y and z passes on the slab's planes only).
          G = dev::Mesh(eng.shared(), slab_ctx, layout);
          const double all = -1., amp_G = 1. / eng.vol();
          dev::check(trvb_shell_slab_batch(c, sub, dn_LM_ref.view(), 0, 0, &all, &all, &amp_G, 1,
                                           slab_x0, slab_nx, G.data()), "trvb_shell_slab_batch (G)");
          G_on_slab = true;
        } else {
          G = dev::Mesh(eng.shared(), sub, layout);
          dev::check(trvb_shell_ifft(c, sub, dn_LM_ref.view(), 0, 0, -1., -1., 1. / eng.vol(),
                                     G.view()), "trvb_shell_ifft (G)");
          G_on_slab = false;
        }
        trvs::count_ifft += 1;
        G_M = t.M; have_G = true;
        dev::profile_mark(c, "G_field");
      }
      const bool same_fields = (params.ell1 == params.ell2 && t.m1 == t.m2);
      // (l, -m) fields are (-1)^(l+m) conj of the (l, m) ones: one set of transforms.
      const bool mirror = layout == TRVB_COMPLEX
        && mirror_harmonics(params, t, dn_00, kbinning.bin_edges.back());
      const int m_b = mirror ? t.m1 : t.m2;
      std::vector<cdouble> bk_comp;
      trvb_mesh G_pairs = G.view();
      if (slab_ctx && !G_on_slab) {   // REAL layout, x slowest: the slab's planes are contiguous
        G_pairs.data = static_cast<char*>(G_pairs.data)
          + sizeof(double) * (size_t)slab_x0 * (size_t)nsub[1] * (size_t)nsub[2];
      }
      reduce_pairs(
        eng, pair_grid, dv, active, same_fields || mirror, G_pairs,
        [&](const std::vector<int>& bins) { return shell_slab(params.ell1, t.m1, bins, layout); },
        [&](const std::vector<int>& bins) { return shell_slab(params.ell2, m_b, bins, layout); },
        bk_comp, mirror);
      const double sign_b = (mirror && ((params.ell2 + t.m1) % 2 != 0)) ? -1. : 1.;
      for (int i = 0; i < dv.dim; i++) {
        if (!active[i]) continue;
        const cdouble comp = sign_b * bk_comp[i];
        bk_dv[i] += t.coupling * vol_cell_sub * (comp + t.factor_mirror * std::conj(comp));
      }
      dev::profile_mark(c, "shells_and_pairs");
    }

    if (share.any_shot) {
      // ---- shot noise ------------------------------------------------------
      if (params.ell1 == 0 && params.ell2 == 0) {   // S|{i = j = k}
        const cdouble S_ijk = t.coupling * Sbar_LM;
        for (int i = 0; i < dv.dim; i++) if (shot_active[i]) sn_dv[i] += S_ijk;
      }
      // The binned statistics depend on (ell, m) only: B_000-like cases reuse
      // one evaluation for both S|{i != j = k} and S|{j != i = k}.
      std::vector<double> pk, sn;
      int binned_ell = -1, binned_m = 0;
      auto binned_term = [&](int ell, int m, bool by_row) {
        if (!(binned_ell == ell && binned_m == m)) {
          std::vector<long long> nm(nb);
          std::vector<double> kk(nb);
          pk.assign(2 * nb, 0.); sn.assign(2 * nb, 0.);
          const double S[2] = {Sbar_LM.real(), Sbar_LM.imag()};
          // Distributed run on a large low-|k| cube: every rank bins a contiguous range of
          // shells and the per-bin means are summed over the ranks (zeros elsewhere).
          double cube_modes = 0.5;
          for (int ax = 0; ax < 3; ax++) {
            cube_modes *= 2. * kbinning.bin_edges.back() * params.boxsize[ax] / (2. * M_PI) + 1.;
          }
          if (dist_mesh && cube_modes > 4.e6 && nb >= params.part_count) {
            const int R = params.part_count, r = params.part_rank;
            const int b0 = static_cast<int>((long long)nb * r / R);
            const int b1 = static_cast<int>((long long)nb * (r + 1) / R);
            dev::check(trvb_twopt_fourier(c, dn_00.view(), N_LM_ref, S, ell, m, /*interlaced=*/0,
                                          kbinning.bin_edges.data() + b0,
                                          kbinning.bin_centres.data() + b0, b1 - b0,
                                          nm.data() + b0, kk.data() + b0, pk.data() + 2 * b0,
                                          sn.data() + 2 * b0), "trvb_twopt_fourier");
            std::vector<double> buf(pk);
            buf.insert(buf.end(), sn.begin(), sn.end());
            dev::allreduce(c, buf.data(), (long long)buf.size());
            std::copy(buf.begin(), buf.begin() + 2 * nb, pk.begin());
            std::copy(buf.begin() + 2 * nb, buf.end(), sn.begin());
          } else {
            dev::check(trvb_twopt_fourier(c, dn_00.view(), N_LM_ref, S, ell, m, /*interlaced=*/0,
                                          kbinning.bin_edges.data(), kbinning.bin_centres.data(),
                                          nb, nm.data(), kk.data(), pk.data(), sn.data()),
                       "trvb_twopt_fourier");
          }
          binned_ell = ell; binned_m = m;
        }
        for (int i = 0; i < dv.dim; i++) {
          if (!shot_active[i]) continue;
          const int b = by_row ? dv.row[i] : dv.col[i];
          const cdouble S_b = t.coupling * (cdouble(pk[2*b], pk[2*b+1]) - cdouble(sn[2*b], sn[2*b+1]));
          sn_dv[i] += S_b + t.factor_mirror * std::conj(S_b);
        }
      };
      if (params.ell2 == 0) binned_term(params.ell1, t.m1, true);    // S|{i != j = k}
      if (params.ell1 == 0) binned_term(params.ell2, t.m2, false);   // S|{j != i = k}
      dev::profile_mark(c, "shot_binned");

      ensure_xi();
      {
        std::vector<double> ka, kb; std::vector<int> where;
        for (int i = 0; i < dv.dim; i++) {
          if (!shot_active[i]) continue;
          ka.push_back(k1eff[i]); kb.push_back(k2eff[i]); where.push_back(i);
        }
        if (dist_mesh && where.empty()) {   // the histogram sum is collective: take part
          ka.push_back(keff[0]); kb.push_back(keff[0]);
        }
        if (!ka.empty()) {
          std::vector<double> S(2 * ka.size());
          if (dist_mesh) {
            int x0 = 0, nx = 0;
            dev::check(trvb_dmesh_planes(eng.dmesh(), &x0, &nx), "trvb_dmesh_planes");
            dev::check(trvb_shot_bispec_reduce_slab(
              c, static_cast<const double*>(xi_planes->data()), x0, nx, comm, params.ell1, t.m1,
              params.ell2, t.m2, ka.data(), kb.data(), (int)ka.size(), S.data()),
              "trvb_shot_bispec_reduce_slab");
          } else {
            dev::check(trvb_shot_bispec_reduce(c, xi.view(), params.ell1, t.m1, params.ell2, t.m2,
                                               ka.data(), kb.data(), (int)ka.size(), S.data()),
                       "trvb_shot_bispec_reduce");
          }
          for (size_t p = 0; p < where.size(); p++) {
            const cdouble S_ij_k = t.coupling * cdouble(S[2*p], S[2*p+1]);
            sn_dv[where[p]] += factor_phase * (
              S_ij_k + t.factor_mirror_w3j * std::conj(S_ij_k));
          }
        }
      }
    }
    dev::profile_mark(c, "shot_reduce");
    dev::check(trvb_ctx_join(c, sub), "trvb_ctx_join");
    if (trvs::currTask == 0) {
      trvs::logger.stat("Bispectrum term computed at orders (m1, m2, M) = +/-(%d, %d, %d).",
                        t.m1, t.m2, t.M);
    }
  }

  // Bins without modes: the reference divides the shell field and its effective
  // wavenumber by a zero mode count (S/field.cpp:1895-1905), so every entry that
  // touches such a bin comes out as NaN there -- reproduced (by the entry's owner,
  // so that the sum over ranks stays NaN) rather than returned as a silent zero.
  {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int i = 0; i < dv.dim; i++) {
      const bool empty_row = nmodes[dv.row[i]] == 0, empty_col = nmodes[dv.col[i]] == 0;
      if (empty_row) k1eff[i] = nan;
      if (empty_col) k2eff[i] = nan;
      if (!(empty_row || empty_col)) continue;
      if (active[i]) bk_dv[i] = cdouble(nan, nan);
      if (shot_active[i]) sn_dv[i] = cdouble(nan, nan);
    }
  }

  trv::BispecMeasurements out;
  out.dim = dv.dim;
  for (int i = 0; i < dv.dim; i++) {
    out.k1_bin.push_back(kbinning.bin_centres[dv.row[i]]);
    out.k2_bin.push_back(kbinning.bin_centres[dv.col[i]]);
    out.k1_eff.push_back(k1eff[i]);
    out.k2_eff.push_back(k2eff[i]);
    out.nmodes_1.push_back(static_cast<int>(nmodes[dv.row[i]]));
    out.nmodes_2.push_back(static_cast<int>(nmodes[dv.col[i]]));
    out.bk_raw.push_back(norm_factor * bk_dv[i]);
    out.bk_shot.push_back(norm_factor * sn_dv[i]);
  }
  return out;
}

// =====================================================================
// Three-point correlation function
// =====================================================================

trv::ThreePCFMeasurements threepcf_impl(
  Engine& eng, trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor,
  bool wide_angle = false
) {
  const bool survey = eng.survey();
  const cdouble factor_phase = std::pow(trvm::M_I, params.ell1 + params.ell2);
  const double factor_parity = std::pow(-1., params.ell1 + params.ell2);
  const int nb = rbinning.num_bins;
  const DataVector dv = make_data_vector(params, nb);
  const std::vector<char> active = active_entries(params, dv);

  trvb_ctx* c = eng.ctx();
  const double vol_cell = eng.vol_cell();

  dev::Mesh dn_00 = eng.density_fluctuation(0, 0);
  dev::Mesh N_00_own;
  if (survey) N_00_own = eng.quadratic_field(0, 0);
  const trvb_mesh N_00 = survey ? N_00_own.view() : eng.quadratic_view_of_fluctuation(dn_00);
  dev::profile_mark(c, "fields_00");

  trvm::SphericalBesselCalculator sj_a(params.ell1), sj_b(params.ell2);
  eng.upload_sjl(sj_a);
  if (params.ell2 != params.ell1) eng.upload_sjl(sj_b);

  std::vector<cdouble> zeta_dv(dv.dim, 0.), sn_dv(dv.dim, 0.);
  std::vector<double> reff(nb, 0.);
  std::vector<long long> npairs(nb, 0);

  const std::vector<Term> terms = enumerate_terms(params, survey);
  dev::Mesh dn_LM, xi, G;
  cdouble Sbar_LM = 0.;
  int cached_M = 0; bool have_LM = false, have_xi = false, have_G = false;
  int count_terms = 0;

  for (const Term& t : terms) {
    if (survey && !(have_LM && cached_M == t.M)) {
      dn_LM = eng.density_fluctuation(params.ELL, t.M);
      Sbar_LM = eng.shotnoise_amp(params.ELL, t.M);
      cached_M = t.M; have_LM = true; have_xi = false; have_G = false;
      dev::profile_mark(c, "fields_LM");
    }
    if (!survey && !have_LM) { Sbar_LM = eng.shotnoise_amp(0, 0); have_LM = true; }
    const dev::Mesh& dn_LM_ref = survey ? dn_LM : dn_00;

    // ---- shot noise first: it also yields r_eff and npairs ---------------
    if (!have_xi) {
      const double S[2] = {Sbar_LM.real(), Sbar_LM.imag()};
      const bool real_xi = dn_LM_ref.layout() == TRVB_HALF && N_00.layout == TRVB_HALF
        && S[1] == 0.;
      xi = dev::Mesh(eng.shared(), c, real_xi ? TRVB_REAL : TRVB_COMPLEX);
      dev::check(trvb_shot_xi(c, dn_LM_ref.view(), N_00, S, /*interlaced=*/0, xi.view()), "trvb_shot_xi");
      trvs::count_ifft += 1;
      have_xi = true;
      dev::profile_mark(c, "shot_xi");
    }
    std::vector<long long> np(nb);
    std::vector<double> rr(nb), xib(2 * nb);
    dev::check(trvb_shot_3pcf_bin(c, xi.view(), params.ell1, t.m1, params.ell2, t.m2,
                                  rbinning.bin_edges.data(), rbinning.bin_centres.data(), nb,
                                  factor_parity, np.data(), rr.data(), xib.data()),
               "trvb_shot_3pcf_bin");
    // Kronecker delta of eq. (51): shot noise on equal bins only
    // (S/threept.cpp:2388-2437).
    for (int i = 0; i < dv.dim; i++) {
      if (!active[i] || dv.row[i] != dv.col[i]) continue;
      if (params.shape == "off-diag" && params.idx_bin != 0) continue;
      const int b = dv.row[i];
      const cdouble x(xib[2*b], xib[2*b+1]);
      sn_dv[i] += t.coupling * factor_parity * (x + t.factor_mirror_w3j * std::conj(x));
    }
    if (count_terms == 0) { reff = rr; npairs = np; }
    dev::profile_mark(c, "shot_3pcf_bin");

    // ---- raw 3PCF ---------------------------------------------------------
    if (!have_G) {
      G = dev::Mesh(eng.shared(), c, TRVB_COMPLEX);
      dev::check(trvb_shell_ifft(c, c, dn_LM_ref.view(), 0, 0, -1., -1., 1. / eng.vol(),
                                 G.view()), "trvb_shell_ifft (G)");
      trvs::count_ifft += 1;
      if (wide_angle) {   // window only: G_LM(x) r^{-i-j} (S/threept.cpp:2978-2980)
        dev::check(trvb_mesh_pow_law(c, G.view(), params.i_wa + params.j_wa),
                   "trvb_mesh_pow_law");
      }
      have_G = true;
    }
    auto sjl_fields = [&](int ell, int m, const std::vector<int>& bins) {
      auto slab = std::make_shared<Slab>(eng.shared(), c, TRVB_COMPLEX, (int)bins.size());
      std::vector<double> radii;
      for (int b : bins) radii.push_back(reff[b]);
      dev::check(trvb_sjl_ifft_batch(c, dn_00.view(), ell, m, radii.data(), 1. / eng.vol(),
                                     (int)bins.size(), slab->data()), "trvb_sjl_ifft_batch");
      trvs::count_ifft += (int)bins.size();
      return slab;
    };
    const bool same_fields = (params.ell1 == params.ell2 && t.m1 == t.m2);
    // (No mirror-harmonic shortcut here: j_l(kr) weights the Nyquist planes too.)
    std::vector<cdouble> zeta_comp;
    reduce_pairs(
      eng, c, dv, active, same_fields, G.view(),
      [&](const std::vector<int>& bins) { return sjl_fields(params.ell1, t.m1, bins); },
      [&](const std::vector<int>& bins) { return sjl_fields(params.ell2, t.m2, bins); },
      zeta_comp);
    for (int i = 0; i < dv.dim; i++) {
      if (!active[i]) continue;
      zeta_dv[i] += t.coupling * vol_cell * factor_phase * (
        zeta_comp[i] + t.factor_mirror * std::conj(zeta_comp[i]));
    }
    dev::profile_mark(c, "sjl_fields_and_pairs");
    count_terms++;
    if (trvs::currTask == 0) {
      trvs::logger.stat(
        "Three-point correlation function term computed at orders "
        "(m1, m2, M) = +/-(%d, %d, %d).", t.m1, t.m2, t.M);
    }
  }

  trv::ThreePCFMeasurements out;
  out.dim = dv.dim;
  for (int i = 0; i < dv.dim; i++) {
    out.r1_bin.push_back(rbinning.bin_centres[dv.row[i]]);
    out.r2_bin.push_back(rbinning.bin_centres[dv.col[i]]);
    out.r1_eff.push_back(reff[dv.row[i]]);
    out.r2_eff.push_back(reff[dv.col[i]]);
    out.npairs_1.push_back(static_cast<int>(npairs[dv.row[i]]));
    out.npairs_2.push_back(static_cast<int>(npairs[dv.col[i]]));
    out.zeta_raw.push_back(norm_factor * zeta_dv[i]);
    out.zeta_shot.push_back(norm_factor * sn_dv[i]);
  }
  return out;
}

}  // namespace

// =====================================================================
// Public entry points
// =====================================================================

namespace {

/// Runs one estimator call over the GPUs.
///  * Caller-driven partition (params.part_count > 1, one process per GPU): the call
///    computes this rank's share; when the process communicator spans part_count ranks the
///    raw and shot-noise vectors are summed over NCCL, so every rank returns the complete
///    measurement (every entry comes from exactly one rank: the sum adds zeros).
///  * Otherwise, when the process sees several usable GPUs (dev::multi_device_count), one
///    host thread per GPU computes a share with its own context and the shares are summed on
///    the host in rank order -- the single-process multi-GPU mode of the reference
///    (S/field.cpp:212-235, S/monitor.cpp:258-324) without its cuFFT-Xt slab transforms.
/// `run(p)` builds the engine on the calling thread's current device and returns the result.
template <class Result, class Run>
Result run_over_gpus(trv::ParameterSet& params, Run run,
                     std::vector<cdouble> Result::* raw, std::vector<cdouble> Result::* shot) {
  if (params.part_count > 1) {
    Result out = run(params);
    trvb_comm* comm = dev::process_comm();
    if (comm != nullptr && trvb_comm_size(comm) == params.part_count) {
      const size_t n = (out.*raw).size();
      std::vector<double> buf(4 * n);
      for (size_t i = 0; i < n; i++) {
        buf[2*i] = (out.*raw)[i].real(); buf[2*i + 1] = (out.*raw)[i].imag();
        buf[2*n + 2*i] = (out.*shot)[i].real(); buf[2*n + 2*i + 1] = (out.*shot)[i].imag();
      }
      dev::allreduce(dev::last_context(), buf.data(), (long long)buf.size());
      for (size_t i = 0; i < n; i++) {
        (out.*raw)[i] = cdouble(buf[2*i], buf[2*i + 1]);
        (out.*shot)[i] = cdouble(buf[2*n + 2*i], buf[2*n + 2*i + 1]);
      }
    }
    return out;
  }
  const int ndev = dev::multi_device_count(params);
  if (ndev <= 1) return run(params);

  std::vector<Result> parts(ndev);
  std::vector<std::string> errors(ndev);
  std::vector<std::thread> workers;
  for (int r = 0; r < ndev; r++) {
    workers.emplace_back([&, r]() {
      try {
        dev::check(trvb_set_current_device(r), "trvb_set_current_device");
        trv::ParameterSet p = params;
        p.part_rank = r; p.part_count = ndev;
        parts[r] = run(p);
      } catch (const std::exception& e) {
        errors[r] = e.what()[0] ? e.what() : "unknown error";
      }
    });
  }
  for (std::thread& w : workers) w.join();
  for (int r = 0; r < ndev; r++) {
    if (!errors[r].empty()) throw trvs::DeviceError("GPU %d: %s", r, errors[r].c_str());
  }
  Result out = std::move(parts[0]);
  for (int r = 1; r < ndev; r++) {
    for (size_t i = 0; i < (out.*raw).size(); i++) {
      (out.*raw)[i] += (parts[r].*raw)[i];
      (out.*shot)[i] += (parts[r].*shot)[i];
    }
  }
  return out;
}

}  // namespace

trv::BispecMeasurements compute_bispec(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& kbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat("Computing bispectrum from paired survey-type catalogues...");
  }
  validate_multipole_coupling(params);
  trv::BispecMeasurements out = run_over_gpus<trv::BispecMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, catalogue_data, &catalogue_rand, los_data, los_rand);
      return bispec_impl(eng, p, kbinning, norm_factor);
    }, &trv::BispecMeasurements::bk_raw, &trv::BispecMeasurements::bk_shot);
  if (trvs::currTask == 0) {
    trvs::logger.stat("... computed bispectrum from paired survey-type catalogues.");
  }
  return out;
}

trv::BispecMeasurements compute_bispec_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning kbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing bispectrum from a periodic-box simulation-type catalogue "
      "in the global plane-parallel approximation...");
  }
  validate_multipole_coupling(params);
  trv::BispecMeasurements out = run_over_gpus<trv::BispecMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, catalogue_data, nullptr, nullptr, nullptr);
      return bispec_impl(eng, p, kbinning, norm_factor);
    }, &trv::BispecMeasurements::bk_raw, &trv::BispecMeasurements::bk_shot);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed bispectrum from a periodic-box simulation-type catalogue "
      "in the global plane-parallel approximation.");
  }
  return out;
}

trv::ThreePCFMeasurements compute_3pcf(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing three-point correlation function from paired survey-type catalogues...");
  }
  validate_multipole_coupling(params);
  trv::ThreePCFMeasurements out = run_over_gpus<trv::ThreePCFMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, catalogue_data, &catalogue_rand, los_data, los_rand);
      return threepcf_impl(eng, p, rbinning, norm_factor);
    }, &trv::ThreePCFMeasurements::zeta_raw, &trv::ThreePCFMeasurements::zeta_shot);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed three-point correlation function from paired survey-type catalogues.");
  }
  return out;
}

trv::ThreePCFMeasurements compute_3pcf_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing three-point correlation function from a periodic-box "
      "simulation-type catalogue in the global plane-parallel approximation...");
  }
  validate_multipole_coupling(params);
  trv::ThreePCFMeasurements out = run_over_gpus<trv::ThreePCFMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, catalogue_data, nullptr, nullptr, nullptr);
      return threepcf_impl(eng, p, rbinning, norm_factor);
    }, &trv::ThreePCFMeasurements::zeta_raw, &trv::ThreePCFMeasurements::zeta_shot);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed three-point correlation function from a periodic-box "
      "simulation-type catalogue in the global plane-parallel approximation.");
  }
  return out;
}

trv::ThreePCFWindowMeasurements compute_3pcf_window(
  ParticleCatalogue& catalogue_rand, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning,
  double alpha, double norm_factor, bool wide_angle
) {
  trvs::logger.reset_level(params.verbose);
  const char* tag = wide_angle ? "wide-angle corrections " : "";
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "Computing three-point correlation function window %sfrom random catalogue...", tag);
  }
  validate_multipole_coupling(params);
  trv::ThreePCFMeasurements res = run_over_gpus<trv::ThreePCFMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, catalogue_rand, los_rand, alpha);
      return threepcf_impl(eng, p, rbinning, norm_factor, wide_angle);
    }, &trv::ThreePCFMeasurements::zeta_raw, &trv::ThreePCFMeasurements::zeta_shot);
  trv::ThreePCFWindowMeasurements out;
  out.dim = res.dim;
  out.r1_bin = std::move(res.r1_bin); out.r2_bin = std::move(res.r2_bin);
  out.r1_eff = std::move(res.r1_eff); out.r2_eff = std::move(res.r2_eff);
  out.npairs_1 = std::move(res.npairs_1); out.npairs_2 = std::move(res.npairs_2);
  out.zeta_raw = std::move(res.zeta_raw); out.zeta_shot = std::move(res.zeta_shot);
  if (trvs::currTask == 0) {
    trvs::logger.stat(
      "... computed three-point correlation function window %sfrom random catalogue.", tag);
  }
  return out;
}

// ---------------------------------------------------------------------
// Array-level entry points (B200 build extension): the periodic-box
// estimators fed from coordinate arrays in host or DEVICE memory, skipping
// the AoS staging copy of ParticleCatalogue.  Same results as the
// catalogue-based overloads.
// ---------------------------------------------------------------------

trv::BispecMeasurements compute_bispec_in_gpp_box(
  long long nparticles, const double* x, const double* y, const double* z,
  bool on_device, trv::ParameterSet& params, trv::Binning kbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  validate_multipole_coupling(params);
  return run_over_gpus<trv::BispecMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, nparticles, x, y, z, on_device);
      return bispec_impl(eng, p, kbinning, norm_factor);
    }, &trv::BispecMeasurements::bk_raw, &trv::BispecMeasurements::bk_shot);
}

trv::ThreePCFMeasurements compute_3pcf_in_gpp_box(
  long long nparticles, const double* x, const double* y, const double* z,
  bool on_device, trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor
) {
  trvs::logger.reset_level(params.verbose);
  validate_multipole_coupling(params);
  return run_over_gpus<trv::ThreePCFMeasurements>(
    params, [&](trv::ParameterSet& p) {
      Engine eng(p, nparticles, x, y, z, on_device);
      return threepcf_impl(eng, p, rbinning, norm_factor);
    }, &trv::ThreePCFMeasurements::zeta_raw, &trv::ThreePCFMeasurements::zeta_shot);
}

}  // namespace trv
