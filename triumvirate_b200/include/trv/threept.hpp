// threept.hpp -- three-point estimators (mirror of I/threept.hpp:115-314).
#ifndef TRV_B200_THREEPT_HPP_
#define TRV_B200_THREEPT_HPP_

#include <complex>

#include "dataobjs.hpp"
#include "field.hpp"
#include "maths.hpp"
#include "monitor.hpp"
#include "parameters.hpp"
#include "particles.hpp"

namespace trv {

/// Spherical orders (m1, m2, M) of one term (S/threept.cpp:40-58).
struct SphericalOrderTriplet {
  int m1, m2, M;
  bool is_zeros() const { return m1 == 0 && m2 == 0 && M == 0; }
  bool is_inverse(const SphericalOrderTriplet& o) const {
    return m1 + o.m1 == 0 && m2 + o.m2 == 0 && M + o.M == 0;
  }
};

/// (2l1+1)(2l2+1)(2L+1) 3j(l1 l2 L; 0 0 0) 3j(l1 l2 L; m1 m2 M)
/// (S/threept.cpp:65-71).
double calc_coupling_coeff_3pt(int ell1, int ell2, int ELL, int m1, int m2, int M);

/// Throws if 3j(l1 l2 L; 0 0 0) vanishes (S/threept.cpp:73-89).
void validate_multipole_coupling(trv::ParameterSet& params);

/// 1 / (alpha sum ws nz^2 wc^3) (S/threept.cpp:96-136).
double calc_bispec_normalisation_from_particles(
  ParticleCatalogue& particles, double alpha = 1.);

/// 1 / (dV sum_x n_w(x)^3) / alpha^3 (S/threept.cpp:138-149).
double calc_bispec_normalisation_from_mesh(
  ParticleCatalogue& particles, trv::ParameterSet& params, double alpha = 1.);

/// sum_data y_LM w^3 + alpha^3 sum_rand y_LM w^3 (S/threept.cpp:156-208).
std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_bispec(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m);

/// alpha^3 sum y_LM w^3 (S/threept.cpp:210-236).
std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_bispec(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m);

/// Survey-type bispectrum, local plane-parallel (S/threept.cpp:248-1012).
trv::BispecMeasurements compute_bispec(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& kbinning, double norm_factor);

/// Survey-type 3PCF, local plane-parallel (S/threept.cpp:1014-1471).
trv::ThreePCFMeasurements compute_3pcf(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor);

/// Periodic-box bispectrum, global plane-parallel (S/threept.cpp:1473-2188).
trv::BispecMeasurements compute_bispec_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning kbinning, double norm_factor);

/// Periodic-box 3PCF, global plane-parallel (S/threept.cpp:2190-2619).
trv::ThreePCFMeasurements compute_3pcf_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor);

/// 3PCF window from a random catalogue, optionally with the wide-angle
/// power-law kernel r^{-i_wa-j_wa} on G_LM (S/threept.cpp:2621-3077).
trv::ThreePCFWindowMeasurements compute_3pcf_window(
  ParticleCatalogue& catalogue_rand, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning,
  double alpha, double norm_factor, bool wide_angle = false);

/// Array-level overloads (B200 build extension): periodic-box estimators
/// from coordinate arrays in host (`on_device` false) or device memory.
trv::BispecMeasurements compute_bispec_in_gpp_box(
  long long nparticles, const double* x, const double* y, const double* z,
  bool on_device, trv::ParameterSet& params, trv::Binning kbinning, double norm_factor);
trv::ThreePCFMeasurements compute_3pcf_in_gpp_box(
  long long nparticles, const double* x, const double* y, const double* z,
  bool on_device, trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor);

/// Multi-GPU work split (B200 build extension): owner rank in [0, world) of
/// every data-vector entry given its (row bin, column bin); the estimators
/// compute the entries owned by `ParameterSet::part_rank` and leave zeros elsewhere.
std::vector<int> partition_owners(
  const std::vector<int>& row, const std::vector<int>& col, int world, bool same_fields = true,
  double handicap_last = 0.);

/// The same for a parameter set's data-vector shape (after validate()).
std::vector<int> partition_owners(const trv::ParameterSet& params, int num_bins, int world);

}  // namespace trv

#endif  // TRV_B200_THREEPT_HPP_
