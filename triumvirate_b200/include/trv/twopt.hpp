// twopt.hpp -- two-point estimators (drop-in for I/twopt.hpp:63-325).
//
// Same names, argument meaning and exceptions as the reference; the work runs
// on the device through include/trvb.h.  Interlacing is honoured here (the only
// estimators for which validate() keeps it, S/parameters.cpp:1240-1249).
#ifndef TRV_B200_TWOPT_HPP_
#define TRV_B200_TWOPT_HPP_

#include <complex>
#include <string>

#include "dataobjs.hpp"
#include "field.hpp"
#include "parameters.hpp"
#include "particles.hpp"

namespace trv {

/// (2l+1)(2L+1) (l 0 L; 0 0 0)(l 0 L; m 0 M), S/twopt.cpp:45-49.
double calc_coupling_coeff_2pt(int ell, int ELL, int m, int M);

/// 1 / (alpha sum ws nz wc^2), S/twopt.cpp:56-96.
double calc_powspec_normalisation_from_particles(ParticleCatalogue& particles, double alpha);

/// Grid-based power-law normalisation of order 2, S/twopt.cpp:98-109.
double calc_powspec_normalisation_from_mesh(
  ParticleCatalogue& particles, trv::ParameterSet& params, double alpha);

/// Mixed-mesh normalisation 1 / (alpha dV sum_x n_data n_rand), S/twopt.cpp:111-228.
double calc_powspec_normalisation_from_meshes(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  trv::ParameterSet& params, double alpha);

/// The same on a dedicated normalisation mesh, S/twopt.cpp:230-263.
double calc_powspec_normalisation_from_meshes(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  trv::ParameterSet& params, double alpha,
  double padding, double cellsize, const std::string& assignment);

/// sum ws^2 wc^2, S/twopt.cpp:271-296.
double calc_powspec_shotnoise_from_particles(ParticleCatalogue& particles, double alpha);

/// sum_data y_lm w^2 + alpha^2 sum_rand y_lm w^2, S/twopt.cpp:298-350.
std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_powspec(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m);

/// alpha^2 sum y_lm w^2, S/twopt.cpp:352-380.
std::complex<double> calc_ylm_wgtd_shotnoise_amp_for_powspec(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m);

/// Survey-type power spectrum, local plane-parallel (S/twopt.cpp:388-496).
trv::PowspecMeasurements compute_powspec(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& kbinning, double norm_factor);

/// Survey-type two-point correlation function (S/twopt.cpp:498-607).
trv::TwoPCFMeasurements compute_corrfunc(
  ParticleCatalogue& catalogue_data, ParticleCatalogue& catalogue_rand,
  LineOfSight* los_data, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor);

/// Periodic-box power spectrum, global plane-parallel (S/twopt.cpp:609-701).
trv::PowspecMeasurements compute_powspec_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning kbinning, double norm_factor);

/// Periodic-box two-point correlation function (S/twopt.cpp:703-793).
trv::TwoPCFMeasurements compute_corrfunc_in_gpp_box(
  ParticleCatalogue& catalogue_data,
  trv::ParameterSet& params, trv::Binning& rbinning, double norm_factor);

/// Two-point correlation function window of a random catalogue
/// (S/twopt.cpp:795-901).
trv::TwoPCFWindowMeasurements compute_corrfunc_window(
  ParticleCatalogue& catalogue_rand, LineOfSight* los_rand,
  trv::ParameterSet& params, trv::Binning rbinning, double alpha, double norm_factor);

}  // namespace trv

#endif  // TRV_B200_TWOPT_HPP_
