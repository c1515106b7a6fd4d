// io.hpp -- measurement output files (mirror of I/io.hpp:40-234; S/io.cpp:102-558).
//
// File format (what downstream readers parse): '#'-prefixed header lines, then a
// tab-separated table with "%.9e" coordinates, "%10d" counts and "% .9e" statistics.
#ifndef TRV_B200_IO_HPP_
#define TRV_B200_IO_HPP_

#include <cstdio>
#include <string>

#include "dataobjs.hpp"
#include "monitor.hpp"
#include "parameters.hpp"
#include "particles.hpp"

namespace trv {

namespace io {

const char comment_delimiter[] = "#";

/// Header for measurements from a pair of catalogues / a single catalogue: sources,
/// sizes, extents, box, mesh, assignment and the three normalisation factors.
void print_measurement_header_to_file(
  std::FILE* fileptr, trv::ParameterSet& params,
  trv::ParticleCatalogue& catalogue_data, trv::ParticleCatalogue& catalogue_rand,
  double norm_factor_part, double norm_factor_mesh, double norm_factor_meshes
);
void print_measurement_header_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ParticleCatalogue& catalogue,
  double norm_factor_part, double norm_factor_mesh, double norm_factor_meshes
);

void print_binned_vectors_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::BinnedVectors& binned_vectors
);

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::PowspecMeasurements& meas_powspec);
void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::TwoPCFMeasurements& meas_2pcf);
void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::TwoPCFWindowMeasurements& meas_2pcf_win);
void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::BispecMeasurements& meas_bispec);
void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ThreePCFMeasurements& meas_3pcf);
void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ThreePCFWindowMeasurements& meas_3pcf_win);

}  // namespace io
}  // namespace trv

#endif  // TRV_B200_IO_HPP_
