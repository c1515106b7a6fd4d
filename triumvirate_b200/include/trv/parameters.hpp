// parameters.hpp -- trv::ParameterSet (mirror of I/parameters.hpp:58-258).
#ifndef TRV_B200_PARAMETERS_HPP_
#define TRV_B200_PARAMETERS_HPP_

#include <string>
#include <vector>

#include "monitor.hpp"

namespace trv {

class ParameterSet {
 public:
  // -- I/O (kept for struct compatibility; file I/O is out of scope) --
  std::string catalogue_dir;
  std::string measurement_dir;
  std::string data_catalogue_file;
  std::string rand_catalogue_file;
  std::string catalogue_columns;
  std::string catalogue_dataset;
  std::string output_tag;

  // -- Mesh sampling --
  double boxsize[3] = {0., 0., 0.};
  int ngrid[3] = {0, 0, 0};
  double expand = 1.;
  double cutoff_nyq = 0.;
  std::string alignment = "centre";
  std::string padscale = "box";
  double padfactor = 0.;
  std::string assignment = "tsc";
  std::string interlace = "false";
  double volume = 0.;        // derived
  long long nmesh = 0;       // derived
  int assignment_order = 0;  // derived

  // -- Measurement --
  std::string catalogue_type;
  std::string statistic_type;
  std::string npoint;  // derived: "2pt" | "3pt" | "none"
  std::string space;   // derived: "fourier" | "config"
  int ell1 = 0;
  int ell2 = 0;
  int ELL = 0;
  int i_wa = 0;
  int j_wa = 0;
  std::string form = "diag";
  std::string norm_convention = "particle";
  std::string shape = "diag";  // derived
  std::string binning = "lin";
  double bin_min = 0.;
  double bin_max = 0.;
  int num_bins = 0;
  int idx_bin = 0;

  // -- Misc --
  std::string fftw_scheme = "measure";     // accepted and ignored (cuFFT)
  unsigned fftw_planner_flag = 0;
  std::string use_fftw_wisdom = "false";   // accepted and ignored
  std::string fftw_wisdom_file_f;
  std::string fftw_wisdom_file_b;
  std::string save_binned_vectors = "false";
  int verbose = 20;
  std::string progbar = "false";

  // -- B200 build extensions --
  /// 1 selects the deterministic (bit-reproducible) mesh assignment; also
  /// taken from the environment variable TRV_DETERMINISTIC at validate().
  int deterministic = 0;
  /// Work partition for multi-GPU runs: this process computes the data-vector
  /// entries that trv::partition_owners() gives to part_rank (compact blocks of
  /// the bin-pair matrix) and returns zeros elsewhere, so a sum over ranks (one
  /// small all-reduce) gives the full result.
  int part_rank = 0;
  int part_count = 1;

  ParameterSet() = default;
  ~ParameterSet() = default;

  /// Validate and derive parameters (S/parameters.cpp:466-1270).
  int validate(bool init = false);
};

}  // namespace trv

#endif  // TRV_B200_PARAMETERS_HPP_
