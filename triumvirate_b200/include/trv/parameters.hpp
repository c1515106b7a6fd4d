// parameters.hpp -- trv::ParameterSet (mirror of I/parameters.hpp:58-258).
#ifndef TRV_B200_PARAMETERS_HPP_
#define TRV_B200_PARAMETERS_HPP_

#include <string>
#include <vector>

#include "monitor.hpp"

namespace trv {

class ParameterSet {
 public:
  // -- I/O (S/parameters.cpp:94-466, 1271-1380) --
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

  /// Read `key = value` lines of a parameter INI file ('#' starts a comment line), then
  /// validate(true) (S/parameters.cpp:94-464).  Returns validate()'s status.
  int read_from_file(char* parameter_filepath);

  /// Validate and derive parameters (S/parameters.cpp:466-1270).  `init`: first
  /// validation after reading a file -- directory and catalogue paths are joined.
  int validate(bool init = false);

  /// Write the parameters in use as `key = value` lines (S/parameters.cpp:1271-1372);
  /// without argument: `<measurement_dir>parameters_used<output_tag>`.
  int print_to_file(char* out_parameter_filepath);
  int print_to_file();

  /// Catalogue file paths after splitting at trv::sys::fn_delimiter and joining with
  /// `catalogue_dir` (filled by validate()).
  std::vector<std::string> data_catalogue_files;
  std::vector<std::string> rand_catalogue_files;
};

/// TRV_OVERRIDE_OUTPUT_TAG / _VERBOSE / _PROGBAR replace the corresponding parameters
/// (S/parameters.cpp:1382-1414; the FFTW overrides have no effect in this build).
void override_paramset_by_envvars(trv::ParameterSet& params);
/// boxsize = expand x coordinate spans (S/parameters.cpp:1416-1432).
void set_boxsize_from_expand(const double* spans, trv::ParameterSet& params);
/// ngrid from the Nyquist cut-off, rounded up to even (S/parameters.cpp:1434-1470).
void set_ngrid_from_cutoff(trv::ParameterSet& params);

}  // namespace trv

#endif  // TRV_B200_PARAMETERS_HPP_
