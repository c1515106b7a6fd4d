// dataobjs.hpp -- Binning, LineOfSight and measurement structs
// (mirror of I/dataobjs.hpp:58-282; S/dataobjs.cpp:134-249).
#ifndef TRV_B200_DATAOBJS_HPP_
#define TRV_B200_DATAOBJS_HPP_

#include <complex>
#include <string>
#include <vector>

#include "monitor.hpp"
#include "parameters.hpp"

namespace trv {

class Binning {
 public:
  std::string space;
  std::string scheme;
  double bin_min = 0.;
  double bin_max = 0.;
  int num_bins = 0;
  std::vector<double> bin_edges;
  std::vector<double> bin_centres;
  std::vector<double> bin_widths;

  explicit Binning(std::string space, std::string scheme);
  explicit Binning(trv::ParameterSet& params);

  void set_bins(double coord_min, double coord_max, int nbin);
  void set_bins();
  void set_bins(double boxsize_max, int ngrid_min);
  void set_bins(std::vector<double> bin_edges);

 private:
  int nbin_pad = 5;
  double dbin_pad_fourier = 1.e-3;
  double dbin_pad_config = 10.;
  void compute_binning();
};

/// Mesh vectors falling in each bin (I/dataobjs.hpp:166-175).
struct BinnedVectors {
  int count = 0;
  int num_bins = 0;
  std::vector<int> indices;
  std::vector<double> lower_edges;
  std::vector<double> upper_edges;
  std::vector<double> vecx;
  std::vector<double> vecy;
  std::vector<double> vecz;
};

struct LineOfSight {
  double pos[3];
};

struct BispecMeasurements {
  int dim = 0;
  std::vector<double> k1_bin;
  std::vector<double> k2_bin;
  std::vector<double> k1_eff;
  std::vector<double> k2_eff;
  std::vector<int> nmodes_1;
  std::vector<int> nmodes_2;
  std::vector< std::complex<double> > bk_raw;
  std::vector< std::complex<double> > bk_shot;
};

struct ThreePCFMeasurements {
  int dim = 0;
  std::vector<double> r1_bin;
  std::vector<double> r2_bin;
  std::vector<double> r1_eff;
  std::vector<double> r2_eff;
  std::vector<int> npairs_1;
  std::vector<int> npairs_2;
  std::vector< std::complex<double> > zeta_raw;
  std::vector< std::complex<double> > zeta_shot;
};

/// 3PCF window measurements (I/dataobjs.hpp:288-303); same members as
/// ThreePCFMeasurements.
struct ThreePCFWindowMeasurements {
  int dim = 0;
  std::vector<double> r1_bin;
  std::vector<double> r2_bin;
  std::vector<double> r1_eff;
  std::vector<double> r2_eff;
  std::vector<int> npairs_1;
  std::vector<int> npairs_2;
  std::vector< std::complex<double> > zeta_raw;
  std::vector< std::complex<double> > zeta_shot;
};

/// Power spectrum measurements (I/dataobjs.hpp:202-211).
struct PowspecMeasurements {
  int dim = 0;
  std::vector<double> kbin;
  std::vector<double> keff;
  std::vector<int> nmodes;
  std::vector< std::complex<double> > pk_raw;
  std::vector< std::complex<double> > pk_shot;
};

/// Two-point correlation function measurements (I/dataobjs.hpp:217-224).
struct TwoPCFMeasurements {
  int dim = 0;
  std::vector<double> rbin;
  std::vector<double> reff;
  std::vector<int> npairs;
  std::vector< std::complex<double> > xi;
};

/// Two-point correlation function window measurements (I/dataobjs.hpp:230-238).
struct TwoPCFWindowMeasurements {
  int dim = 0;
  std::vector<double> rbin;
  std::vector<double> reff;
  std::vector<int> npairs;
  std::vector< std::complex<double> > xi;
};

}  // namespace trv

#endif  // TRV_B200_DATAOBJS_HPP_
