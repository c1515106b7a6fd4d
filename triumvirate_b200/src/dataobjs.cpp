// dataobjs.cpp -- Binning (S/dataobjs.cpp:134-249).  Edge, centre and width
// expressions are kept operation-for-operation because shell membership is
// decided against these doubles.
#include "trv/dataobjs.hpp"

#include <algorithm>
#include <cmath>

namespace trvs = trv::sys;

namespace trv {

Binning::Binning(std::string space, std::string scheme)
  : space(space), scheme(scheme) {}

Binning::Binning(trv::ParameterSet& params)
  : Binning::Binning(params.space, params.binning) {
  this->bin_min = params.bin_min;
  this->bin_max = params.bin_max;
  this->num_bins = params.num_bins;
  const double box_max = *std::max_element(params.boxsize, params.boxsize + 3);
  const int ngrid_min = *std::min_element(params.ngrid, params.ngrid + 3);
  this->dbin_pad_config = (1 + 5.e-3) * box_max / ngrid_min;
  this->dbin_pad_fourier = (1 + 5.e-3) * (2 * M_PI) / box_max;
}

void Binning::set_bins(double coord_min, double coord_max, int nbin) {
  if (coord_min < 0.) {
    throw trvs::InvalidParameterError("Bin range must be non-negative.");
  }
  if (nbin <= 0) {
    throw trvs::InvalidParameterError("Number of bins must be positive.");
  }
  this->bin_min = coord_min;
  this->bin_max = coord_max;
  this->num_bins = nbin;
  this->set_bins();
}

void Binning::set_bins() {
  this->bin_edges.clear();
  this->bin_centres.clear();
  this->bin_widths.clear();
  this->compute_binning();
}

void Binning::set_bins(double boxsize_max, int ngrid_min) {
  this->scheme = "lin";
  this->bin_min = 0.;
  double bin_width = 0.;
  if (this->space == "config") {
    bin_width = boxsize_max / double(ngrid_min);
    this->bin_max = boxsize_max / 2.;
  } else if (this->space == "fourier") {
    bin_width = 2 * M_PI / boxsize_max;
    this->bin_max = M_PI * double(ngrid_min) / boxsize_max;
  }
  this->bin_max += bin_width / 2.;
  this->num_bins = ngrid_min / 2;
  this->set_bins();
}

void Binning::set_bins(std::vector<double> edges) {
  this->bin_min = edges.front();
  this->bin_max = edges.back();
  this->num_bins = static_cast<int>(edges.size()) - 1;
  this->scheme = "custom";
  this->bin_edges = edges;
  this->bin_centres.clear();
  this->bin_widths.clear();
  for (int ibin = 0; ibin < this->num_bins; ibin++) {
    this->bin_centres.push_back((edges[ibin] + edges[ibin + 1]) / 2.);
    this->bin_widths.push_back(edges[ibin + 1] - edges[ibin]);
  }
}

void Binning::compute_binning() {
  const double dbin_pad =
    (this->space == "config") ? this->dbin_pad_config : this->dbin_pad_fourier;
  const bool padded = (this->scheme == "linpad" || this->scheme == "logpad");
  const bool logarithmic = (this->scheme == "log" || this->scheme == "logpad");
  if (!(this->scheme == "lin" || padded || logarithmic)) {
    throw trvs::InvalidParameterError(
      "Unrecognised/unsupported binning `scheme`: %s.", this->scheme.c_str());
  }

  auto push = [this](double left, double centre, double width) {
    this->bin_edges.push_back(left);
    this->bin_centres.push_back(centre);
    this->bin_widths.push_back(width);
  };

  int first = 0;
  double lower = this->bin_min;
  if (padded) {
    for (int ibin = 0; ibin < this->nbin_pad; ibin++) {
      const double left = dbin_pad * ibin;
      push(left, left + dbin_pad / 2., dbin_pad);
    }
    first = this->nbin_pad;
    lower = dbin_pad * this->nbin_pad;
  }

  if (!logarithmic) {
    const double dbin = (this->bin_max - lower) / double(this->num_bins - first);
    for (int ibin = first; ibin < this->num_bins; ibin++) {
      const double left = lower + dbin * (ibin - first);
      push(left, left + dbin / 2., dbin);
    }
  } else {
    if (lower == 0.) {
      throw trvs::InvalidParameterError(
        "Cannot use logarithmic binning when the lowest edge is zero.");
    }
    const double dlnbin = (std::log(this->bin_max) - std::log(lower))
      / double(this->num_bins - first);
    for (int ibin = first; ibin < this->num_bins; ibin++) {
      const double left = lower * std::exp(dlnbin * (ibin - first));
      const double right = lower * std::exp(dlnbin * (ibin - first + 1));
      push(left, (left + right) / 2., right - left);
    }
  }
  this->bin_edges.push_back(this->bin_max);
}

}  // namespace trv
