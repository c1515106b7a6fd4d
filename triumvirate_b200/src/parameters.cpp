// parameters.cpp -- ParameterSet::validate: the derivations and checks of
// S/parameters.cpp:466-1270 that the three-point path depends on.
#include "trv/parameters.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace trvs = trv::sys;

namespace trv {

namespace {

bool one_of(const std::string& v, std::initializer_list<const char*> opts) {
  for (const char* o : opts) if (v == o) return true;
  return false;
}

[[noreturn]] void fail(const char* what, const std::string& key, const std::string& val) {
  if (trvs::currTask == 0) trvs::logger.error("%s: `%s` = '%s'.", what, key.c_str(), val.c_str());
  throw trvs::InvalidParameterError("%s: `%s` = '%s'.", what, key.c_str(), val.c_str());
}

}  // namespace

int ParameterSet::validate(bool init) {
  (void)init;  // path transmutations only matter for file I/O (out of scope)

  // Catalogue types (S/parameters.cpp:523-628).
  if (!one_of(this->catalogue_type, {"survey", "random", "sim", "none"})) {
    fail("Catalogue type is not recognised", "catalogue_type", this->catalogue_type);
  }

  // Alignment and padding (S/parameters.cpp:636-657).
  if (!one_of(this->alignment, {"centre", "pad"})) {
    fail("Box alignment choice is not recognised", "alignment", this->alignment);
  }
  if (!one_of(this->padscale, {"box", "grid"})) {
    fail("Pad scale is not recognised", "padscale", this->padscale);
  }

  // Assignment order (S/parameters.cpp:659-683).
  if (this->assignment == "ngp") this->assignment_order = 1;
  else if (this->assignment == "cic") this->assignment_order = 2;
  else if (this->assignment == "tsc") this->assignment_order = 3;
  else if (this->assignment == "pcs") this->assignment_order = 4;
  else fail("Mesh assignment scheme is not supported", "assignment", this->assignment);

  // Interlacing switch (S/parameters.cpp:685-702).
  if (one_of(this->interlace, {"true", "on"})) this->interlace = "true";
  else if (one_of(this->interlace, {"false", "off"})) this->interlace = "false";
  else fail("Interlacing must be 'true'/'on' or 'false'/'off'", "interlace", this->interlace);

  // Statistic type -> npoint, space (S/parameters.cpp:704-827).
  struct StatRule { const char* stat; const char* npoint; const char* space; int needs; };
  // needs: 0 any catalogue but random/none, 1 random only, 2 anything.
  static const StatRule rules[] = {
    {"powspec", "2pt", "fourier", 0}, {"2pcf", "2pt", "config", 0},
    {"2pcf-win", "2pt", "config", 1}, {"bispec", "3pt", "fourier", 0},
    {"3pcf", "3pt", "config", 0}, {"3pcf-win", "3pt", "config", 1},
    {"3pcf-win-wa", "3pt", "config", 1}, {"modes", "none", "fourier", 2},
    {"pairs", "none", "config", 2},
  };
  bool found = false;
  for (const StatRule& r : rules) {
    if (this->statistic_type != r.stat) continue;
    found = true;
    this->npoint = r.npoint; this->space = r.space;
    const bool is_rand_like = one_of(this->catalogue_type, {"random", "none"});
    if ((r.needs == 0 && is_rand_like)
        || (r.needs == 1 && this->catalogue_type != "random")) {
      fail("Statistic and catalogue types are incompatible", "catalogue_type",
           this->catalogue_type);
    }
  }
  if (!found) fail("Statistic type is not recognised", "statistic_type", this->statistic_type);

  // Form -> shape (S/parameters.cpp:829-849): 'full' with equal degrees is
  // measured on the upper triangle only.
  if (!one_of(this->form, {"full", "diag", "off-diag", "row"})) {
    fail("`form` must be 'full', 'diag', 'off-diag' or 'row'", "form", this->form);
  }
  this->shape = (this->form == "full" && this->ell1 == this->ell2) ? "triu" : this->form;

  if (!one_of(this->norm_convention, {"none", "particle", "mesh", "mesh-mixed"})) {
    fail("`norm_convention` is not recognised", "norm_convention", this->norm_convention);
  }
  if (this->norm_convention == "mesh-mixed" && this->npoint != "2pt") {
    fail("'mesh-mixed' normalisation only applies to two-point statistics", "npoint",
         this->npoint);
  }
  if (!one_of(this->binning, {"lin", "log", "linpad", "logpad", "custom"})) {
    fail("Binning scheme is unrecognised", "binning", this->binning);
  }

  // FFTW options are accepted for compatibility and neutralised, as the
  // reference does in GPU mode (S/parameters.cpp:924-943).
  this->fftw_scheme = "";
  this->use_fftw_wisdom = "";
  if (this->save_binned_vectors == "false") this->save_binned_vectors = "";

  // Derived mesh quantities (S/parameters.cpp:1028-1031).
  this->volume = this->boxsize[0] * this->boxsize[1] * this->boxsize[2];
  this->nmesh = static_cast<long long>(this->ngrid[0]) * this->ngrid[1] * this->ngrid[2];
  if (this->volume <= 0.) {
    if (this->expand < 1.) {
      throw trvs::InvalidParameterError(
        "Box expansion factor must be >= 1: `expand` = %lg.", this->expand);
    }
  }
  if (this->nmesh <= 0 && this->cutoff_nyq < 0.) {
    throw trvs::InvalidParameterError(
      "Nyquist cutoff must be non-negative: `cutoff_nyq` = %lg.", this->cutoff_nyq);
  }

  if (this->alignment == "pad") {
    if (this->padfactor < 0.) {
      throw trvs::InvalidParameterError(
        "Padding is negative: `padfactor` = %lg.", this->padfactor);
    }
    if (this->padscale == "box" && this->padfactor >= 1.) {
      throw trvs::InvalidParameterError(
        "Padding is too large (exceeding box size): `padfactor` = %lg.", this->padfactor);
    }
    if (this->padscale == "grid" && (
          this->padfactor >= this->ngrid[0] || this->padfactor >= this->ngrid[1]
          || this->padfactor >= this->ngrid[2])) {
      throw trvs::InvalidParameterError(
        "Padding is too large (exceeding mesh size): `padfactor` = %lg.", this->padfactor);
    }
  }

  // Measurement range (S/parameters.cpp:1110-1238).
  if (this->bin_min < 0.) {
    throw trvs::InvalidParameterError("Measurement range limits must be non-negative.");
  }
  if (this->bin_min >= this->bin_max) {
    throw trvs::InvalidParameterError(
      "Measurement range lower limit must be less than the upper limit.");
  }
  if (this->nmesh > 0 && this->volume > 0. && trvs::currTask == 0) {
    const int ngrid_min = *std::min_element(this->ngrid, this->ngrid + 3);
    const double box_max = *std::max_element(this->boxsize, this->boxsize + 3);
    if (this->space == "fourier" && this->bin_min > M_PI * ngrid_min / box_max) {
      trvs::logger.warn("Measurement range lower limit exceeds the Nyquist wavenumber.");
    }
    if (this->space == "config" && this->bin_max < 2 * box_max / ngrid_min) {
      trvs::logger.warn("Measurement range upper limit is below the Nyquist separation.");
    }
  }
  if (this->num_bins < 2) {
    throw trvs::InvalidParameterError("Number of bins `num_bins` must be >= 2.");
  }
  if (this->idx_bin < 0 && this->npoint == "3pt" && this->form == "row") {
    throw trvs::InvalidParameterError(
      "Fixed row bin index `idx_bin` must be >= 0 when `form` = 'row'.");
  }
  if (this->binning == "linpad" || this->binning == "logpad") {
    const int nbin_pad = 5;
    if (this->num_bins < nbin_pad + 2) {
      throw trvs::InvalidParameterError(
        "Binning scheme '%s' requires `num_bins` >= %d.", this->binning.c_str(), nbin_pad + 2);
    }
  }
  if (std::abs(this->idx_bin) >= this->num_bins) {
    throw trvs::InvalidParameterError(
      "Bin index `idx_bin` must be less than `num_bins` in absolute value.");
  }

  // Interlacing is unsupported for three-point statistics and is switched
  // off here exactly as the reference does (S/parameters.cpp:1240-1249).
  if (this->npoint == "3pt" && this->interlace == "true") {
    this->interlace = "false";
    if (trvs::currTask == 0) {
      trvs::logger.warn(
        "Interlacing is unsupported for 3-point measurements; `interlace` is set to 'false'.");
    }
  }

  // B200 extension: deterministic assignment from the environment.
  const char* det = std::getenv("TRV_DETERMINISTIC");
  if (det != nullptr && std::string(det) != "0" && std::string(det) != "") {
    this->deterministic = 1;
  }
  if (this->part_count < 1 || this->part_rank < 0 || this->part_rank >= this->part_count) {
    throw trvs::InvalidParameterError(
      "Invalid work partition: rank %d of %d.", this->part_rank, this->part_count);
  }
  return 0;
}

}  // namespace trv
