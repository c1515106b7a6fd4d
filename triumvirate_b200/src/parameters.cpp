// parameters.cpp -- ParameterSet::validate: the derivations and checks of
// S/parameters.cpp:466-1270 that the three-point path depends on.
#include "trv/parameters.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

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

namespace {

bool is_blank(const std::string& v) {
  return v.find_first_not_of(" \t\n\r\v\f") == std::string::npos;
}

/// Catalogue paths of one `*_catalogue_file` entry: split at the multi-file delimiter,
/// relative names joined with the catalogue directory on the first validation,
/// ${VAR} expanded (S/parameters.cpp:489-600).
std::vector<std::string> resolve_catalogue_files(std::string& entry, const std::string& dir,
                                                 bool init) {
  std::vector<std::string> files;
  if (is_blank(entry)) return files;
  if (!trvs::has_extension(entry, trvs::fn_delimiter)) entry += trvs::fn_delimiter;
  for (std::string f : trvs::split_string(entry, trvs::fn_delimiter)) {
    if (init && f.rfind("/", 0) != 0) f = dir + f;
    trvs::expand_envar_in_path(f);
    files.push_back(f);
  }
  return files;
}

}  // namespace

int ParameterSet::read_from_file(char* parameter_filepath) {
  std::ifstream fin(parameter_filepath);
  std::map<std::string, std::string> kv;
  std::string line;
  while (std::getline(fin, line)) {
    if (line.rfind("#", 0) == 0) continue;                  // comment line
    std::istringstream iss(line);
    std::string key, eq, value;
    if (!(iss >> key >> eq >> value) || eq != "=") continue;   // not `key = value`
    kv[key] = value;
  }
  auto str = [&](const char* key, std::string& dst, bool keep_default) {
    auto it = kv.find(key);
    if (it != kv.end()) dst = it->second;
    else if (!keep_default) dst.clear();
  };
  auto num = [&](const char* key, double& dst) {
    auto it = kv.find(key);
    if (it != kv.end()) dst = std::strtod(it->second.c_str(), nullptr);
  };
  auto integer = [&](const char* key, int& dst) {
    auto it = kv.find(key);
    if (it != kv.end()) dst = static_cast<int>(std::strtol(it->second.c_str(), nullptr, 10));
  };
  // strings without a built-in default are reset when absent, the others keep theirs
  str("catalogue_dir", this->catalogue_dir, false);
  str("measurement_dir", this->measurement_dir, false);
  str("data_catalogue_file", this->data_catalogue_file, false);
  str("rand_catalogue_file", this->rand_catalogue_file, false);
  str("catalogue_columns", this->catalogue_columns, false);
  str("catalogue_dataset", this->catalogue_dataset, false);
  str("output_tag", this->output_tag, false);
  str("catalogue_type", this->catalogue_type, false);
  str("statistic_type", this->statistic_type, false);
  str("alignment", this->alignment, true);
  str("padscale", this->padscale, true);
  str("assignment", this->assignment, true);
  str("interlace", this->interlace, true);
  str("form", this->form, true);
  str("norm_convention", this->norm_convention, true);
  str("binning", this->binning, true);
  str("fftw_scheme", this->fftw_scheme, true);
  str("use_fftw_wisdom", this->use_fftw_wisdom, true);
  str("save_binned_vectors", this->save_binned_vectors, true);
  str("progbar", this->progbar, true);
  double box[3] = {0., 0., 0.};
  int grid[3] = {0, 0, 0};
  num("boxsize_x", box[0]); num("boxsize_y", box[1]); num("boxsize_z", box[2]);
  integer("ngrid_x", grid[0]); integer("ngrid_y", grid[1]); integer("ngrid_z", grid[2]);
  for (int ax = 0; ax < 3; ax++) { this->boxsize[ax] = box[ax]; this->ngrid[ax] = grid[ax]; }
  num("expand", this->expand);
  num("padfactor", this->padfactor);
  num("cutoff_nyq", this->cutoff_nyq);
  integer("ell1", this->ell1); integer("ell2", this->ell2); integer("ELL", this->ELL);
  integer("i_wa", this->i_wa); integer("j_wa", this->j_wa);
  num("bin_min", this->bin_min); num("bin_max", this->bin_max);
  integer("num_bins", this->num_bins); integer("idx_bin", this->idx_bin);
  integer("verbose", this->verbose);
  this->volume = box[0] * box[1] * box[2];
  this->nmesh = static_cast<long long>(grid[0]) * grid[1] * grid[2];
  return this->validate(true);
}

int ParameterSet::validate(bool init) {
  trvs::logger.reset_level(this->verbose);

  // Directories and catalogue files (S/parameters.cpp:470-600).
  if (init && !is_blank(this->catalogue_dir)) this->catalogue_dir += "/";
  if (is_blank(this->measurement_dir)) this->measurement_dir = "./";
  else if (init) this->measurement_dir += "/";
  trvs::expand_envar_in_path(this->catalogue_dir);
  trvs::expand_envar_in_path(this->measurement_dir);
  if (this->catalogue_type == "random") this->data_catalogue_file = "";
  if (this->catalogue_type == "sim") this->rand_catalogue_file = "";
  this->data_catalogue_files.clear();
  this->rand_catalogue_files.clear();
  if (this->catalogue_type == "survey" || this->catalogue_type == "sim") {
    this->data_catalogue_files =
      resolve_catalogue_files(this->data_catalogue_file, this->catalogue_dir, init);
  }
  if (this->catalogue_type == "survey" || this->catalogue_type == "random") {
    this->rand_catalogue_files =
      resolve_catalogue_files(this->rand_catalogue_file, this->catalogue_dir, init);
  }
  // the `*_catalogue_file` strings become the resolved paths joined by the delimiter
  // (S/parameters.cpp:627-632): what load_catalogue_file() is handed
  auto rejoin = [](const std::vector<std::string>& files) {
    std::string joined;
    for (size_t i = 0; i < files.size(); i++) {
      if (i) joined += trvs::fn_delimiter;
      joined += files[i];
    }
    return joined;
  };
  this->data_catalogue_file = rejoin(this->data_catalogue_files);
  this->rand_catalogue_file = rejoin(this->rand_catalogue_files);

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

int ParameterSet::print_to_file(char* out_parameter_filepath) {
  std::FILE* out = std::fopen(out_parameter_filepath, "w");
  if (out == nullptr) {
    if (trvs::currTask == 0) {
      trvs::logger.error("Non-existent or unwritable output directory: %s",
                         this->measurement_dir.c_str());
    }
    throw trvs::IOError("Non-existent or unwritable output directory: %s",
                        this->measurement_dir.c_str());
  }
  auto s_ = [out](const char* key, const std::string& v) { std::fprintf(out, "%s = %s\n", key, v.c_str()); };
  auto i_ = [out](const char* key, long long v) { std::fprintf(out, "%s = %lld\n", key, v); };
  auto f_ = [out](const char* key, const char* fmt, double v) {
    std::fprintf(out, "%s = ", key); std::fprintf(out, fmt, v); std::fprintf(out, "\n");
  };
  s_("catalogue_dir", this->catalogue_dir);
  s_("measurement_dir", this->measurement_dir);
  s_("data_catalogue_file", this->data_catalogue_file);
  s_("rand_catalogue_file", this->rand_catalogue_file);
  s_("catalogue_columns", this->catalogue_columns);
  s_("catalogue_dataset", this->catalogue_dataset);
  s_("output_tag", this->output_tag);
  f_("boxsize_x", "%.3f", this->boxsize[0]);
  f_("boxsize_y", "%.3f", this->boxsize[1]);
  f_("boxsize_z", "%.3f", this->boxsize[2]);
  i_("ngrid_x", this->ngrid[0]); i_("ngrid_y", this->ngrid[1]); i_("ngrid_z", this->ngrid[2]);
  f_("volume", "%.6e", this->volume);
  i_("nmesh", this->nmesh);
  f_("expand", "%.4f", this->expand);
  s_("alignment", this->alignment);
  s_("padscale", this->padscale);
  f_("padfactor", "%.4f", this->padfactor);
  s_("assignment", this->assignment);
  s_("interlace", this->interlace);
  i_("assignment_order", this->assignment_order);
  s_("catalogue_type", this->catalogue_type);
  s_("statistic_type", this->statistic_type);
  s_("npoint", this->npoint);
  s_("space", this->space);
  i_("ell1", this->ell1); i_("ell2", this->ell2); i_("ELL", this->ELL);
  i_("i_wa", this->i_wa); i_("j_wa", this->j_wa);
  s_("form", this->form);
  s_("norm_convention", this->norm_convention);
  s_("binning", this->binning);
  s_("shape", this->shape);
  f_("bin_min", "%.4f", this->bin_min);
  f_("bin_max", "%.4f", this->bin_max);
  i_("num_bins", this->num_bins);
  i_("idx_bin", this->idx_bin);
  s_("fftw_scheme", this->fftw_scheme);
  s_("use_fftw_wisdom", this->use_fftw_wisdom);
  s_("fftw_wisdom_file_f", this->fftw_wisdom_file_f);
  s_("fftw_wisdom_file_b", this->fftw_wisdom_file_b);
  s_("save_binned_vectors", this->save_binned_vectors);
  s_("progbar", this->progbar);
  i_("verbose", this->verbose);
  i_("fftw_planner_flag", this->fftw_planner_flag);
  std::fclose(out);
  if (trvs::currTask == 0) {
    trvs::logger.info("Check used-parameter file for reference: %s", out_parameter_filepath);
  }
  return 0;
}

int ParameterSet::print_to_file() {
  std::string path = this->measurement_dir + "parameters_used" + this->output_tag;
  return this->print_to_file(&path[0]);
}

void override_paramset_by_envvars(trv::ParameterSet& params) {
  if (const char* v = std::getenv("TRV_OVERRIDE_OUTPUT_TAG")) params.output_tag = v;
  if (const char* v = std::getenv("TRV_OVERRIDE_VERBOSE")) params.verbose = std::stoi(v);
  if (const char* v = std::getenv("TRV_OVERRIDE_PROGBAR")) params.progbar = v;
  // TRV_OVERRIDE_FFTW_SCHEME / TRV_OVERRIDE_USE_FFTW_WISDOM only act in the reference's
  // CPU mode (S/parameters.cpp:1389-1397); this build always runs on the device.
  params.validate();
}

void set_boxsize_from_expand(const double* spans, trv::ParameterSet& params) {
  for (int ax = 0; ax < 3; ax++) params.boxsize[ax] = spans[ax] * params.expand;
  params.validate();
  if (trvs::currTask == 0) {
    trvs::logger.info(
      "Box size has been set from particle coordinate spans and expansion factor: "
      "(%.3f, %.3f, %.3f).", params.boxsize[0], params.boxsize[1], params.boxsize[2]);
  }
}

void set_ngrid_from_cutoff(trv::ParameterSet& params) {
  for (int ax = 0; ax < 3; ax++) {
    double cells;
    if (params.space == "fourier") cells = params.boxsize[ax] * params.cutoff_nyq / M_PI;
    else if (params.space == "config") cells = 2. * params.boxsize[ax] / params.cutoff_nyq;
    else throw trvs::InvalidParameterError(
      "Space must be 'fourier' or 'config': `space` = '%s'.", params.space.c_str());
    const int n = static_cast<int>(std::ceil(cells));
    params.ngrid[ax] = n + (n % 2);
  }
  params.validate();
  if (trvs::currTask == 0) {
    trvs::logger.info(
      "Mesh grid numbers have been set from Nyquist cutoff and box size: (%d, %d, %d).",
      params.ngrid[0], params.ngrid[1], params.ngrid[2]);
  }
}

}  // namespace trv
