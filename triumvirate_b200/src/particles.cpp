// particles.cpp -- ParticleCatalogue: host-side AoS catalogue with the
// reference's public members and alignment helpers (S/particles.cpp:512-888).
#include "trv/particles.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace trvs = trv::sys;

namespace trv {

namespace {

void require_data(const ParticleCatalogue& c) {
  if (c.pdata == nullptr) {
    if (trvs::currTask == 0) trvs::logger.error("Particle data are uninitialised.");
    throw trvs::InvalidDataError("Particle data are uninitialised.");
  }
}

void warn_if_overflow(ParticleCatalogue& c, const double boxsize[3], const char* op) {
  c.calc_pos_extents(false);
  for (int ax = 0; ax < 3; ax++) {
    if (c.pos_span[ax] > boxsize[ax] && trvs::currTask == 0) {
      trvs::logger.warn(
        "Catalogue extent exceeds the box size along axis %d: span = %.3f, "
        "boxsize = %.3f (source=%s). Some particles may lie outside the box after %s.",
        ax, c.pos_span[ax], boxsize[ax], c.source.c_str(), op);
    }
  }
}

}  // namespace

ParticleCatalogue::ParticleCatalogue(int verbose) {
  if (verbose >= 0) trvs::logger.reset_level(verbose);
}

ParticleCatalogue::~ParticleCatalogue() { this->finalise_particles(); }

void ParticleCatalogue::initialise_particles(const int num) {
  if (num <= 0) {
    if (trvs::currTask == 0) trvs::logger.error("Number of particles is non-positive.");
    throw trvs::InvalidParameterError("Number of particles is non-positive.");
  }
  this->reset_particles();
  this->ntotal = num;
  this->pdata = new ParticleData[num];
  trvs::gbytesMem += trvs::size_in_gb<ParticleData>(num);
  trvs::update_maxmem();
}

void ParticleCatalogue::finalise_particles() { this->reset_particles(); }

void ParticleCatalogue::reset_particles() {
  if (this->pdata != nullptr) {
    delete[] this->pdata;
    this->pdata = nullptr;
    trvs::gbytesMem -= trvs::size_in_gb<ParticleData>(this->ntotal);
  }
}

ParticleData& ParticleCatalogue::operator[](const int pid) { return this->pdata[pid]; }

namespace {

/// One catalogue file in memory with the start of every data line (not empty, not '#').
struct TextFile {
  std::string bytes;
  std::vector<size_t> line_start;
};

TextFile slurp_data_lines(const std::string& path) {
  TextFile f;
  std::FILE* fp = std::fopen(path.c_str(), "rb");
  if (fp == nullptr) {
    if (trvs::currTask == 0) trvs::logger.error("Failed to open file: %s", path.c_str());
    throw trvs::IOError("Failed to open file: %s", path.c_str());
  }
  std::fseek(fp, 0, SEEK_END);
  const long size = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  f.bytes.resize(size > 0 ? (size_t)size : 0);
  if (size > 0 && std::fread(&f.bytes[0], 1, (size_t)size, fp) != (size_t)size) {
    std::fclose(fp);
    throw trvs::IOError("Failed to read file: %s", path.c_str());
  }
  std::fclose(fp);
  f.bytes.push_back('\n');   // sentinel: every line ends in a newline
  size_t at = 0;
  const size_t n = f.bytes.size();
  while (at < n) {
    const size_t eol = f.bytes.find('\n', at);
    if (eol > at && f.bytes[at] != '#') f.line_start.push_back(at);
    at = eol + 1;
  }
  return f;
}

}  // namespace

int ParticleCatalogue::load_catalogue_file(
  const std::string& catalogue_filepath, const std::string& catalogue_columns,
  const std::string& catalogue_dataset, double volume
) {
  (void)catalogue_dataset;   // HDF5 only
  if (!this->source.empty()) {
    if (trvs::currTask == 0) {
      trvs::logger.error("Catalogue already loaded from another source: %s.", this->source.c_str());
    }
    throw trvs::InvalidDataError(
      "Catalogue already loaded from another source: %s.", this->source.c_str());
  }
  this->source = "extfile:" + catalogue_filepath;
  if (trvs::has_extension(catalogue_filepath, ".h5")
      || trvs::has_extension(catalogue_filepath, ".hdf5")) {
    if (trvs::currTask == 0) {
      trvs::logger.error("HDF5 file format is not supported in this build: %s",
                         catalogue_filepath.c_str());
    }
    throw trvs::InvalidDataError("HDF5 file format is not supported in this build: %s",
                                 catalogue_filepath.c_str());
  }

  // column of each of x, y, z, nz, ws, wc in the file (-1: absent)
  const char* wanted[6] = {"x", "y", "z", "nz", "ws", "wc"};
  int column[6] = {-1, -1, -1, -1, -1, -1};
  {
    std::istringstream names(catalogue_columns);
    std::string name;
    for (int col = 0; std::getline(names, name, ','); col++) {
      for (int q = 0; q < 6; q++) if (column[q] < 0 && name == wanted[q]) column[q] = col;
    }
  }
  if (column[0] < 0 || column[1] < 0 || column[2] < 0) {
    throw trvs::InvalidDataError(
      "Catalogue columns must name 'x', 'y' and 'z': `catalogue_columns` = '%s'.",
      catalogue_columns.c_str());
  }
  const int ncol_needed = 1 + *std::max_element(column, column + 6);

  std::vector<TextFile> files;
  long long nentry = 0;
  for (const std::string& path : trvs::split_string(catalogue_filepath, trvs::fn_delimiter)) {
    files.push_back(slurp_data_lines(path));
    nentry += (long long)files.back().line_start.size();
  }
  if (nentry > 2147483647LL) {
    throw trvs::InvalidDataError("Catalogue holds more than 2^31 - 1 particles.");
  }
  this->initialise_particles(static_cast<int>(nentry));
  if (column[3] < 0 && trvs::currTask == 0) {
    trvs::logger.info(
      "Catalogue 'nz' field is unavailable and will be set to the mean density in the "
      "bounding box (source=%s).", this->source.c_str());
  }
  const double nz_default = (volume > 0.) ? this->ntotal / volume : 0.;

  // the lines are independent: parsed in parallel (strtod on the in-memory text)
  long long base = 0;
  bool short_row = false;
  for (const TextFile& f : files) {
    const long long nline = (long long)f.line_start.size();
#pragma omp parallel for schedule(static) reduction(||:short_row)
    for (long long ln = 0; ln < nline; ln++) {
      const char* p = f.bytes.data() + f.line_start[ln];
      double row[64];
      int nread = 0;
      while (nread < 64) {
        char* end = nullptr;
        const double v = std::strtod(p, &end);
        if (end == p) break;
        row[nread++] = v;
        p = end;
        while (*p == ' ' || *p == '\t' || *p == '\r') p++;
        if (*p == '\n') break;
      }
      if (nread < ncol_needed) { short_row = true; continue; }
      ParticleData& pt = this->pdata[base + ln];
      pt.pos[0] = row[column[0]]; pt.pos[1] = row[column[1]]; pt.pos[2] = row[column[2]];
      pt.nz = column[3] >= 0 ? row[column[3]] : nz_default;
      pt.ws = column[4] >= 0 ? row[column[4]] : 1.;
      pt.wc = column[5] >= 0 ? row[column[5]] : 1.;
      pt.w = pt.ws * pt.wc;
    }
    base += nline;
  }
  if (short_row) {
    throw trvs::InvalidDataError(
      "Catalogue rows hold fewer columns than `catalogue_columns` names (source=%s).",
      this->source.c_str());
  }
  this->calc_total_weights();
  this->calc_pos_extents();
  return 0;
}

int ParticleCatalogue::load_particle_data(
  std::vector<double> x, std::vector<double> y, std::vector<double> z,
  std::vector<double> nz, std::vector<double> ws, std::vector<double> wc
) {
  const std::size_t n = x.size();
  if (!(y.size() == n && z.size() == n && nz.size() == n && ws.size() == n
        && wc.size() == n)) {
    this->source = "extdata";
    if (trvs::currTask == 0) {
      trvs::logger.error("Inconsistent particle data dimensions (source=%s).",
                         this->source.c_str());
    }
    throw trvs::InvalidDataError("Inconsistent particle data dimensions (source=%s).",
                                 this->source.c_str());
  }
  return this->load_particle_arrays(static_cast<int>(n), x.data(), y.data(), z.data(),
                                    nz.data(), ws.data(), wc.data());
}

int ParticleCatalogue::load_particle_arrays(
  int n, const double* x, const double* y, const double* z,
  const double* nz, const double* ws, const double* wc
) {
  this->source = "extdata";
  this->initialise_particles(n);
#pragma omp parallel for
  for (int pid = 0; pid < n; pid++) {
    ParticleData& p = this->pdata[pid];
    p.pos[0] = x[pid]; p.pos[1] = y[pid]; p.pos[2] = z[pid];
    p.nz = nz ? nz[pid] : 0.;
    p.ws = ws ? ws[pid] : 1.;
    p.wc = wc ? wc[pid] : 1.;
    p.w = p.ws * p.wc;   // S/particles.cpp:556
  }
  this->calc_total_weights();
  this->calc_pos_extents();
  return 0;
}

void ParticleCatalogue::calc_total_weights() {
  require_data(*this);
  // Fixed chunks summed in parallel, chunk sums added in order: the totals (and the
  // alpha contrast derived from them) are the same bits whatever the thread count or
  // schedule -- an OpenMP `reduction` (S/particles.cpp:578-591) combines in an
  // unspecified order, which made survey statistics differ in the last bit from call
  // to call.
  const int chunk = 1 << 16;
  const int nchunks = (this->ntotal + chunk - 1) / chunk;
  std::vector<double> part_w(nchunks, 0.), part_ws(nchunks, 0.);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nchunks; c++) {
    const int lo = c * chunk, hi = std::min(this->ntotal, lo + chunk);
    double wt = 0., wst = 0.;
    for (int pid = lo; pid < hi; pid++) {
      wt += this->pdata[pid].w;
      wst += this->pdata[pid].ws;
    }
    part_w[c] = wt; part_ws[c] = wst;
  }
  double wt = 0., wst = 0.;
  for (int c = 0; c < nchunks; c++) { wt += part_w[c]; wst += part_ws[c]; }
  this->wtotal = wt;
  this->wstotal = wst;
}

void ParticleCatalogue::calc_pos_extents(bool init) {
  (void)init;
  require_data(*this);
  double lo[3], hi[3];
  for (int ax = 0; ax < 3; ax++) lo[ax] = hi[ax] = this->pdata[0].pos[ax];
#pragma omp parallel for reduction(min:lo[:3]) reduction(max:hi[:3])
  for (int pid = 0; pid < this->ntotal; pid++) {
    for (int ax = 0; ax < 3; ax++) {
      const double v = this->pdata[pid].pos[ax];
      if (v < lo[ax]) lo[ax] = v;
      if (v > hi[ax]) hi[ax] = v;
    }
  }
  for (int ax = 0; ax < 3; ax++) {
    this->pos_min[ax] = lo[ax];
    this->pos_max[ax] = hi[ax];
    this->pos_span[ax] = hi[ax] - lo[ax];
  }
}

void ParticleCatalogue::offset_coords(const double dpos[3]) {
  require_data(*this);
#pragma omp parallel for
  for (int pid = 0; pid < this->ntotal; pid++) {
    for (int ax = 0; ax < 3; ax++) this->pdata[pid].pos[ax] -= dpos[ax];
  }
  this->calc_pos_extents();
}

void ParticleCatalogue::offset_coords_for_periodicity(const double boxsize[3]) {
  // Wrap only, no centring (S/particles.cpp:678-699).
#pragma omp parallel for
  for (int pid = 0; pid < this->ntotal; pid++) {
    for (int ax = 0; ax < 3; ax++) {
      double& v = this->pdata[pid].pos[ax];
      if (v >= boxsize[ax]) v = std::fmod(v, boxsize[ax]);
      else if (v < 0.) v = std::fmod(v, boxsize[ax]) + boxsize[ax];
    }
  }
  this->calc_pos_extents();
}

void ParticleCatalogue::centre_in_box(ParticleCatalogue& catalogue, const double boxsize[3]) {
  warn_if_overflow(catalogue, boxsize, "centring");
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = (catalogue.pos_min[ax] + catalogue.pos_max[ax]) / 2. - boxsize[ax] / 2.;
  }
  catalogue.offset_coords(dvec);
}

void ParticleCatalogue::centre_in_box(
  ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref, const double boxsize[3]
) {
  warn_if_overflow(catalogue, boxsize, "centring");
  warn_if_overflow(catalogue_ref, boxsize, "centring");
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = (catalogue_ref.pos_min[ax] + catalogue_ref.pos_max[ax]) / 2.
      - boxsize[ax] / 2.;
  }
  catalogue_ref.offset_coords(dvec);
  catalogue.offset_coords(dvec);
}

void ParticleCatalogue::pad_in_box(
  ParticleCatalogue& catalogue, const double boxsize[3], const double boxsize_pad[3]
) {
  warn_if_overflow(catalogue, boxsize, "padding");
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = catalogue.pos_min[ax] - boxsize_pad[ax] * boxsize[ax];
  }
  catalogue.offset_coords(dvec);
}

void ParticleCatalogue::pad_in_box(
  ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref,
  const double boxsize[3], const double boxsize_pad[3]
) {
  warn_if_overflow(catalogue, boxsize, "padding");
  warn_if_overflow(catalogue_ref, boxsize, "padding");
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = catalogue_ref.pos_min[ax] - boxsize_pad[ax] * boxsize[ax];
  }
  catalogue_ref.offset_coords(dvec);
  catalogue.offset_coords(dvec);
}

void ParticleCatalogue::pad_grids(
  ParticleCatalogue& catalogue,
  const double boxsize[3], const int ngrid[3], const double ngrid_pad[3]
) {
  catalogue.calc_pos_extents(false);
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = catalogue.pos_min[ax];
    dvec[ax] -= ngrid_pad[ax] * boxsize[ax] / double(ngrid[ax]);
  }
  catalogue.offset_coords(dvec);
}

void ParticleCatalogue::pad_grids(
  ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref,
  const double boxsize[3], const int ngrid[3], const double ngrid_pad[3]
) {
  catalogue_ref.calc_pos_extents(false);
  double dvec[3];
  for (int ax = 0; ax < 3; ax++) {
    dvec[ax] = catalogue_ref.pos_min[ax]
      - ngrid_pad[ax] * boxsize[ax] / double(ngrid[ax]);
  }
  catalogue_ref.offset_coords(dvec);
  catalogue.offset_coords(dvec);
}

}  // namespace trv
