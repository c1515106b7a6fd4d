// particles.hpp -- trv::ParticleData / trv::ParticleCatalogue
// (mirror of I/particles.hpp:63-90; S/particles.cpp:512-888).
#ifndef TRV_B200_PARTICLES_HPP_
#define TRV_B200_PARTICLES_HPP_

#include <string>
#include <vector>

#include "monitor.hpp"

namespace trv {

struct ParticleData {
  double pos[3];  ///< particle position vector
  double nz;      ///< redshift-dependent expected number density
  double ws;      ///< particle sample weight
  double wc;      ///< particle clustering weight
  double w;       ///< particle overall weight
};

class ParticleCatalogue {
 public:
  std::string source;
  ParticleData* pdata = nullptr;
  int ntotal = 0;
  double wtotal = 0.;
  double wstotal = 0.;
  double pos_min[3] = {0., 0., 0.};
  double pos_max[3] = {0., 0., 0.};
  double pos_span[3] = {0., 0., 0.};

  explicit ParticleCatalogue(int verbose = -1);
  ~ParticleCatalogue();
  ParticleCatalogue(const ParticleCatalogue&) = delete;
  ParticleCatalogue& operator=(const ParticleCatalogue&) = delete;

  void initialise_particles(const int num);
  void finalise_particles();
  void reset_particles();
  ParticleData& operator[](const int pid);

  /// Read a (multi-file, `::`-delimited) plain-text catalogue: whitespace-separated
  /// numeric columns, '#' comment lines; `catalogue_columns` names the columns present,
  /// comma-separated, among x, y, z, nz, ws, wc (missing nz: ntotal / volume; missing
  /// weights: 1) (S/particles.cpp:94-510).  HDF5 files are not supported by this build.
  int load_catalogue_file(
    const std::string& catalogue_filepath,
    const std::string& catalogue_columns,
    const std::string& catalogue_dataset = "",
    double volume = 0.
  );
  int load_particle_data(
    std::vector<double> x, std::vector<double> y, std::vector<double> z,
    std::vector<double> nz, std::vector<double> ws, std::vector<double> wc
  );
  /// Zero-copy-friendly variant of the above taking raw arrays; `nz`, `ws`,
  /// `wc` may be null (0, 1, 1).
  int load_particle_arrays(
    int n, const double* x, const double* y, const double* z,
    const double* nz, const double* ws, const double* wc
  );

  void calc_total_weights();
  void calc_pos_extents(bool init = true);
  void offset_coords(const double dpos[3]);
  void offset_coords_for_periodicity(const double boxsize[3]);

  static void centre_in_box(ParticleCatalogue& catalogue, const double boxsize[3]);
  static void centre_in_box(
    ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref,
    const double boxsize[3]
  );
  static void pad_in_box(
    ParticleCatalogue& catalogue,
    const double boxsize[3], const double boxsize_pad[3]
  );
  static void pad_in_box(
    ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref,
    const double boxsize[3], const double boxsize_pad[3]
  );
  static void pad_grids(
    ParticleCatalogue& catalogue,
    const double boxsize[3], const int ngrid[3], const double ngrid_pad[3]
  );
  static void pad_grids(
    ParticleCatalogue& catalogue, ParticleCatalogue& catalogue_ref,
    const double boxsize[3], const int ngrid[3], const double ngrid_pad[3]
  );
};

}  // namespace trv

#endif  // TRV_B200_PARTICLES_HPP_
