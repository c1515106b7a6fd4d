// io.cpp -- writers of the measurement files (format of S/io.cpp:102-558: the lines and
// number formats below are the file format itself and have to match byte for byte; the
// code around them is organised by table layout instead of one function per statistic).
#include "trv/io.hpp"

#include <complex>
#include <vector>

namespace trv {
namespace io {

namespace {

void header_catalogue(std::FILE* f, const char* label_source, const char* label_size,
                      const char* label_extents, trv::ParticleCatalogue& c) {
  std::fprintf(f, "%s %s: %s\n", comment_delimiter, label_source, c.source.c_str());
  std::fprintf(f, "%s %s: ntotal = %d, wtotal = %.3f, wstotal = %.3f\n",
               comment_delimiter, label_size, c.ntotal, c.wtotal, c.wstotal);
  std::fprintf(f, "%s %s: [(%.3f, %.3f), (%.3f, %.3f), (%.3f, %.3f)]\n",
               comment_delimiter, label_extents,
               c.pos_min[0], c.pos_max[0], c.pos_min[1], c.pos_max[1], c.pos_min[2], c.pos_max[2]);
}

void header_mesh_and_norm(std::FILE* f, trv::ParameterSet& p, double norm_part,
                          double norm_mesh, double norm_meshes) {
  std::fprintf(f, "%s Box size: [%.3f, %.3f, %.3f]\n", comment_delimiter,
               p.boxsize[0], p.boxsize[1], p.boxsize[2]);
  std::fprintf(f, "%s Box alignment: %s\n", comment_delimiter, p.alignment.c_str());
  std::fprintf(f, "%s Mesh number: [%d, %d, %d]\n", comment_delimiter,
               p.ngrid[0], p.ngrid[1], p.ngrid[2]);
  std::fprintf(f, "%s Mesh assignment and interlacing: %s, %s\n", comment_delimiter,
               p.assignment.c_str(), p.interlace.c_str());
  const struct { const char* name; double value; } conventions[] = {
    {"none", 1.}, {"particle", norm_part}, {"mesh", norm_mesh}, {"mesh-mixed", norm_meshes}};
  for (const auto& c : conventions) {
    if (p.norm_convention == c.name) {
      std::fprintf(f, "%s Normalisation factor: %.9e (%s)\n", comment_delimiter, c.value,
                   p.norm_convention.c_str());
    }
  }
  std::fprintf(f, "%s Normalisation factor alternatives: "
               "%.9e (particle), %.9e (mesh), %.9e (mesh-mixed)\n",
               comment_delimiter, norm_part, norm_mesh, norm_meshes);
}

typedef std::vector< std::complex<double> > cvec;

/// Two-point table: centre, effective coordinate, count, then one or two complex columns.
void table_2pt(std::FILE* f, int dim, const std::vector<double>& cen,
               const std::vector<double>& eff, const std::vector<int>& count,
               const cvec& first, const cvec* second) {
  for (int i = 0; i < dim; i++) {
    std::fprintf(f, "%.9e\t%.9e\t%10d\t% .9e\t% .9e", cen[i], eff[i], count[i],
                 first[i].real(), first[i].imag());
    if (second) std::fprintf(f, "\t% .9e\t% .9e", (*second)[i].real(), (*second)[i].imag());
    std::fprintf(f, "\n");
  }
}

/// Three-point table: two (centre, effective coordinate, count) triplets, raw and shot.
void table_3pt(std::FILE* f, trv::ParameterSet& p, const char* coord, const char* count_name,
               const char* stat, int dim,
               const std::vector<double>& c1, const std::vector<double>& e1, const std::vector<int>& n1,
               const std::vector<double>& c2, const std::vector<double>& e2, const std::vector<int>& n2,
               const cvec& raw, const cvec& shot) {
  char mp[16];
  std::snprintf(mp, sizeof(mp), "%d%d%d", p.ell1, p.ell2, p.ELL);
  std::fprintf(f, "%s [0] %s1_cen, [1] %s1_eff, [2] %s_1, [3] %s2_cen, [4] %s2_eff, [5] %s_2, "
               "[6] Re{%s%s_raw}, [7] Im{%s%s_raw}, [8] Re{%s%s_shot}, [9] Im{%s%s_shot}\n",
               comment_delimiter, coord, coord, count_name, coord, coord, count_name,
               stat, mp, stat, mp, stat, mp, stat, mp);
  for (int i = 0; i < dim; i++) {
    std::fprintf(f, "%.9e\t%.9e\t%10d\t%.9e\t%.9e\t%10d\t% .9e\t% .9e\t% .9e\t% .9e\n",
                 c1[i], e1[i], n1[i], c2[i], e2[i], n2[i],
                 raw[i].real(), raw[i].imag(), shot[i].real(), shot[i].imag());
  }
}

}  // namespace

void print_measurement_header_to_file(
  std::FILE* fileptr, trv::ParameterSet& params,
  trv::ParticleCatalogue& catalogue_data, trv::ParticleCatalogue& catalogue_rand,
  double norm_factor_part, double norm_factor_mesh, double norm_factor_meshes
) {
  header_catalogue(fileptr, "Data catalogue source", "Data catalogue size",
                   "Data-source particle extents", catalogue_data);
  header_catalogue(fileptr, "Random catalogue source", "Random catalogue size",
                   "Random-source particle extents", catalogue_rand);
  header_mesh_and_norm(fileptr, params, norm_factor_part, norm_factor_mesh, norm_factor_meshes);
}

void print_measurement_header_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ParticleCatalogue& catalogue,
  double norm_factor_part, double norm_factor_mesh, double norm_factor_meshes
) {
  header_catalogue(fileptr, "Catalogue source", "Catalogue size", "Catalogue particle extents",
                   catalogue);
  header_mesh_and_norm(fileptr, params, norm_factor_part, norm_factor_mesh, norm_factor_meshes);
}

void print_binned_vectors_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::BinnedVectors& binned_vectors
) {
  std::fprintf(fileptr, "%s Box size: [%.3f, %.3f, %.3f]\n", comment_delimiter,
               params.boxsize[0], params.boxsize[1], params.boxsize[2]);
  std::fprintf(fileptr, "%s Mesh number: [%d, %d, %d]\n", comment_delimiter,
               params.ngrid[0], params.ngrid[1], params.ngrid[2]);
  std::fprintf(fileptr, "%s Vector count: %d\n", comment_delimiter, binned_vectors.count);
  std::fprintf(fileptr, "%s Bin number: %d\n", comment_delimiter, binned_vectors.num_bins);
  std::fprintf(fileptr, "%s [0] bin_index, [1] bin_lower, [2] bin_upper, "
               "[3] vec_x, [4] vec_y, [5] vec_z\n", comment_delimiter);
  for (int i = 0; i < binned_vectors.count; i++) {
    std::fprintf(fileptr, "%d\t%.9e\t%.9e\t% .9e\t% .9e\t% .9e\n",
                 binned_vectors.indices[i], binned_vectors.lower_edges[i],
                 binned_vectors.upper_edges[i], binned_vectors.vecx[i],
                 binned_vectors.vecy[i], binned_vectors.vecz[i]);
  }
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::PowspecMeasurements& m
) {
  std::fprintf(fileptr, "%s [0] k_cen, [1] k_eff, [2] nmodes, [3] Re{pk%d_raw}, [4] Im{pk%d_raw}, "
               "[5] Re{pk%d_shot}, [6] Im{pk%d_shot}\n", comment_delimiter,
               params.ELL, params.ELL, params.ELL, params.ELL);
  table_2pt(fileptr, m.dim, m.kbin, m.keff, m.nmodes, m.pk_raw, &m.pk_shot);
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::TwoPCFMeasurements& m
) {
  std::fprintf(fileptr, "%s [0] r_cen, [1] r_eff, [2] npairs, [3] Re{xi%d}, [4] Im{xi%d}\n",
               comment_delimiter, params.ELL, params.ELL);
  table_2pt(fileptr, m.dim, m.rbin, m.reff, m.npairs, m.xi, nullptr);
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::TwoPCFWindowMeasurements& m
) {
  std::fprintf(fileptr, "%s [0] r_cen, [1] r_eff, [2] npairs, [3] Re{xi%d}, [4] Im{xi%d}\n",
               comment_delimiter, params.ELL, params.ELL);
  table_2pt(fileptr, m.dim, m.rbin, m.reff, m.npairs, m.xi, nullptr);
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::BispecMeasurements& m
) {
  table_3pt(fileptr, params, "k", "nmodes", "bk", m.dim, m.k1_bin, m.k1_eff, m.nmodes_1,
            m.k2_bin, m.k2_eff, m.nmodes_2, m.bk_raw, m.bk_shot);
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ThreePCFMeasurements& m
) {
  table_3pt(fileptr, params, "r", "npairs", "zeta", m.dim, m.r1_bin, m.r1_eff, m.npairs_1,
            m.r2_bin, m.r2_eff, m.npairs_2, m.zeta_raw, m.zeta_shot);
}

void print_measurement_datatab_to_file(
  std::FILE* fileptr, trv::ParameterSet& params, trv::ThreePCFWindowMeasurements& m
) {
  table_3pt(fileptr, params, "r", "npairs", "zeta", m.dim, m.r1_bin, m.r1_eff, m.npairs_1,
            m.r2_bin, m.r2_eff, m.npairs_2, m.zeta_raw, m.zeta_shot);
}

}  // namespace io
}  // namespace trv
