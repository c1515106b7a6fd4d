// maths.hpp -- trv::maths calculators (mirror of I/maths.hpp; S/maths.cpp:
// 167-375) with GSL replaced by self-contained evaluations.
#ifndef TRV_B200_MATHS_HPP_
#define TRV_B200_MATHS_HPP_

#include <complex>
#include <vector>

#include "monitor.hpp"

namespace trv {
namespace maths {

extern const std::complex<double> M_I;
extern const double eps_coupling;  // 1e-9, S/maths.cpp:165

double get_vec3d_magnitude(const double* vec);

/// Wigner 3-j symbol (S/maths.cpp:167-169; Racah formula).
double wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3);

/// Spherical Bessel function of the first kind, accurate to ~1 ulp.
double sph_bessel_jl(int ell, double x);

/// sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(x), Condon-Shortley phase included.
double sph_legendre_plm(int ell, int m, double x);

class SphericalHarmonicCalculator {
 public:
  /// Reduced spherical harmonic (S/maths.cpp:171-220).
  static std::complex<double> calc_reduced_spherical_harmonic(
    const int ell, const int m, double pos[3]
  );
  /// y_lm at the wavevector of every mesh cell, signed FFT index order, index
  /// (i n_y + j) n_z + k (S/maths.cpp:222-261).  The device estimators never build these
  /// tables (y_lm is evaluated in registers); they are kept for C++ callers of the
  /// reference API.  Throws trv::sys::InvalidDataError when `ylm_out` is not sized n_x n_y n_z.
  static void store_reduced_spherical_harmonic_in_fourier_space(
    const int ell, const int m,
    const double boxsize[3], const int ngrid[3],
    std::vector< std::complex<double> >& ylm_out
  );
  /// The same at the signed position vector of every cell (S/maths.cpp:263-302).
  static void store_reduced_spherical_harmonic_in_config_space(
    const int ell, const int m,
    const double boxsize[3], const int ngrid[3],
    std::vector< std::complex<double> >& ylm_out
  );
};

/// Interpolated spherical Bessel function (I/maths.hpp:262-312; S/maths.cpp:
/// 309-375): natural cubic spline with step 0.05 on [0, max(1000, ell^2)],
/// direct evaluation above.  The (y, c) table is what the device evaluates.
class SphericalBesselCalculator {
 public:
  int order;
  double split = 1000.;
  double step = 0.05;
  std::vector<double> x;   ///< knots
  std::vector<double> y;   ///< j_ell at the knots
  std::vector<double> c;   ///< spline coefficients (half second derivatives)

  explicit SphericalBesselCalculator(const int ell);
  SphericalBesselCalculator(const SphericalBesselCalculator& other) = default;
  SphericalBesselCalculator& operator=(const SphericalBesselCalculator& other) = default;
  double eval(double x);

 private:
  void build_table();
};

}  // namespace maths
}  // namespace trv

#endif  // TRV_B200_MATHS_HPP_
