// maths.cpp -- trv::maths calculators without GSL.
//
// The reference delegates to GSL (S/maths.cpp:168 gsl_sf_coupling_3j, :211
// gsl_sf_legendre_sphPlm, :331/:370 gsl_sf_bessel_jl, :335-373 gsl cspline).
// Here each is evaluated from its definition; the spline reproduces GSL's
// natural-cspline convention because the reference's j_l values BETWEEN the
// knots are spline values (interpolation error ~1e-8 absolute), so an exact
// j_l on the device would fail 1e-8 parity.
#include "trv/maths.hpp"

#include <map>
#include <mutex>

#include <cmath>

namespace trv {
namespace maths {

const std::complex<double> M_I(0., 1.);
const double eps_coupling = 1.e-9;

double get_vec3d_magnitude(const double* vec) {
  return std::sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
}

// ---------------------------------------------------------------------
// Wigner 3-j (Racah's single-sum formula, exact in long double for the
// degrees used here).
// ---------------------------------------------------------------------

namespace {

long double fact(int n) {
  long double f = 1.L;
  for (int i = 2; i <= n; i++) f *= i;
  return f;
}

}  // namespace

double wigner_3j(int j1, int j2, int j3, int m1, int m2, int m3) {
  if (j1 < 0 || j2 < 0 || j3 < 0) return 0.;
  if (m1 + m2 + m3 != 0) return 0.;
  if (std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m3) > j3) return 0.;
  if (j3 > j1 + j2 || j3 < std::abs(j1 - j2)) return 0.;
  const int t1 = j2 - m1 - j3, t2 = j1 + m2 - j3;
  const int t3 = j1 + j2 - j3, t4 = j1 - m1, t5 = j2 + m2;
  const int tmin = std::max(0, std::max(t1, t2));
  const int tmax = std::min(t3, std::min(t4, t5));
  long double sum = 0.L;
  for (int t = tmin; t <= tmax; t++) {
    long double den = fact(t) * fact(t - t1) * fact(t - t2)
      * fact(t3 - t) * fact(t4 - t) * fact(t5 - t);
    sum += ((t & 1) ? -1.L : 1.L) / den;
  }
  long double tri = fact(j1 + j2 - j3) * fact(j1 - j2 + j3) * fact(-j1 + j2 + j3)
    / fact(j1 + j2 + j3 + 1);
  long double pref = sqrtl(
    tri * fact(j1 + m1) * fact(j1 - m1) * fact(j2 + m2) * fact(j2 - m2)
    * fact(j3 + m3) * fact(j3 - m3));
  const int ph = j1 - j2 - m3;
  return (double)(((std::abs(ph) & 1) ? -1.L : 1.L) * pref * sum);
}

// ---------------------------------------------------------------------
// Spherical Bessel j_l.
// ---------------------------------------------------------------------

double sph_bessel_jl(int ell, double x_) {
  long double x = x_;
  if (x == 0.L) return ell == 0 ? 1. : 0.;
  if (x < 0.L) return ((ell & 1) ? -1. : 1.) * sph_bessel_jl(ell, -x_);
  if (x >= (long double)ell && x >= 1.L) {
    // Forward recurrence from the closed forms, stable for x >= l.
    const long double s = sinl(x), c = cosl(x);
    long double jm = s / x;
    if (ell == 0) return (double)jm;
    long double jc = (s / x - c) / x;
    for (int n = 1; n < ell; n++) {
      const long double jn = (2 * n + 1) / x * jc - jm;
      jm = jc; jc = jn;
    }
    return (double)jc;
  }
  // Miller's backward recurrence, normalised with j_0 = sin(x)/x (x < l) or
  // by the power series when x is tiny.
  if (x < 1.e-2L) {
    // Leading terms of the ascending series suffice: relative error
    // O(x^6 / l^3) < 1e-19 here.
    long double pre = 1.L;
    for (int i = 1; i <= ell; i++) pre *= x / (2 * i + 1);
    const long double h = -0.5L * x * x;
    long double term = 1.L, sum = 1.L;
    for (int k = 1; k < 8; k++) {
      term *= h / (k * (2.L * ell + 2 * k + 1));
      sum += term;
    }
    return (double)(pre * sum);
  }
  // The trial sequence is normalised with the sum rule
  // sum_n (2n+1) j_n(x)^2 = 1 (robust where j_0 or j_1 vanish); the sign is
  // fixed against whichever of j_0, j_1 is larger in magnitude.
  const int nstart = ell + 60 + (int)(2 * x);
  long double jp1 = 0.L, jc = 1.e-300L, target = 0.L;
  long double sumsq = (2.L * nstart + 1.L) * jc * jc;
  long double t0 = 0.L, t1 = 0.L;
  for (int n = nstart; n >= 1; n--) {
    const long double jm1 = (2 * n + 1) / x * jc - jp1;
    jp1 = jc; jc = jm1;
    sumsq += (2.L * (n - 1) + 1.L) * jc * jc;
    if (n - 1 == ell) target = jc;
    if (fabsl(jc) > 1.e300L) {
      jc *= 1.e-300L; jp1 *= 1.e-300L; target *= 1.e-300L; sumsq *= 1.e-600L;
    }
  }
  t0 = jc; t1 = jp1;   // unnormalised j_0, j_1
  const long double s = sinl(x), c = cosl(x);
  const long double j0 = s / x, j1 = (s / x - c) / x;
  long double sign;
  if (fabsl(j0) >= fabsl(j1)) sign = ((j0 < 0.L) == (t0 < 0.L)) ? 1.L : -1.L;
  else sign = ((j1 < 0.L) == (t1 < 0.L)) ? 1.L : -1.L;
  return (double)(sign * target / sqrtl(sumsq));
}

// ---------------------------------------------------------------------
// Normalised associated Legendre function and reduced harmonics.
// ---------------------------------------------------------------------

double sph_legendre_plm(int ell, int m, double x_) {
  if (m < 0 || m > ell) return 0.;
  const long double x = x_;
  const long double pi = 3.141592653589793238462643383279502884L;
  const long double sx = sqrtl((1.L - x) * (1.L + x));
  long double pmm = sqrtl(1.L / (4.L * pi));
  for (int i = 1; i <= m; i++) pmm *= -sqrtl((2.L * i + 1.L) / (2.L * i)) * sx;
  if (ell == m) return (double)pmm;
  long double pm1 = x * sqrtl(2.L * m + 3.L) * pmm;
  for (int l = m + 2; l <= ell; l++) {
    const long double a = sqrtl((4.L * l * l - 1.L) / ((long double)l * l - (long double)m * m));
    const long double b = sqrtl((((long double)l - 1) * (l - 1) - (long double)m * m)
                                / (4.L * (l - 1) * (l - 1) - 1.L));
    const long double p = a * (x * pm1 - b * pmm);
    pmm = pm1; pm1 = p;
  }
  return (double)pm1;
}

std::complex<double> SphericalHarmonicCalculator::calc_reduced_spherical_harmonic(
  const int ell, const int m, double pos[3]
) {
  const double eps = 1.e-9;   // S/maths.cpp:176
  if (ell == 0 && m == 0) return 1.;
  double r2 = 0.;
  for (int ax = 0; ax < 3; ax++) r2 += pos[ax] * pos[ax];
  const double r = std::sqrt(r2);
  if (std::fabs(r) < eps) return 0.;
  const double mu = pos[2] / r;
  const double rxy = std::sqrt(pos[0] * pos[0] + pos[1] * pos[1]);
  double phi = 0.;
  if (std::fabs(rxy) >= eps) {
    phi = std::acos(pos[0] / rxy);
    if (pos[1] < 0.) phi = -phi + 2. * M_PI;
  }
  std::complex<double> ylm = std::exp(M_I * double(m) * phi)
    * sph_legendre_plm(ell, std::abs(m), mu);
  ylm = std::pow(-1, (m - std::abs(m)) / 2) * std::conj(ylm);
  ylm *= std::sqrt(4. * M_PI / (2. * ell + 1.));
  return ylm;
}

namespace {

/// Fill `out` with y_lm of (signed index) x (step per axis) for every mesh cell.
void store_ylm_on_mesh(int ell, int m, const double step[3], const int ngrid[3],
                       std::vector< std::complex<double> >& out) {
  const long long nmesh = (long long)ngrid[0] * ngrid[1] * ngrid[2];
  if ((long long)out.size() < nmesh) {
    throw trv::sys::InvalidDataError(
      "Output vector for reduced spherical harmonics holds %zu of %lld mesh cells.",
      out.size(), nmesh);
  }
#pragma omp parallel for collapse(2)
  for (int i = 0; i < ngrid[0]; i++) {
    for (int j = 0; j < ngrid[1]; j++) {
      double v[3];
      v[0] = ((i < ngrid[0] / 2) ? i : i - ngrid[0]) * step[0];
      v[1] = ((j < ngrid[1] / 2) ? j : j - ngrid[1]) * step[1];
      std::complex<double>* row = out.data() + ((long long)i * ngrid[1] + j) * ngrid[2];
      for (int k = 0; k < ngrid[2]; k++) {
        v[2] = ((k < ngrid[2] / 2) ? k : k - ngrid[2]) * step[2];
        row[k] = SphericalHarmonicCalculator::calc_reduced_spherical_harmonic(ell, m, v);
      }
    }
  }
}

}  // namespace

void SphericalHarmonicCalculator::store_reduced_spherical_harmonic_in_fourier_space(
  const int ell, const int m, const double boxsize[3], const int ngrid[3],
  std::vector< std::complex<double> >& ylm_out
) {
  const double dk[3] = {2. * M_PI / boxsize[0], 2. * M_PI / boxsize[1], 2. * M_PI / boxsize[2]};
  store_ylm_on_mesh(ell, m, dk, ngrid, ylm_out);
}

void SphericalHarmonicCalculator::store_reduced_spherical_harmonic_in_config_space(
  const int ell, const int m, const double boxsize[3], const int ngrid[3],
  std::vector< std::complex<double> >& ylm_out
) {
  const double dr[3] = {boxsize[0] / double(ngrid[0]), boxsize[1] / double(ngrid[1]),
                        boxsize[2] / double(ngrid[2])};
  store_ylm_on_mesh(ell, m, dr, ngrid, ylm_out);
}

// ---------------------------------------------------------------------
// Spline-interpolated spherical Bessel calculator.
// ---------------------------------------------------------------------

namespace {
// The table is a pure function of ell: build it once per process (the
// reference rebuilds it -- 20001 exact j_l evaluations and a tridiagonal
// solve -- in every estimator call, S/threept.cpp:1560-1561).
std::mutex g_sjl_mutex;
std::map<int, SphericalBesselCalculator> g_sjl_cache;
}  // namespace

SphericalBesselCalculator::SphericalBesselCalculator(const int ell) : order(ell) {
  {
    std::lock_guard<std::mutex> lock(g_sjl_mutex);
    auto it = g_sjl_cache.find(ell);
    if (it != g_sjl_cache.end()) { *this = it->second; return; }
  }
  this->build_table();
  std::lock_guard<std::mutex> lock(g_sjl_mutex);
  g_sjl_cache.emplace(ell, *this);
}

void SphericalBesselCalculator::build_table() {
  this->split = (this->split >= this->order * this->order)
    ? this->split : this->order * this->order;            // S/maths.cpp:313-314
  const double xmin = 0., xmax = this->split, dx = this->step;
  const int nsample = int((xmax - xmin) / dx) + 1;         // S/maths.cpp:320
  this->x.resize(nsample);
  this->y.resize(nsample);
  this->c.assign(nsample, 0.);
#pragma omp parallel for
  for (int i = 0; i < nsample; i++) {
    this->x[i] = xmin + dx * i;
    this->y[i] = sph_bessel_jl(this->order, this->x[i]);
  }
  // Natural cubic spline: c_0 = c_{n-1} = 0 and, for the interior,
  //   h_{i-1} c_{i-1} + 2 (h_{i-1} + h_i) c_i + h_i c_{i+1}
  //     = 3 [ (y_{i+1} - y_i)/h_i - (y_i - y_{i-1})/h_{i-1} ],
  // solved by forward elimination / back substitution.
  const int n = nsample;
  if (n < 3) return;
  const int m = n - 2;
  std::vector<double> diag(m), off(m), rhs(m);
  for (int i = 0; i < m; i++) {
    const double h0 = this->x[i + 1] - this->x[i];
    const double h1 = this->x[i + 2] - this->x[i + 1];
    const double d0 = this->y[i + 1] - this->y[i];
    const double d1 = this->y[i + 2] - this->y[i + 1];
    off[i] = h1;
    diag[i] = 2.0 * (h1 + h0);
    rhs[i] = 3.0 * (d1 * (1.0 / h1) - d0 * (1.0 / h0));
  }
  std::vector<double> gam(m), z(m);
  gam[0] = off[0] / diag[0];
  z[0] = rhs[0] / diag[0];
  for (int i = 1; i < m; i++) {
    const double den = diag[i] - off[i - 1] * gam[i - 1];
    gam[i] = off[i] / den;
    z[i] = (rhs[i] - off[i - 1] * z[i - 1]) / den;
  }
  this->c[m] = z[m - 1];
  for (int i = m - 2; i >= 0; i--) this->c[i + 1] = z[i] - gam[i] * this->c[i + 2];
}

double SphericalBesselCalculator::eval(double xv) {
  if (xv >= this->split) return sph_bessel_jl(this->order, xv);
  const int n = static_cast<int>(this->x.size());
  int i = static_cast<int>(xv / this->step);
  if (i > n - 2) i = n - 2;
  if (i < 0) i = 0;
  while (i < n - 2 && this->x[i + 1] <= xv) i++;
  while (i > 0 && this->x[i] > xv) i--;
  const double x_lo = this->x[i], x_hi = this->x[i + 1];
  const double y_lo = this->y[i], y_hi = this->y[i + 1];
  const double dx = x_hi - x_lo, dy = y_hi - y_lo;
  const double b_i = (dy / dx) - dx * (this->c[i + 1] + 2.0 * this->c[i]) / 3.0;
  const double d_i = (this->c[i + 1] - this->c[i]) / (3.0 * dx);
  const double delx = xv - x_lo;
  return y_lo + delx * (b_i + delx * (this->c[i] + delx * d_i));
}

}  // namespace maths
}  // namespace trv
