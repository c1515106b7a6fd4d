// TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
//
// Implementation of the GSL subset declared in oracle/shim/gsl/*.h so that the
// UNMODIFIED reference sources can be compiled in an image without GSL
// (dependency: gsl >= 2.7, deploy/pkg/conda_recipe/meta.yaml:40; call sites
// S/maths.cpp:124,168,211,331-338,350-373).  Each function restates the
// PUBLISHED mathematical definition of the GSL routine, evaluated in long
// double so the double result is correct to ~1 ulp:
//   gsl_sf_bessel_jl        spherical Bessel function of the first kind
//   gsl_sf_legendre_sphPlm  sqrt((2l+1)/(4pi) (l-m)!/(l+m)!) P_l^m(x), with
//                           the Condon-Shortley phase
//   gsl_sf_coupling_3j      Wigner 3-j symbol, arguments doubled
//   gsl_interp_cspline      natural cubic spline (c_0 = c_{n-1} = 0), GSL's
//                           coefficient convention and evaluation formula
//   gsl_sf_lngamma_complex_e  log-gamma for complex argument (FFTLog only)

#include "gsl/gsl_interp.h"
#include "gsl/gsl_sf_bessel.h"
#include "gsl/gsl_sf_coupling.h"
#include "gsl/gsl_sf_gamma.h"
#include "gsl/gsl_sf_legendre.h"
#include "gsl/gsl_spline.h"
#include "gsl/gsl_version.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

namespace {

typedef long double ld;

const ld PI_L = 3.141592653589793238462643383279502884L;

ld sph_bessel_series(int l, ld x) {
  // j_l(x) = x^l / (2l+1)!! * sum_k (-x^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1))
  ld pre = 1.L;
  for (int i = 1; i <= l; i++) pre *= x / (ld)(2 * i + 1);
  ld term = 1.L, sum = 1.L;
  const ld h = -0.5L * x * x;
  for (int k = 1; k < 500; k++) {
    term *= h / ((ld)k * (ld)(2 * l + 2 * k + 1));
    sum += term;
    if (fabsl(term) < 1e-22L * fabsl(sum)) break;
  }
  return pre * sum;
}

ld sph_bessel(int l, ld x) {
  if (x == 0.L) return (l == 0) ? 1.L : 0.L;
  if (x < 0.L) {
    ld v = sph_bessel(l, -x);
    return (l % 2 == 0) ? v : -v;
  }
  // Ascending series in the monotone region (stable: little cancellation for
  // x below the first turning point ~ l + 1/2).
  if (x < 1.L || x < (ld)l) return sph_bessel_series(l, x);
  // Upward recurrence is stable for x >= l.
  ld s = sinl(x), c = cosl(x);
  ld j0 = s / x;
  if (l == 0) return j0;
  ld j1 = (s / x - c) / x;
  if (l == 1) return j1;
  ld jm = j0, jc = j1;
  for (int n = 1; n < l; n++) {
    ld jn = (ld)(2 * n + 1) / x * jc - jm;
    jm = jc; jc = jn;
  }
  return jc;
}

ld factorial_l(int n) {
  ld f = 1.L;
  for (int i = 2; i <= n; i++) f *= (ld)i;
  return f;
}

}  // namespace

extern "C" {

const char* gsl_version = GSL_VERSION;

double gsl_sf_bessel_jl(const int l, const double x) {
  return (double)sph_bessel(l, (ld)x);
}

double gsl_sf_legendre_sphPlm(const int l, const int m, const double x) {
  // Requires 0 <= m <= l, |x| <= 1.
  if (m < 0 || l < m) return 0.;
  ld xl = (ld)x;
  ld somx2 = sqrtl((1.L - xl) * (1.L + xl));
  // Normalised P_m^m: N_mm P_mm = (-1)^m sqrt((2m+1)!!/(2m)!!/(4pi)) (1-x^2)^{m/2}
  ld pmm = sqrtl(1.L / (4.L * PI_L));
  for (int i = 1; i <= m; i++) {
    pmm *= -sqrtl((ld)(2 * i + 1) / (ld)(2 * i)) * somx2;
  }
  if (l == m) return (double)pmm;
  // Normalised upward recurrence in l.
  ld pmmp1 = xl * sqrtl((ld)(2 * m + 3)) * pmm;
  if (l == m + 1) return (double)pmmp1;
  ld pll = 0.L;
  for (int ll = m + 2; ll <= l; ll++) {
    ld a = sqrtl(((ld)(4 * ll * ll) - 1.L) / ((ld)(ll * ll) - (ld)(m * m)));
    ld b = sqrtl((((ld)(ll - 1) * (ld)(ll - 1)) - (ld)(m * m))
                 / ((ld)(4 * (ll - 1) * (ll - 1)) - 1.L));
    pll = a * (xl * pmmp1 - b * pmm);
    pmm = pmmp1; pmmp1 = pll;
  }
  return (double)pll;
}

double gsl_sf_coupling_3j(
  int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc
) {
  // Racah formula; all arguments are twice the physical values.
  if (two_ja < 0 || two_jb < 0 || two_jc < 0) return 0.;
  if (two_ma + two_mb + two_mc != 0) return 0.;
  if (std::abs(two_ma) > two_ja || std::abs(two_mb) > two_jb
      || std::abs(two_mc) > two_jc) return 0.;
  if ((two_ja + two_ma) % 2 || (two_jb + two_mb) % 2 || (two_jc + two_mc) % 2)
    return 0.;
  if ((two_ja + two_jb + two_jc) % 2) return 0.;
  if (two_jc > two_ja + two_jb || two_jc < std::abs(two_ja - two_jb)) return 0.;

  const int jca = (-two_ja + two_jb + two_jc) / 2;
  const int jcb = (two_ja - two_jb + two_jc) / 2;
  const int jcc = (two_ja + two_jb - two_jc) / 2;
  const int jmma = (two_ja - two_ma) / 2, jpma = (two_ja + two_ma) / 2;
  const int jmmb = (two_jb - two_mb) / 2, jpmb = (two_jb + two_mb) / 2;
  const int jmmc = (two_jc - two_mc) / 2, jpmc = (two_jc + two_mc) / 2;
  const int jsum = (two_ja + two_jb + two_jc) / 2;

  const int kmin = std::max(std::max(0, jpmb - jmmc), jmma - jpmc);
  const int kmax = std::min(std::min(jcc, jmma), jpmb);

  ld norm = sqrtl(
    factorial_l(jca) * factorial_l(jcb) * factorial_l(jcc)
    / factorial_l(jsum + 1)
    * factorial_l(jpma) * factorial_l(jmma)
    * factorial_l(jpmb) * factorial_l(jmmb)
    * factorial_l(jpmc) * factorial_l(jmmc)
  );
  ld sum = 0.L;
  for (int k = kmin; k <= kmax; k++) {
    ld den = factorial_l(k) * factorial_l(jcc - k) * factorial_l(jmma - k)
      * factorial_l(jpmb - k) * factorial_l(jmmc - jpmb + k)
      * factorial_l(jpmc - jmma + k);
    sum += ((k % 2) ? -1.L : 1.L) / den;
  }
  const int phase_exp = (two_ja - two_jb - two_mc) / 2;
  ld phase = (std::abs(phase_exp) % 2) ? -1.L : 1.L;
  return (double)(phase * norm * sum);
}

int gsl_sf_lngamma_complex_e(
  double zr, double zi, gsl_sf_result* lnr, gsl_sf_result* arg
) {
  // Lanczos (g = 7, n = 9) with reflection; used by FFTLog only.
  typedef std::complex<long double> cl;
  static const ld coef[9] = {
    0.99999999999980993227684700473478L, 676.520368121885098567009190444019L,
    -1259.13921672240287047156078755283L, 771.3234287776530788486528258894L,
    -176.61502916214059906584551354L, 12.507343278686904814458936853L,
    -0.13857109526572011689554707L, 9.984369578019570859563e-6L,
    1.50563273514931155834e-7L
  };
  cl z((ld)zr, (ld)zi);
  cl res;
  auto lanczos = [&](cl w) {
    w -= 1.L;
    cl x = coef[0];
    for (int i = 1; i < 9; i++) x += coef[i] / (w + (ld)i);
    cl t = w + 7.5L;
    return 0.5L * logl(2.L * PI_L) + (w + 0.5L) * std::log(t) - t + std::log(x);
  };
  if (zr < 0.5) {
    cl s = std::sin(PI_L * z);
    res = logl(PI_L) - std::log(s) - lanczos(1.L - z);
  } else {
    res = lanczos(z);
  }
  lnr->val = (double)res.real(); lnr->err = 0.;
  // Reduce the argument to (-pi, pi].
  ld th = res.imag();
  th = remainderl(th, 2.L * PI_L);
  arg->val = (double)th; arg->err = 0.;
  return 0;
}

// ---------------------------------------------------------------------
// Interpolation: natural cubic spline, GSL conventions.
// ---------------------------------------------------------------------

static const gsl_interp_type trvshim_cspline_type = {"cspline", 3};
const gsl_interp_type* gsl_interp_cspline = &trvshim_cspline_type;

gsl_interp_accel* gsl_interp_accel_alloc(void) {
  gsl_interp_accel* a = (gsl_interp_accel*)std::malloc(sizeof(gsl_interp_accel));
  a->cache = 0; a->miss_count = 0; a->hit_count = 0;
  return a;
}

void gsl_interp_accel_free(gsl_interp_accel* a) { std::free(a); }

gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t size) {
  gsl_spline* s = (gsl_spline*)std::malloc(sizeof(gsl_spline));
  s->interp = nullptr;
  s->size = size;
  s->x = (double*)std::malloc(sizeof(double) * size);
  s->y = (double*)std::malloc(sizeof(double) * size);
  s->c = (double*)std::malloc(sizeof(double) * size);
  return s;
}

int gsl_spline_init(
  gsl_spline* s, const double xa[], const double ya[], size_t size
) {
  for (size_t i = 0; i < size; i++) { s->x[i] = xa[i]; s->y[i] = ya[i]; }
  s->size = size;
  // Solve for c_i (half second derivatives) with c_0 = c_{n-1} = 0:
  //   h_{i-1} c_{i-1} + 2 (h_{i-1} + h_i) c_i + h_i c_{i+1}
  //     = 3 [ (y_{i+1}-y_i)/h_i - (y_i-y_{i-1})/h_{i-1} ].
  const size_t n = size;
  for (size_t i = 0; i < n; i++) s->c[i] = 0.;
  if (n < 3) return 0;
  const size_t m = n - 2;  // interior unknowns c_1..c_{n-2}
  std::vector<double> diag(m), off(m), rhs(m);
  for (size_t i = 0; i < m; i++) {
    const double h_i = xa[i + 1] - xa[i];
    const double h_ip1 = xa[i + 2] - xa[i + 1];
    const double ydiff_i = ya[i + 1] - ya[i];
    const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    rhs[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
  }
  // Symmetric tridiagonal solve (Thomas algorithm; the system is strictly
  // diagonally dominant).
  std::vector<double> cp(m), dp(m);
  cp[0] = off[0] / diag[0];
  dp[0] = rhs[0] / diag[0];
  for (size_t i = 1; i < m; i++) {
    const double den = diag[i] - off[i - 1] * cp[i - 1];
    cp[i] = off[i] / den;
    dp[i] = (rhs[i] - off[i - 1] * dp[i - 1]) / den;
  }
  s->c[m] = dp[m - 1];
  for (size_t i = m - 1; i-- > 0;) {
    s->c[i + 1] = dp[i] - cp[i] * s->c[i + 2];
  }
  return 0;
}

double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel* a) {
  const size_t n = s->size;
  // Locate i with x_i <= x < x_{i+1} (right end inclusive in last interval).
  size_t i = a ? a->cache : 0;
  if (!(i < n - 1 && s->x[i] <= x && x < s->x[i + 1])) {
    size_t lo = 0, hi = n - 1;
    while (hi > lo + 1) {
      size_t mid = (lo + hi) / 2;
      if (s->x[mid] > x) hi = mid; else lo = mid;
    }
    i = lo;
    if (a) a->cache = i;
  }
  const double x_lo = s->x[i], x_hi = s->x[i + 1];
  const double y_lo = s->y[i], y_hi = s->y[i + 1];
  const double dx = x_hi - x_lo;
  const double dy = y_hi - y_lo;
  const double c_i = s->c[i], c_ip1 = s->c[i + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  const double delx = x - x_lo;
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

void gsl_spline_free(gsl_spline* s) {
  if (!s) return;
  std::free(s->x); std::free(s->y); std::free(s->c); std::free(s);
}

}  // extern "C"
