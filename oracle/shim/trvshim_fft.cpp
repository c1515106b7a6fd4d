// TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
//
// Implementation of the FFTW3 subset declared in oracle/shim/fftw3.h so that
// the UNMODIFIED reference sources (/root/reference/src/triumvirate/src/*.cpp)
// can be compiled and linked in an image without FFTW3.  Convention matches
// FFTW: unnormalised, row-major, sign -1 forward / +1 backward; in-place and
// out-of-place both supported.  Algorithm: 1-D mixed-radix Cooley-Tukey
// (radix 4/2 fast paths, generic small primes, O(n^2) DFT for a large prime
// factor) with an exactly tabulated twiddle table; the 3-D transform applies it
// line by line along each axis (OpenMP over lines).  Round-off differs from
// FFTW's by O(1e-16 log n) relative, far below the parity tolerance.

#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::complex<double> cplx;

struct Fft1d {
  int n = 0;
  int sign = -1;
  std::vector<cplx> tw;       // tw[k] = exp(sign * 2 pi i k / n)
  std::vector<int> factors;   // radix sequence

  void init(int n_, int sign_) {
    n = n_; sign = sign_;
    tw.resize(n > 0 ? n : 1);
    const long double twopi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; k++) {
      long double a = twopi * (long double)k / (long double)n;
      tw[k] = cplx((double)cosl(a), (double)(sign * sinl(a)));
    }
    factors.clear();
    int m = n;
    while (m % 4 == 0) { factors.push_back(4); m /= 4; }
    while (m % 2 == 0) { factors.push_back(2); m /= 2; }
    for (int p = 3; (long long)p * p <= m; p += 2) {
      while (m % p == 0) { factors.push_back(p); m /= p; }
    }
    if (m > 1) factors.push_back(m);
  }

  // Recursive decimation-in-time: out[0..n_sub) = DFT of in[0], in[stride], ...
  void rec(const cplx* in, cplx* out, int n_sub, int stride, int level,
           cplx* scratch) const {
    if (n_sub == 1) { out[0] = in[0]; return; }
    const int p = factors[level];
    const int m = n_sub / p;
    for (int q = 0; q < p; q++) {
      rec(in + (size_t)q * stride, out + (size_t)q * m, m, stride * p,
          level + 1, scratch);
    }
    // Butterflies: X[k + m r] = sum_q W_n_sub^{q (k + m r)} Y_q[k].
    const int tstep = n / n_sub;   // twiddle stride into the master table
    if (p == 2) {
      for (int k = 0; k < m; k++) {
        cplx a = out[k];
        cplx b = out[k + m] * tw[(size_t)k * tstep];
        out[k] = a + b;
        out[k + m] = a - b;
      }
    } else if (p == 4) {
      const cplx jj = (sign < 0) ? cplx(0., -1.) : cplx(0., 1.);
      for (int k = 0; k < m; k++) {
        cplx a0 = out[k];
        cplx a1 = out[k + m] * tw[(size_t)k * tstep];
        cplx a2 = out[k + 2 * m] * tw[(size_t)2 * k * tstep];
        cplx a3 = out[k + 3 * m] * tw[(size_t)3 * k * tstep];
        cplx s02 = a0 + a2, d02 = a0 - a2;
        cplx s13 = a1 + a3, d13 = (a1 - a3) * jj;
        out[k] = s02 + s13;
        out[k + m] = d02 + d13;
        out[k + 2 * m] = s02 - s13;
        out[k + 3 * m] = d02 - d13;
      }
    } else {
      // Generic radix p (scratch holds p values).
      for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
          scratch[q] = out[k + (size_t)q * m]
            * tw[((size_t)q * k * tstep) % (size_t)n];
        }
        for (int r = 0; r < p; r++) {
          cplx acc(0., 0.);
          for (int q = 0; q < p; q++) {
            // W_p^{q r} = tw[(q r mod p) * n / p]
            acc += scratch[q] * tw[(size_t)((q * (long long)r) % p) * (n / p)];
          }
          out[k + (size_t)r * m] = acc;
        }
      }
    }
  }

  // Transform contiguous buffer `buf` (length n) using `work` (length n).
  void run(cplx* buf, cplx* work, cplx* scratch) const {
    if (n <= 1) return;
    rec(buf, work, n, 1, 0, scratch);
    std::memcpy((void*)buf, (const void*)work, sizeof(cplx) * (size_t)n);
  }
};

}  // namespace

struct trvshim_fftw_plan_s {
  int rank;
  int n[3];
  int sign;
  fftw_complex* in;
  fftw_complex* out;
  Fft1d ax[3];
};

extern "C" {

const char fftw_version[] = "trvshim-fft-3.3-compatible";

void* fftw_malloc(size_t n) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, n > 0 ? n : 64) != 0) return nullptr;
  return p;
}

fftw_complex* fftw_alloc_complex(size_t n) {
  return (fftw_complex*)fftw_malloc(sizeof(fftw_complex) * n);
}

void fftw_free(void* p) { std::free(p); }

fftw_plan fftw_plan_dft_3d(
  int n0, int n1, int n2, fftw_complex* in, fftw_complex* out,
  int sign, unsigned /*flags*/
) {
  trvshim_fftw_plan_s* p = new trvshim_fftw_plan_s();
  p->rank = 3;
  p->n[0] = n0; p->n[1] = n1; p->n[2] = n2;
  p->sign = sign; p->in = in; p->out = out;
  p->ax[0].init(n0, sign); p->ax[1].init(n1, sign); p->ax[2].init(n2, sign);
  return p;
}

fftw_plan fftw_plan_dft_1d(
  int n, fftw_complex* in, fftw_complex* out, int sign, unsigned /*flags*/
) {
  trvshim_fftw_plan_s* p = new trvshim_fftw_plan_s();
  p->rank = 1;
  p->n[0] = 1; p->n[1] = 1; p->n[2] = n;
  p->sign = sign; p->in = in; p->out = out;
  p->ax[0].init(1, sign); p->ax[1].init(1, sign); p->ax[2].init(n, sign);
  return p;
}

void fftw_execute_dft(const fftw_plan p, fftw_complex* in, fftw_complex* out) {
  const long long n0 = p->n[0], n1 = p->n[1], n2 = p->n[2];
  const long long ntot = n0 * n1 * n2;
  cplx* a = reinterpret_cast<cplx*>(out);
  if (in != out) {
    std::memcpy((void*)out, (const void*)in, sizeof(fftw_complex) * (size_t)ntot);
  }
  const int nmax = (int)std::max(n0, std::max(n1, n2));

#ifdef _OPENMP
#pragma omp parallel
#endif
  {
    const int B = 8;  // lines gathered together along strided axes
    std::vector<cplx> line((size_t)nmax * B), work((size_t)nmax), scr(64 + nmax);

    // Axis 2 (contiguous).
    if (n2 > 1) {
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
      for (long long ij = 0; ij < n0 * n1; ij++) {
        p->ax[2].run(a + ij * n2, work.data(), scr.data());
      }
    }
    // Axis 1 (stride n2): gather B adjacent k-columns at a time.
    if (n1 > 1) {
      const long long nkb = (n2 + B - 1) / B;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
      for (long long t = 0; t < n0 * nkb; t++) {
        const long long i = t / nkb, k0 = (t % nkb) * B;
        const int nb = (int)std::min<long long>(B, n2 - k0);
        cplx* base = a + i * n1 * n2 + k0;
        for (long long j = 0; j < n1; j++)
          for (int b = 0; b < nb; b++)
            line[(size_t)b * n1 + j] = base[j * n2 + b];
        for (int b = 0; b < nb; b++)
          p->ax[1].run(line.data() + (size_t)b * n1, work.data(), scr.data());
        for (long long j = 0; j < n1; j++)
          for (int b = 0; b < nb; b++)
            base[j * n2 + b] = line[(size_t)b * n1 + j];
      }
    }
    // Axis 0 (stride n1*n2).
    if (n0 > 1) {
      const long long plane = n1 * n2;
      const long long nkb = (plane + B - 1) / B;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
      for (long long t = 0; t < nkb; t++) {
        const long long k0 = t * B;
        const int nb = (int)std::min<long long>(B, plane - k0);
        cplx* base = a + k0;
        for (long long i = 0; i < n0; i++)
          for (int b = 0; b < nb; b++)
            line[(size_t)b * n0 + i] = base[i * plane + b];
        for (int b = 0; b < nb; b++)
          p->ax[0].run(line.data() + (size_t)b * n0, work.data(), scr.data());
        for (long long i = 0; i < n0; i++)
          for (int b = 0; b < nb; b++)
            base[i * plane + b] = line[(size_t)b * n0 + i];
      }
    }
  }
}

void fftw_execute(const fftw_plan p) { fftw_execute_dft(p, p->in, p->out); }

void fftw_destroy_plan(fftw_plan p) { delete p; }

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}
void fftw_cleanup(void) {}

int fftw_import_wisdom_from_filename(const char*) { return 0; }
int fftw_export_wisdom_to_filename(const char*) { return 0; }

}  // extern "C"
