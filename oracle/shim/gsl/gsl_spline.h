/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call sites: maths.cpp:336-338,351-364,373.
 * The reference's copy constructor reads spline->x, ->y, ->size directly. */
#ifndef TRV_ORACLE_SHIM_GSL_SPLINE_H_
#define TRV_ORACLE_SHIM_GSL_SPLINE_H_
#include <stddef.h>
#include "gsl_interp.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  void* interp;   /* unused */
  double* x;
  double* y;
  size_t size;
  double* c;      /* natural-spline second-derivative-like coefficients */
} gsl_spline;
gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t size);
int gsl_spline_init(gsl_spline* spline, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline* spline, double x, gsl_interp_accel* a);
void gsl_spline_free(gsl_spline* spline);
#ifdef __cplusplus
}
#endif
#endif
