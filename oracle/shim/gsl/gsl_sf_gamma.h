/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call site: maths.cpp:124 (FFTLog only). */
#ifndef TRV_ORACLE_SHIM_GSL_SF_GAMMA_H_
#define TRV_ORACLE_SHIM_GSL_SF_GAMMA_H_
#include "gsl_sf_result.h"
#ifdef __cplusplus
extern "C" {
#endif
int gsl_sf_lngamma_complex_e(
  double zr, double zi, gsl_sf_result* lnr, gsl_sf_result* arg
);
#ifdef __cplusplus
}
#endif
#endif
