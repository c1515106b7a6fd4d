/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call sites: maths.cpp:331,370. */
#ifndef TRV_ORACLE_SHIM_GSL_SF_BESSEL_H_
#define TRV_ORACLE_SHIM_GSL_SF_BESSEL_H_
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_bessel_jl(const int l, const double x);
#ifdef __cplusplus
}
#endif
#endif
