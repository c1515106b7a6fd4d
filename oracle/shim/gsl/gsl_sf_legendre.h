/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call site: maths.cpp:211. */
#ifndef TRV_ORACLE_SHIM_GSL_SF_LEGENDRE_H_
#define TRV_ORACLE_SHIM_GSL_SF_LEGENDRE_H_
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_legendre_sphPlm(const int l, const int m, const double x);
#ifdef __cplusplus
}
#endif
#endif
