/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call site: maths.cpp:168. */
#ifndef TRV_ORACLE_SHIM_GSL_SF_COUPLING_H_
#define TRV_ORACLE_SHIM_GSL_SF_COUPLING_H_
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_coupling_3j(
  int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc
);
#ifdef __cplusplus
}
#endif
#endif
