/* TEST INFRASTRUCTURE ONLY (oracle shim). */
#ifndef TRV_ORACLE_SHIM_GSL_SF_RESULT_H_
#define TRV_ORACLE_SHIM_GSL_SF_RESULT_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { double val; double err; } gsl_sf_result;
#ifdef __cplusplus
}
#endif
#endif
