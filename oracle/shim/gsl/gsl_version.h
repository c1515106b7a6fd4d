/* TEST INFRASTRUCTURE ONLY (oracle shim; GSL is not installed in this image). */
#ifndef TRV_ORACLE_SHIM_GSL_VERSION_H_
#define TRV_ORACLE_SHIM_GSL_VERSION_H_
#define GSL_VERSION "2.7-trvshim"
#define GSL_MAJOR_VERSION 2
#define GSL_MINOR_VERSION 7
#ifdef __cplusplus
extern "C" {
#endif
extern const char* gsl_version;
#ifdef __cplusplus
}
#endif
#endif
