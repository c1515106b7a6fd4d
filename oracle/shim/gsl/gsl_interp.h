/* TEST INFRASTRUCTURE ONLY (oracle shim).  Call sites: maths.cpp:335-338,350-364,373. */
#ifndef TRV_ORACLE_SHIM_GSL_INTERP_H_
#define TRV_ORACLE_SHIM_GSL_INTERP_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t cache; size_t miss_count; size_t hit_count; } gsl_interp_accel;
typedef struct { const char* name; unsigned int min_size; } gsl_interp_type;
extern const gsl_interp_type* gsl_interp_cspline;
gsl_interp_accel* gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel* a);
#ifdef __cplusplus
}
#endif
#endif
