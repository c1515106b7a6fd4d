/* TEST INFRASTRUCTURE ONLY -- part of oracle/, never linked into the product.
 *
 * Declaration-compatible stand-in for the subset of <fftw3.h> that the
 * reference sources under /root/reference/src/triumvirate/src use
 * (call sites: field.cpp:56,246,299,339,1552-1554,1714-1716,2182;
 * fftlog.cpp 1-D plans).  FFTW3 itself is not installed in this image and
 * cannot be fetched; the implementation lives in trvshim_fft.cpp (own
 * mixed-radix complex FFT, unnormalised, in-place, row-major, sign -1
 * forward / +1 backward: the FFTW convention).
 */
#ifndef TRV_ORACLE_SHIM_FFTW3_H_
#define TRV_ORACLE_SHIM_FFTW3_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct trvshim_fftw_plan_s* fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

extern const char fftw_version[];

void* fftw_malloc(size_t n);
fftw_complex* fftw_alloc_complex(size_t n);
void fftw_free(void* p);

fftw_plan fftw_plan_dft_3d(
  int n0, int n1, int n2, fftw_complex* in, fftw_complex* out,
  int sign, unsigned flags
);
fftw_plan fftw_plan_dft_1d(
  int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags
);
void fftw_execute(const fftw_plan p);
void fftw_execute_dft(const fftw_plan p, fftw_complex* in, fftw_complex* out);
void fftw_destroy_plan(fftw_plan p);

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);
void fftw_cleanup(void);

int fftw_import_wisdom_from_filename(const char* filename);
int fftw_export_wisdom_to_filename(const char* filename);

#ifdef __cplusplus
}
#endif

#endif  /* TRV_ORACLE_SHIM_FFTW3_H_ */
