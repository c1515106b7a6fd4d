// Development harness: cuFFT timings for candidate sub-grid sizes (batched vs looped, Z2Z vs Z2D).
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
int main() {
  int sizes[] = {128, 135, 140, 144, 150, 160, 256, 512};
  const int batch = 20;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int n : sizes) {
    int b = n >= 256 ? 1 : batch;
    size_t elems = (size_t)n * n * n;
    cufftDoubleComplex* d; cudaMalloc(&d, elems * 16 * b); cudaMemset(d, 0, elems * 16 * b);
    double* r; cudaMalloc(&r, elems * 8 * b);
    cufftHandle p1, pb, pr; int dims[3] = {n, n, n};
    cufftPlan3d(&p1, n, n, n, CUFFT_Z2Z);
    cufftPlanMany(&pb, 3, dims, nullptr, 1, (int)elems, nullptr, 1, (int)elems, CUFFT_Z2Z, b);
    cufftPlanMany(&pr, 3, dims, nullptr, 1, n * n * (n / 2 + 1), nullptr, 1, (int)elems, CUFFT_Z2D, b);
    float ms_loop, ms_batch, ms_real;
    for (int w = 0; w < 2; w++) {
      cudaEventRecord(e0);
      for (int rep = 0; rep < 5; rep++) for (int i = 0; i < b; i++) cufftExecZ2Z(p1, d + i * elems, d + i * elems, CUFFT_INVERSE);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_loop, e0, e1);
      cudaEventRecord(e0);
      for (int rep = 0; rep < 5; rep++) cufftExecZ2Z(pb, d, d, CUFFT_INVERSE);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_batch, e0, e1);
      cudaEventRecord(e0);
      for (int rep = 0; rep < 5; rep++) cufftExecZ2D(pr, d, r);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_real, e0, e1);
    }
    printf("n=%4d batch=%2d  Z2Z loop %8.3f ms  Z2Z batched %8.3f ms  Z2D batched %8.3f ms  (per transform: %.1f / %.1f / %.1f us)\n",
           n, b, ms_loop / 5, ms_batch / 5, ms_real / 5, 1e3 * ms_loop / 5 / b, 1e3 * ms_batch / 5 / b, 1e3 * ms_real / 5 / b);
    cufftDestroy(p1); cufftDestroy(pb); cufftDestroy(pr); cudaFree(d); cudaFree(r);
  }
  return 0;
}
