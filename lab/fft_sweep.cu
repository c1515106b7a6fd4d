// Development harness: cuFFT fp64 3-D batched Z2D / Z2Z cost per cell for every 7-smooth length in a range.
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
int main(int argc, char** argv) {
  int lo = argc > 1 ? atoi(argv[1]) : 64, hi = argc > 2 ? atoi(argv[2]) : 800;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t cap_cells = 420u * 1000u * 1000u;
  cufftDoubleComplex* d; cudaMalloc(&d, cap_cells * 16); cudaMemset(d, 0, cap_cells * 16);
  double* r; cudaMalloc(&r, cap_cells * 8);
  printf("# n batch z2d_ns_per_kcell z2z_ns_per_kcell\n");
  for (int n = lo; n <= hi; n++) {
    int m = n; for (int p : {2, 3, 5, 7}) while (m % p == 0) m /= p;
    if (m != 1) continue;
    size_t elems = (size_t)n * n * n;
    int b = (int)std::max<size_t>(1, std::min<size_t>(20, cap_cells / elems));
    int dims[3] = {n, n, n};
    cufftHandle pb, pr;
    if (cufftPlanMany(&pb, 3, dims, nullptr, 1, (int)elems, nullptr, 1, (int)elems, CUFFT_Z2Z, b) != CUFFT_SUCCESS) continue;
    if (cufftPlanMany(&pr, 3, dims, nullptr, 1, n * n * (n / 2 + 1), nullptr, 1, (int)elems, CUFFT_Z2D, b) != CUFFT_SUCCESS) { cufftDestroy(pb); continue; }
    float ms_c = 0, ms_r = 0;
    int reps = elems * b > 100000000 ? 2 : 5;
    for (int w = 0; w < 2; w++) {
      cudaEventRecord(e0); for (int q = 0; q < reps; q++) cufftExecZ2Z(pb, d, d, CUFFT_INVERSE); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c, e0, e1);
      cudaEventRecord(e0); for (int q = 0; q < reps; q++) cufftExecZ2D(pr, d, r); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_r, e0, e1);
    }
    double kc = (double)elems * b / 1000.;
    printf("%d %d %.2f %.2f\n", n, b, 1e6 * ms_r / reps / kc, 1e6 * ms_c / reps / kc); fflush(stdout);
    cufftDestroy(pb); cufftDestroy(pr);
  }
  return 0;
}
