// Host-side emulation of k_shell_ypass (csrc/trvb_zpass.cuh).
//   ypass_host N n0 mc1 x0 nx in.bin out.bin   (in: A[K1][n0] complex, K1 = 2 mc1 + 1;
//                                               out: B[nx][N] complex for x = x0 .. x0+nx-1)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "trvb_zpass.cuh"

using namespace xpass;

template <int N>
int run(int n0, int mc1, int x0, int nx, const std::vector<double2>& A, std::vector<double2>& out) {
  constexpr int XT = 4, NT = 128, NS = Radix<N>::NS;
  const int K1 = 2 * mc1 + 1;
  std::vector<double2> tw(N);
  for (int t = 0; t < N; t++) {
    const long double a = -2.0L * M_PIl * t / N;
    tw[t] = make_double2((double)cosl(a), (double)sinl(a));
  }
  std::vector<double2> tile((size_t)XT * zp_pitch<N, XT>(), make_double2(1.e300, -1.e300));
  for (int xi0 = 0; xi0 < nx; xi0 += XT) {
    for (int tid = 0; tid < NT; tid++) ystage_load<N, XT, NT>(tid, A.data(), K1, mc1, n0, x0 + xi0, x0 + nx, tile.data());
    if constexpr (NS >= 4) for (int tid = 0; tid < NT; tid++) zstage<N, XT, NT, 3>(tid, tile.data(), tw.data());
    if constexpr (NS >= 3) for (int tid = 0; tid < NT; tid++) zstage<N, XT, NT, 2>(tid, tile.data(), tw.data());
    for (int tid = 0; tid < NT; tid++) zstage<N, XT, NT, 1>(tid, tile.data(), tw.data());
    for (int tid = 0; tid < NT; tid++) ystage_store<N, XT, NT>(tid, tile.data(), tw.data(), xi0, nx, N, out.data());
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 8) return 2;
  const int N = atoi(argv[1]), n0 = atoi(argv[2]), mc1 = atoi(argv[3]), x0 = atoi(argv[4]), nx = atoi(argv[5]);
  const int K1 = 2 * mc1 + 1;
  std::vector<double2> A((size_t)K1 * n0);
  std::vector<double2> out((size_t)nx * N, make_double2(-7., -7.));
  FILE* f = fopen(argv[6], "rb");
  if (!f || fread(A.data(), sizeof(double2), A.size(), f) != A.size()) return 3;
  fclose(f);
  int st = 4;
  switch (N) {
#define CASE(n) case n: st = run<n>(n0, mc1, x0, nx, A, out); break;
    CASE(64) CASE(72) CASE(96) CASE(108) CASE(128) CASE(144) CASE(160) CASE(180) CASE(192)
    CASE(216) CASE(240) CASE(256) CASE(270) CASE(288) CASE(320) CASE(360) CASE(384) CASE(432)
    CASE(480) CASE(512) CASE(540) CASE(576) CASE(600) CASE(640) CASE(720)
#undef CASE
    default: break;
  }
  if (st) return st;
  f = fopen(argv[7], "wb"); fwrite(out.data(), sizeof(double2), out.size(), f); fclose(f);
  return 0;
}
