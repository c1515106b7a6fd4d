// Host-side emulation of k_shell_zpass (csrc/trvb_zpass.cuh): stage functions called for
// tid = 0 .. NT-1 in turn with one "block barrier" between stages.
//   zpass_host N n1 K2 in.bin out.bin     (in: B[K2][n1] complex; out: [n1][N] doubles)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "trvb_zpass.cuh"

using namespace xpass;

template <int N>
int run(int n1, int K2, const std::vector<double2>& B, std::vector<double>& out) {
  constexpr int LP = 4, NT = 128, NS = Radix<N>::NS;
  std::vector<double2> tw(N);
  for (int t = 0; t < N; t++) {
    const long double a = -2.0L * M_PIl * t / N;
    tw[t] = make_double2((double)cosl(a), (double)sinl(a));
  }
  std::vector<double2> tile((size_t)LP * zp_pitch<N, LP>(), make_double2(1.e300, -1.e300));
  for (int y0 = 0; y0 < n1; y0 += 2 * LP) {
    for (int tid = 0; tid < NT; tid++) zstage_load<N, LP, NT>(tid, B.data(), K2, n1, y0, tile.data());
    if constexpr (NS >= 4) for (int tid = 0; tid < NT; tid++) zstage<N, LP, NT, 3>(tid, tile.data(), tw.data());
    if constexpr (NS >= 3) for (int tid = 0; tid < NT; tid++) zstage<N, LP, NT, 2>(tid, tile.data(), tw.data());
    for (int tid = 0; tid < NT; tid++) zstage<N, LP, NT, 1>(tid, tile.data(), tw.data());
    for (int tid = 0; tid < NT; tid++) zstage_store<N, LP, NT>(tid, tile.data(), tw.data(), n1, y0, out.data());
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 6) return 2;
  const int N = atoi(argv[1]), n1 = atoi(argv[2]), K2 = atoi(argv[3]);
  std::vector<double2> B((size_t)K2 * n1);
  std::vector<double> out((size_t)n1 * N, -7.);
  FILE* f = fopen(argv[4], "rb");
  if (!f || fread(B.data(), sizeof(double2), B.size(), f) != B.size()) return 3;
  fclose(f);
  int st = 4;
  switch (N) {
#define CASE(n) case n: st = run<n>(n1, K2, B, out); break;
    CASE(64) CASE(72) CASE(96) CASE(108) CASE(128) CASE(144) CASE(160) CASE(180) CASE(192)
    CASE(216) CASE(240) CASE(256) CASE(270) CASE(288) CASE(320) CASE(360) CASE(384) CASE(432)
    CASE(480) CASE(512) CASE(540) CASE(576) CASE(600) CASE(640) CASE(720)
#undef CASE
    default: break;
  }
  if (st) return st;
  f = fopen(argv[5], "wb"); fwrite(out.data(), sizeof(double), out.size(), f); fclose(f);
  return 0;
}
