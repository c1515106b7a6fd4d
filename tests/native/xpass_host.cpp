// Host-side emulation of k_xpass_fused (csrc/trvb_xpass.cuh): the stage functions are
// called for tid = 0 .. NT-1 in turn, one "block barrier" between stages, on a random
// [n0][n1][nh] array read from stdin-named files.  Used by tests/test_xpass_host.py.
//   xpass_host N n1 nh s0 s1 s2 add_a add_b S_re inv_vol in.bin ralias.bin out.bin lowk.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "trvb_xpass.cuh"

using namespace xpass;

template <int N, int CK, int NT>
void run_tile(long long c0, double2* T, const Pointwise& pw, const std::vector<double2>& tw) {
  std::vector<double2> tile((size_t)N * CK);
  constexpr int NS = Radix<N>::NS;
  auto col_of = [&](int tid) { return column_of(pw, c0 + tid % CK); };
  if (getenv("XP_ASYNC")) {   // the cp.async variant: tile loaded first, stage 0 on "shared memory"
    for (int tid = 0; tid < NT; tid++) stage_load_async<N, CK, NT>(tid, T, pw.ncols, c0, pw.ncols, tile.data());
    for (int tid = 0; tid < NT; tid++) stage_fwd<N, CK, NT, 0>(tid, tile.data(), tw.data());
  } else {
    for (int tid = 0; tid < NT; tid++) stage_first<N, CK, NT>(tid, T, pw.ncols, col_of(tid), tile.data(), tw.data());
  }
  if constexpr (NS >= 3) for (int tid = 0; tid < NT; tid++) stage_fwd<N, CK, NT, 1>(tid, tile.data(), tw.data());
  if constexpr (NS >= 4) for (int tid = 0; tid < NT; tid++) stage_fwd<N, CK, NT, 2>(tid, tile.data(), tw.data());
  for (int tid = 0; tid < NT; tid++) stage_junction<N, CK, NT>(tid, tile.data(), pw, col_of(tid));
  if constexpr (NS >= 4) for (int tid = 0; tid < NT; tid++) stage_inv<N, CK, NT, 2>(tid, tile.data(), tw.data());
  if constexpr (NS >= 3) for (int tid = 0; tid < NT; tid++) stage_inv<N, CK, NT, 1>(tid, tile.data(), tw.data());
  for (int tid = 0; tid < NT; tid++) stage_last<N, CK, NT>(tid, tile.data(), tw.data(), col_of(tid), T, pw.ncols);
}

template <int N, int CK>
void run_all(double2* T, const Pointwise& pw) {
  std::vector<double2> tw(N);
  for (int t = 0; t < N; t++) {
    const long double a = -2.0L * M_PIl * t / N;
    tw[t] = make_double2((double)cosl(a), (double)sinl(a));
  }
  for (long long c0 = 0; c0 < pw.ncols; c0 += CK) run_tile<N, CK, 256>(c0, T, pw, tw);
}

int main(int argc, char** argv) {
  if (argc != 15) { fprintf(stderr, "usage\n"); return 2; }
  const int N = atoi(argv[1]), n1 = atoi(argv[2]), nh = atoi(argv[3]);
  Pointwise pw;
  pw.n0 = N; pw.n1 = n1; pw.nh = nh; pw.ncols = (long long)n1 * nh;
  pw.s0 = atoi(argv[4]); pw.s1 = atoi(argv[5]); pw.s2 = atoi(argv[6]); pw.sh = pw.s2 / 2 + 1;
  pw.add_a = atof(argv[7]); pw.add_b = atof(argv[8]); pw.S_re = atof(argv[9]); pw.S_im = 0.;
  pw.inv_vol = atof(argv[10]);
  std::vector<double2> T((size_t)N * n1 * nh);
  std::vector<double> ral((size_t)N + n1 + nh);
  FILE* f = fopen(argv[11], "rb");
  if (!f || fread(T.data(), sizeof(double2), T.size(), f) != T.size()) return 3;
  fclose(f);
  f = fopen(argv[12], "rb");
  if (!f || fread(ral.data(), sizeof(double), ral.size(), f) != ral.size()) return 3;
  fclose(f);
  pw.ralias0 = ral.data(); pw.ralias1 = ral.data() + N; pw.ralias2 = ral.data() + N + n1;
  std::vector<double2> lowk((size_t)pw.s0 * pw.s1 * pw.sh, make_double2(0., 0.));
  pw.lowk = lowk.data();
  switch (N) {
    case 32: run_all<32, 8>(T.data(), pw); break;
    case 64: run_all<64, 8>(T.data(), pw); break;
    case 128: run_all<128, 8>(T.data(), pw); break;
    case 256: run_all<256, 8>(T.data(), pw); break;
    case 512: if (getenv("XP_CK4")) run_all<512, 4>(T.data(), pw); else run_all<512, 8>(T.data(), pw); break;
    case 1024: run_all<1024, 4>(T.data(), pw); break;
    case 2048: run_all<2048, 4>(T.data(), pw); break;
    default: return 4;
  }
  f = fopen(argv[13], "wb"); fwrite(T.data(), sizeof(double2), T.size(), f); fclose(f);
  f = fopen(argv[14], "wb"); fwrite(lowk.data(), sizeof(double2), lowk.size(), f); fclose(f);
  return 0;
}
