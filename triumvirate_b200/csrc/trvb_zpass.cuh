// trvb_zpass.cuh -- last pass of the pruned shell transform: complex-to-real along z with
// pruned input, hand-written (S/field.cpp:1792-1906 builds each shell field with a full 3-D
// inverse FFT; trvb_shell_slab_batch splits it into per-axis passes over the non-zero modes).
//
// After the x and y passes a shell field is B[r][kz][y], kz = 0 .. K2-1 (K2 - 1 = the
// shell's cut-off: a quarter of the z-extent N), r = (shell, x-plane); what is wanted is
// out[r][y][z], N reals per line.  cuFFT's c2r reads zero-padded lines of N/2 + 1 modes, so
// the padded copy (k_shell_zlines) was written and read back: 2 x N^2 (N/2+1) x 16 B per
// shell on top of the result itself.  Here a CTA takes 2 LP adjacent lines of one r:
//
//   load    B[r][kz][y0 .. y0 + 2 LP): two adjacent lines A, B are packed into ONE complex
//           line C = F_A + i F_B (C[k] = F_A[k] + i F_B[k], C[N-k] = conj F_A[k] + i conj F_B[k]),
//           stored at the digit-reversed slot of its frequency; the other slots are zeroed;
//   stages  decimation-in-time mixed-radix inverse FFT in shared memory, radices up to 12
//           built from 2/3/4/5-point kernels, in place, one block barrier per stage;
//   store   the last stage leaves natural order: Re -> line A, Im -> line B, lanes along z.
//
// HBM traffic per shell: N^2 K2 x 16 B in, N^3 x 8 B out -- nothing else.
//
// Like trvb_xpass.cuh, every function takes the thread index as an argument and compiles
// as plain C++ (tests/test_xpass_host.py checks all supported lengths against numpy).
#ifndef TRVB_ZPASS_CUH_
#define TRVB_ZPASS_CUH_

#include "trvb_xpass.cuh"

namespace xpass {

// Stage radices of the supported z-extents (first = the LAST stage executed, whose stride
// N / r(0) is the lane-contiguous one of the global stores).
#define ZP_RADIX3(N_, A_, B_, C_) \
  template <> struct Radix<N_> { static constexpr int NS = 3; \
    XP_HD static constexpr int r(int t) { return t == 0 ? A_ : (t == 1 ? B_ : C_); } };
#define ZP_RADIX2(N_, A_, B_) \
  template <> struct Radix<N_> { static constexpr int NS = 2; \
    XP_HD static constexpr int r(int t) { return t == 0 ? A_ : B_; } };
ZP_RADIX2(72, 8, 9)
ZP_RADIX2(96, 8, 12)
ZP_RADIX2(108, 12, 9)
ZP_RADIX2(144, 12, 12)
ZP_RADIX3(160, 4, 4, 10)
ZP_RADIX3(180, 6, 6, 5)
ZP_RADIX3(192, 4, 4, 12)
ZP_RADIX3(216, 6, 6, 6)
ZP_RADIX3(240, 4, 6, 10)
ZP_RADIX3(270, 6, 5, 9)
ZP_RADIX3(288, 4, 8, 9)
ZP_RADIX3(320, 8, 8, 5)
ZP_RADIX3(360, 6, 6, 10)
ZP_RADIX3(384, 8, 8, 6)
ZP_RADIX3(432, 8, 6, 9)
ZP_RADIX3(480, 8, 6, 10)
ZP_RADIX3(540, 6, 10, 9)
ZP_RADIX3(576, 8, 8, 9)
ZP_RADIX3(600, 6, 10, 10)
ZP_RADIX3(640, 8, 8, 10)
ZP_RADIX3(720, 8, 10, 9)
#undef ZP_RADIX2
#undef ZP_RADIX3

template <int SIGN> struct Dft<3, SIGN> {
  XP_HD static void run(double2* x) {
    const double h = 0.86602540378443864676;   // sqrt(3) / 2
    const double2 t = cadd(x[1], x[2]);
    const double2 m = make_double2(x[0].x - 0.5 * t.x, x[0].y - 0.5 * t.y);
    const double2 d = csub(x[1], x[2]);
    const double2 s = mul_i<SIGN>(make_double2(h * d.x, h * d.y));
    x[0] = cadd(x[0], t); x[1] = cadd(m, s); x[2] = csub(m, s);
  }
};
template <int SIGN> struct Dft<5, SIGN> {
  XP_HD static void run(double2* x) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    const double2 a1 = cadd(x[1], x[4]), a2 = cadd(x[2], x[3]);
    const double2 b1 = csub(x[1], x[4]), b2 = csub(x[2], x[3]);
    const double2 r1 = make_double2(x[0].x + c1 * a1.x + c2 * a2.x, x[0].y + c1 * a1.y + c2 * a2.y);
    const double2 r2 = make_double2(x[0].x + c2 * a1.x + c1 * a2.x, x[0].y + c2 * a1.y + c1 * a2.y);
    const double2 i1 = mul_i<SIGN>(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const double2 i2 = mul_i<SIGN>(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    x[0] = cadd(x[0], cadd(a1, a2));
    x[1] = cadd(r1, i1); x[4] = csub(r1, i1);
    x[2] = cadd(r2, i2); x[3] = csub(r2, i2);
  }
};

// R-point DFT in registers with the stage's twiddle table at hand: composite radices are
// one Cooley-Tukey step (n = n1 R2 + n2, k = k1 + R1 k2) whose inner twiddles
// w_R^(n2 k1) = w_N^(n2 k1 N / R) come from the table of the length-N transform (R | N).
// `tw[t] = exp(-2 pi i t / N)`; SIGN = +1 uses the conjugates.
template <int R, int SIGN, int N> struct DftT {
  XP_HD static void run(double2* x, const double2*) { Dft<R, SIGN>::run(x); }
};
template <int R1, int R2, int SIGN, int N>
XP_HD void dft_ct(double2* x, const double2* tw) {
  constexpr int R = R1 * R2;
  double2 y[R];
#pragma unroll
  for (int n2 = 0; n2 < R2; n2++) {
    double2 col[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; n1++) col[n1] = x[n1 * R2 + n2];
    Dft<R1, SIGN>::run(col);
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) {
      if (n2 * k1 == 0) y[k1 * R2 + n2] = col[k1];
      else {
        const double2 w = tw[n2 * k1 * (N / R)];
        y[k1 * R2 + n2] = SIGN > 0 ? cmul_conj(col[k1], w) : cmul(col[k1], w);
      }
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) {
    double2 row[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; n2++) row[n2] = y[k1 * R2 + n2];
    Dft<R2, SIGN>::run(row);
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) x[k1 + R1 * k2] = row[k2];
  }
}
template <int SIGN, int N> struct DftT<6, SIGN, N> {
  XP_HD static void run(double2* x, const double2* tw) { dft_ct<2, 3, SIGN, N>(x, tw); }
};
template <int SIGN, int N> struct DftT<9, SIGN, N> {
  XP_HD static void run(double2* x, const double2* tw) { dft_ct<3, 3, SIGN, N>(x, tw); }
};
template <int SIGN, int N> struct DftT<10, SIGN, N> {
  XP_HD static void run(double2* x, const double2* tw) { dft_ct<2, 5, SIGN, N>(x, tw); }
};
template <int SIGN, int N> struct DftT<12, SIGN, N> {
  XP_HD static void run(double2* x, const double2* tw) { dft_ct<4, 3, SIGN, N>(x, tw); }
};

// Tile slot of frequency k (inverse of freq_of_pos).
template <int N> XP_HD int pos_of_freq(int k) {
  int pos = 0, rem = k;
  for (int t = 0; t < Radix<N>::NS; t++) {
    const int R = Radix<N>::r(t), s = block_len<N>(t) / R;
    pos += (rem % R) * s;
    rem /= R;
  }
  return pos;
}

// Row pitch of the tile [LP][pitch] in complex elements.  A 16-byte access is served eight
// lanes at a time, one 16-byte bank group each.  With eight or more lines per tile those
// lanes are eight lines at one slot: an odd pitch spreads them over the eight groups.  With
// four lines they are four lines at two slots one apart (consecutive butterflies; nine
// apart in the unit-stride stage of 540 = 6 x 10 x 9, the same modulo 8): a pitch = 2 (mod 8)
// puts the lines on the even groups and the second slot on the odd ones (N + 1 left every
// access two-way conflicted: 46 % of the wavefronts at N = 540, ncu).
template <int N, int LP> XP_HD constexpr int zp_pitch() {
  if (LP >= 8) return N + 1;
  int p = N + 1;
  while (p % 8 != 2) p++;
  return p;
}

// Load: the K2 non-zero modes of 2 LP adjacent lines -> packed complex lines in the tile.
//   Bq = B + r K2 n1 (this (shell, plane)'s block [kz][y]); lines y0 .. y0 + 2 LP - 1.
template <int N, int LP, int NT>
XP_HD void zstage_load(int tid, const double2* __restrict__ Bq, int K2, int n1, int y0,
                       double2* tile) {
  constexpr int P = zp_pitch<N, LP>();
  // slots of the frequencies K2 .. N - K2 hold zeros
  const int nzero = N - 2 * K2 + 1;
  for (int e = tid; e < nzero * LP; e += NT) {
    const int j = e % LP, k = K2 + e / LP;
    tile[j * P + pos_of_freq<N>(k)] = make_double2(0., 0.);
  }
  // UN elements per thread with all their global loads issued before the first use
  constexpr int UN = 4;
  for (int e0 = tid; e0 < K2 * LP; e0 += NT * UN) {
    double2 fa[UN], fb[UN];
#pragma unroll
    for (int i = 0; i < UN; i++) {
      const int e = e0 + i * NT;
      const int j = e % LP, kz = e / LP, ya = y0 + 2 * j;
      const double2* src = Bq + (long long)kz * n1 + ya;
      fa[i] = (e < K2 * LP && ya < n1) ? src[0] : make_double2(0., 0.);
      fb[i] = (e < K2 * LP && ya + 1 < n1) ? src[1] : make_double2(0., 0.);
    }
#pragma unroll
    for (int i = 0; i < UN; i++) {
      const int e = e0 + i * NT;
      if (e >= K2 * LP) break;
      const int j = e % LP, kz = e / LP;
      if (kz == 0) {
        // the imaginary part of the k_z = 0 mode is dropped, as a c2r transform does
        tile[j * P + pos_of_freq<N>(0)] = make_double2(fa[i].x, fb[i].x);
      } else {
        // C[k] = F_A + i F_B,  C[N - k] = conj(F_A) + i conj(F_B)
        tile[j * P + pos_of_freq<N>(kz)] = make_double2(fa[i].x - fb[i].y, fa[i].y + fb[i].x);
        tile[j * P + pos_of_freq<N>(N - kz)] = make_double2(fa[i].x + fb[i].y, fb[i].x - fa[i].y);
      }
    }
  }
}

// Decimation-in-time stage t (NS - 1 down to 1), in place on the tile.
template <int N, int LP, int NT, int STAGE>
XP_HD void zstage(int tid, double2* tile, const double2* tw) {
  constexpr int R = Radix<N>::r(STAGE), L = block_len<N>(STAGE), S = L / R, NB = N / R;
  constexpr int P = zp_pitch<N, LP>();
  for (int u = tid; u < NB * LP; u += NT) {
    // lanes across the lines: the odd pitch puts the LP rows of one slot on distinct
    // 16-byte bank groups whatever the stride of the stage (lanes along the butterflies of
    // one line conflict two ways wherever a group of eight straddles a block boundary and
    // S is not a multiple of 8: a quarter of all wavefronts at N = 540, ncu)
    const int j = u % LP, b = u / LP;
    const int base = (b / S) * L, p = b % S;
    double2* line = tile + j * P + base + p;
    double2 x[R];
#pragma unroll
    for (int q = 0; q < R; q++) x[q] = line[q * S];
    if (S > 1) {
#pragma unroll
      for (int q = 1; q < R; q++) x[q] = cmul_conj(x[q], tw[q * p * (N / L)]);
    }
    DftT<R, +1, N>::run(x, tw);
#pragma unroll
    for (int m = 0; m < R; m++) line[m * S] = x[m];
  }
}

// Last stage (t = 0): natural order, Re -> line y0 + 2 j, Im -> line y0 + 2 j + 1 of
// out_r = out + r n1 N.
template <int N, int LP, int NT>
XP_HD void zstage_store(int tid, const double2* tile, const double2* tw, int n1, int y0,
                        double* __restrict__ out_r) {
  constexpr int R = Radix<N>::r(0), S = N / R;
  constexpr int P = zp_pitch<N, LP>();
  for (int u = tid; u < S * LP; u += NT) {
    const int j = u / S, p = u % S;
    const double2* line = tile + j * P + p;
    double2 x[R];
#pragma unroll
    for (int q = 0; q < R; q++) x[q] = line[q * S];
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul_conj(x[q], tw[q * p]);
    DftT<R, +1, N>::run(x, tw);
    const int ya = y0 + 2 * j;
    if (ya < n1) {
      double* oa = out_r + (long long)ya * N + p;
#pragma unroll
      for (int m = 0; m < R; m++) oa[m * S] = x[m].x;
      if (ya + 1 < n1) {
#pragma unroll
        for (int m = 0; m < R; m++) oa[N + m * S] = x[m].y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// The y pass in the same scheme: complex-to-complex inverse transform of length N along y
// with pruned input.  After the x pass a shell is A[q][c][b][x] (c = k_z, b = k_y + mc1 in
// [0, K1), x over the whole line); the z pass wants B[q][xi][c][y] for the planes
// xi in [0, nx) (x = x0 + xi).  A CTA takes XT adjacent x of one (q, c): loads are
// coalesced along x, each line's K1 modes go to the digit-reversed slots of their
// frequencies, the rest of the line is zeroed, and the last stage stores natural order with
// lanes along y.  Replaces k_shell_ylines (zero-padded transpose) + a batched cuFFT Z2Z.
// ---------------------------------------------------------------------------------------

// Aqc = A + (q K2 + c) K1 n0: rows b = 0 .. K1-1 of n0 elements; xs = first x of the tile.
template <int N, int XT, int NT>
XP_HD void ystage_load(int tid, const double2* __restrict__ Aqc, int K1, int mc1, int n0, int xs,
                       int xe /* one past the last valid x */, double2* tile) {
  constexpr int P = zp_pitch<N, XT>();
  // frequencies mc1 + 1 .. N - mc1 - 1 are zero
  const int nzero = N - K1;
  for (int e = tid; e < nzero * XT; e += NT) {
    const int j = e % XT, k = mc1 + 1 + e / XT;
    tile[j * P + pos_of_freq<N>(k)] = make_double2(0., 0.);
  }
  constexpr int UN = 4;
  for (int e0 = tid; e0 < K1 * XT; e0 += NT * UN) {
    double2 v[UN];
#pragma unroll
    for (int i = 0; i < UN; i++) {
      const int e = e0 + i * NT;
      const int j = e % XT, b = e / XT;
      v[i] = (e < K1 * XT && xs + j < xe) ? Aqc[(long long)b * n0 + xs + j] : make_double2(0., 0.);
    }
#pragma unroll
    for (int i = 0; i < UN; i++) {
      const int e = e0 + i * NT;
      if (e >= K1 * XT) break;
      const int j = e % XT, b = e / XT;
      const int mj = b - mc1;
      tile[j * P + pos_of_freq<N>(mj >= 0 ? mj : mj + N)] = v[i];
    }
  }
}

// Last stage: natural order in y.  Bq = B + q nx K2 n1 + c n1; line xi lives at
// Bq + xi K2 n1 (row_stride = K2 n1).
template <int N, int XT, int NT>
XP_HD void ystage_store(int tid, const double2* tile, const double2* tw, int xi0, int nx,
                        long long row_stride, double2* __restrict__ Bq) {
  constexpr int R = Radix<N>::r(0), S = N / R;
  constexpr int P = zp_pitch<N, XT>();
  for (int u = tid; u < S * XT; u += NT) {
    const int j = u / S, p = u % S;
    const double2* line = tile + j * P + p;
    double2 x[R];
#pragma unroll
    for (int q = 0; q < R; q++) x[q] = line[q * S];
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul_conj(x[q], tw[q * p]);
    DftT<R, +1, N>::run(x, tw);
    if (xi0 + j < nx) {
      double2* o = Bq + (long long)(xi0 + j) * row_stride + p;
#pragma unroll
      for (int m = 0; m < R; m++) o[m * S] = x[m];
    }
  }
}

}  // namespace xpass

#endif  // TRVB_ZPASS_CUH_
