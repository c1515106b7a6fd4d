// trvb_xpass.cuh -- the x pass of the box estimators' two full-grid transforms, fused.
//
// The periodic-box bispectrum needs two things from the full-grid spectrum delta n(k)
// (S/threept.cpp:1554-1558, S/field.cpp:3273-3345):
//   * its low-|k| modes (what the sub-grid of the pair phase represents), and
//   * xi(r) = IFFT[(delta n(k) conj N(k) / C1(k) - S) / V] for the shot-noise term.
// delta n(k) itself is never read again.  With the 3-D transforms split as 2-D (planes,
// cuFFT) + 1-D (x), the x pass of the forward transform, the spectrum arithmetic and the
// x pass of the inverse transform touch the same column of 16-byte elements:
//
//   T[i][c] (c = j nh + k, after the 2-D D2Z of the planes)
//     -> forward FFT along i            (decimation in frequency, digit-reversed result)
//     -> low-|k| modes stored to the sub-grid's HALF mesh; spectrum formed in registers
//     -> inverse FFT along i            (decimation in time, natural-order result)
//     -> T[i][c] in place               (then the 2-D Z2D of the planes gives xi(r))
//
// one read and one write of the half spectrum instead of (forward x pass: read + write),
// (spectrum kernel: read + write) and (inverse x pass: read + write).
//
// A CTA owns CK adjacent columns (128 contiguous bytes per x-plane for CK = 8) and all n0
// elements of each: a 64 KB tile in shared memory.  The transform is a mixed-radix
// (8, 8, 8 | 8, 8, 4, 4 | ...) in-place FFT; a stage reads and writes the same R slots per
// butterfly, so the only synchronisation is one block barrier between stages.  The first
// forward stage reads global memory and the last inverse stage writes it; the last forward
// stage, the pointwise arithmetic and the first inverse stage run on registers.
//
// Every function here is __host__ __device__ and takes the thread index as an argument:
// tests/test_xpass_host.py compiles this header with g++ and checks the stage sequence
// against a direct DFT on the CPU (no GPU needed for the index arithmetic).
#ifndef TRVB_XPASS_CUH_
#define TRVB_XPASS_CUH_

#ifdef __CUDACC__
#define XP_HD __host__ __device__ __forceinline__
#else
#define XP_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace xpass {

// Radices of the stages, first forward stage first.  The last one is the radix of the
// register-resident junction (forward last stage / inverse first stage).
template <int N> struct Radix;
template <> struct Radix<32>   { static constexpr int NS = 2; XP_HD static constexpr int r(int t) { return t == 0 ? 8 : 4; } };
template <> struct Radix<64>   { static constexpr int NS = 2; XP_HD static constexpr int r(int t) { return 8; } };
template <> struct Radix<128>  { static constexpr int NS = 3; XP_HD static constexpr int r(int t) { return t == 0 ? 8 : 4; } };
template <> struct Radix<256>  { static constexpr int NS = 3; XP_HD static constexpr int r(int t) { return t < 2 ? 8 : 4; } };
template <> struct Radix<512>  { static constexpr int NS = 3; XP_HD static constexpr int r(int t) { return 8; } };
template <> struct Radix<1024> { static constexpr int NS = 4; XP_HD static constexpr int r(int t) { return t < 2 ? 8 : 4; } };
template <> struct Radix<2048> { static constexpr int NS = 4; XP_HD static constexpr int r(int t) { return t < 3 ? 8 : 4; } };

// Block length of stage t (N / product of the earlier radices) and the butterfly stride.
template <int N> XP_HD constexpr int block_len(int t) {
  int L = N;
  for (int u = 0; u < t; u++) L /= Radix<N>::r(u);
  return L;
}

XP_HD double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
XP_HD double2 cmul_conj(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
XP_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
XP_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a * (SIGN i)
template <int SIGN> XP_HD double2 mul_i(double2 a) {
  return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// In-register DFT of R points: x_q <- sum_m x_m exp(SIGN 2 pi i m q / R).
template <int R, int SIGN> struct Dft;
template <int SIGN> struct Dft<2, SIGN> {
  XP_HD static void run(double2* x) {
    const double2 a = x[0], b = x[1];
    x[0] = cadd(a, b); x[1] = csub(a, b);
  }
};
template <int SIGN> struct Dft<4, SIGN> {
  XP_HD static void run(double2* x) {
    const double2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    const double2 t2 = cadd(x[1], x[3]), t3 = mul_i<SIGN>(csub(x[1], x[3]));
    x[0] = cadd(t0, t2); x[2] = csub(t0, t2);
    x[1] = cadd(t1, t3); x[3] = csub(t1, t3);
  }
};
template <int SIGN> struct Dft<8, SIGN> {
  XP_HD static void run(double2* x) {
    double2 e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
    Dft<4, SIGN>::run(e);
    Dft<4, SIGN>::run(o);
    const double h = 0.70710678118654752440;
    // w^1 = (1 + SIGN i) / sqrt 2, w^2 = SIGN i, w^3 = (-1 + SIGN i) / sqrt 2
    const double2 o1 = SIGN > 0 ? make_double2(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                                : make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    const double2 o2 = mul_i<SIGN>(o[2]);
    const double2 o3 = SIGN > 0 ? make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                                : make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
    x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
    x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
    x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
    x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
  }
};

// Physical row of logical position `pos` in the tile.  With CK = 4 a 16-byte access of
// eight lanes covers two positions; they must differ in parity to hit distinct bank groups,
// which positions R_last apart (the junction) do not: bit log2(R_last) is folded into bit 0.
template <int CK, int RLAST> XP_HD int phys_row(int pos) {
  if (CK >= 8) return pos;
  return pos ^ ((pos / RLAST) & 1);
}

// Frequency index of tile position `pos` after all forward stages (digit reversal of the
// mixed radices): pos = sum_t q_t s_t  ->  k = q_0 + R_0 (q_1 + R_1 (q_2 + ...)).
template <int N> XP_HD int freq_of_pos(int pos) {
  int k = 0, mult = 1, rem = pos;
  for (int t = 0; t < Radix<N>::NS; t++) {
    const int s = block_len<N>(t) / Radix<N>::r(t);
    const int q = rem / s;
    rem -= q * s;
    k += q * mult;
    mult *= Radix<N>::r(t);
  }
  return k;
}

// What the junction needs to know about the mesh (all plain values and device pointers).
struct Pointwise {
  int n0, n1, nh;               // extents of T: [n0][n1][nh]
  long long ncols;              // n1 * nh
  const double* ralias0;        // 1 / (per-axis factor of C1), axis 0 .. 2
  const double* ralias1;
  const double* ralias2;
  double add_a, add_b;          // fa = U + add_a delta_k0, fb = U + add_b delta_k0
  double S_re, S_im, inv_vol;
  int s0, s1, s2, sh;           // extents of the low-|k| HALF mesh (0: none requested)
  double2* lowk;                // [s0][s1][sh]
};

// Per-thread constants of a column (the thread's ck never changes: NT % CK == 0).
struct Column {
  long long c;       // flattened column index j nh + k; < 0: beyond the mesh
  int j, k;
  double r12;        // ralias1[j] * ralias2[k]
  int js;            // row of the low-|k| mesh, < 0: not represented
};

XP_HD Column column_of(const Pointwise& pw, long long c) {
  Column col;
  col.c = c < pw.ncols ? c : -1;
  col.j = 0; col.k = 0; col.r12 = 0.; col.js = -1;
  if (col.c < 0) return col;
  col.j = (int)(c / pw.nh);
  col.k = (int)(c - (long long)col.j * pw.nh);
  col.r12 = pw.ralias1[col.j] * pw.ralias2[col.k];
  if (pw.s0 > 0) {
    const int mj = col.j < pw.n1 / 2 ? col.j : col.j - pw.n1;
    const int amj = mj < 0 ? -mj : mj;
    if (2 * amj < pw.s1 && 2 * col.k < pw.s2) col.js = mj >= 0 ? mj : mj + pw.s1;
  }
  return col;
}

// --- stages.  `tile` is [N][CK] double2 (physical rows), `tw[t] = exp(-2 pi i t / N)`. ---

// First forward stage: rows of global memory -> tile.
template <int N, int CK, int NT>
XP_HD void stage_first(int tid, const double2* __restrict__ T, long long plane_stride,
                       const Column& col, double2* tile, const double2* tw) {
  constexpr int R = Radix<N>::r(0), S = N / R, RL = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  for (int u = tid; u < S * CK; u += NT) {
    const int p = u / CK;
    double2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) {
      x[m] = col.c >= 0 ? T[(long long)(p + m * S) * plane_stride + col.c] : make_double2(0., 0.);
    }
    Dft<R, -1>::run(x);
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul(x[q], tw[q * p]);
#pragma unroll
    for (int q = 0; q < R; q++) tile[phys_row<CK, RL>(q * S + p) * CK + ck] = x[q];
  }
}

// Asynchronous variant of the first stage's read: the whole tile goes from global to shared
// memory with 16-byte cp.async copies (no registers, every byte of the tile in flight at
// once), then stage_fwd<.., 0> runs on shared memory like the middle stages.  Columns beyond
// the mesh are zero-filled.  Followed by stage_load_wait + a block barrier.
template <int N, int CK, int NT>
XP_HD void stage_load_async(int tid, const double2* __restrict__ T, long long plane_stride,
                            long long c0, long long ncols, double2* tile) {
  constexpr int RL = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  const bool inside = c0 + ck < ncols;
  for (int e = tid; e < N * CK; e += NT) {
    const int row = e / CK;
    double2* dst = tile + phys_row<CK, RL>(row) * CK + ck;
    if (!inside) { *dst = make_double2(0., 0.); continue; }
    const double2* src = T + (long long)row * plane_stride + c0 + ck;
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src) : "memory");
#else
    *dst = *src;
#endif
  }
}
XP_HD void stage_load_wait() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Forward stage t (0 <= t < NS - 1), in place on the tile.
template <int N, int CK, int NT, int STAGE>
XP_HD void stage_fwd(int tid, double2* tile, const double2* tw) {
  constexpr int R = Radix<N>::r(STAGE), L = block_len<N>(STAGE), S = L / R;
  constexpr int RL = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  for (int u = tid; u < (N / R) * CK; u += NT) {
    const int b = u / CK, base = (b / S) * L, p = b % S;
    double2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) x[m] = tile[phys_row<CK, RL>(base + p + m * S) * CK + ck];
    Dft<R, -1>::run(x);
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul(x[q], tw[q * p * (N / L)]);
#pragma unroll
    for (int q = 0; q < R; q++) tile[phys_row<CK, RL>(base + q * S + p) * CK + ck] = x[q];
  }
}

// Junction: last forward stage (no twiddles), low-|k| store, spectrum, first inverse stage.
template <int N, int CK, int NT>
XP_HD void stage_junction(int tid, double2* tile, const Pointwise& pw, const Column& col) {
  constexpr int R = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  for (int u = tid; u < (N / R) * CK; u += NT) {
    const int b = u / CK;
    double2 x[R];
#pragma unroll
    for (int m = 0; m < R; m++) x[m] = tile[phys_row<CK, R>(b * R + m) * CK + ck];
    Dft<R, -1>::run(x);
    const int i0 = freq_of_pos<N>(b * R);
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int i = i0 + q * (N / R);          // x[q] = U(i, j, k)
      double2 a = x[q], bb = x[q];
      if ((i | col.j | col.k) == 0) { a.x += pw.add_a; bb.x += pw.add_b; }
      if (col.js >= 0) {
        const int mi = i < N / 2 ? i : i - N;
        const int ami = mi < 0 ? -mi : mi;
        if (2 * ami < pw.s0) {
          const int is = mi >= 0 ? mi : mi + pw.s0;
          pw.lowk[((long long)is * pw.s1 + col.js) * pw.sh + col.k] = a;
        }
      }
      const double rc1 = pw.ralias0[i] * col.r12;
      const double re = (a.x * bb.x + a.y * bb.y) * rc1 - pw.S_re;
      const double im = (a.y * bb.x - a.x * bb.y) * rc1 - pw.S_im;
      x[q] = make_double2(re * pw.inv_vol, im * pw.inv_vol);
    }
    Dft<R, +1>::run(x);
#pragma unroll
    for (int m = 0; m < R; m++) tile[phys_row<CK, R>(b * R + m) * CK + ck] = x[m];
  }
}

// Inverse stage t (0 < t < NS - 1), in place on the tile (run in DEcreasing t).
template <int N, int CK, int NT, int STAGE>
XP_HD void stage_inv(int tid, double2* tile, const double2* tw) {
  constexpr int R = Radix<N>::r(STAGE), L = block_len<N>(STAGE), S = L / R;
  constexpr int RL = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  for (int u = tid; u < (N / R) * CK; u += NT) {
    const int b = u / CK, base = (b / S) * L, p = b % S;
    double2 x[R];
#pragma unroll
    for (int q = 0; q < R; q++) x[q] = tile[phys_row<CK, RL>(base + q * S + p) * CK + ck];
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul_conj(x[q], tw[q * p * (N / L)]);
    Dft<R, +1>::run(x);
#pragma unroll
    for (int m = 0; m < R; m++) tile[phys_row<CK, RL>(base + p + m * S) * CK + ck] = x[m];
  }
}

// Last inverse stage: tile -> rows of global memory.
template <int N, int CK, int NT>
XP_HD void stage_last(int tid, const double2* tile, const double2* tw, const Column& col,
                      double2* __restrict__ T, long long plane_stride) {
  constexpr int R = Radix<N>::r(0), S = N / R, RL = Radix<N>::r(Radix<N>::NS - 1);
  const int ck = tid % CK;
  for (int u = tid; u < S * CK; u += NT) {
    const int p = u / CK;
    double2 x[R];
#pragma unroll
    for (int q = 0; q < R; q++) x[q] = tile[phys_row<CK, RL>(q * S + p) * CK + ck];
#pragma unroll
    for (int q = 1; q < R; q++) x[q] = cmul_conj(x[q], tw[q * p]);
    Dft<R, +1>::run(x);
    if (col.c >= 0) {
#pragma unroll
      for (int m = 0; m < R; m++) T[(long long)(p + m * S) * plane_stride + col.c] = x[m];
    }
  }
}

}  // namespace xpass

#endif  // TRVB_XPASS_CUH_
