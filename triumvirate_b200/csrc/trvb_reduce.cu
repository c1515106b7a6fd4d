// trvb_reduce.cu -- grid reductions and binned statistics of libtrvb.so.
//
//   trvb_gram_reduce         sum_x A_a B_b G for ALL (a, b) pairs in one pass
//                            (replaces the per-pair loops S/threept.cpp:1708-1717)
//   trvb_shot_bispec_reduce  the same tiled reduction with tiles COMPUTED
//                            (j_l splines, y_lm) instead of loaded
//                            (replaces S/field.cpp:3362-3393, once per call
//                            instead of once per bin pair -- SURVEY.md F4)
//   trvb_shell_stats, trvb_twopt_fourier, trvb_shot_xi, trvb_shot_3pcf_bin,
//   trvb_mesh_sum_pow3
//
// All sums are two-stage (block tree -> fixed-order pass over blocks), so a
// given launch configuration reproduces its result bit for bit.
#include "trvb_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace {

struct Tables {
  const double* sinc[3];
  const double* alias[3];
  const double* ralias[3];   // 1 / alias
};

Tables tables_of(const trvb_ctx* ctx) {
  const trvb_ctx* root = ctx->parent ? ctx->parent : ctx;
  Tables t;
  for (int a = 0; a < 3; a++) {
    t.sinc[a] = root->d_sinc[a]; t.alias[a] = root->d_alias[a]; t.ralias[a] = root->d_ralias[a];
  }
  return t;
}


__global__ void k_sum_cols(const double* __restrict__ partial, int nblocks, int width,
                           double* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= width) return;
  double s = 0.;
  for (int b = 0; b < nblocks; b++) s += partial[(long long)b * width + t];
  out[t] = s;
}

// partial is [bin][nblocks][width]; block `bin` sums its slab in fixed order.
__global__ void k_sum_binned(const double* __restrict__ partial, int nblocks, int width,
                             double* __restrict__ out) {
  const int t = threadIdx.x;
  if (t >= width) return;
  const double* slab = partial + (size_t)blockIdx.x * nblocks * width;
  double s = 0.;
  for (int b = 0; b < nblocks; b++) s += slab[(size_t)b * width + t];
  out[(size_t)blockIdx.x * width + t] = s;
}

// ---------------------------------------------------------------------
// Tiled all-pairs ("Gram") reduction:  out[a][b] = sum_x A_a(x) * H_b(x),
// H_b = B_b * G.  It is a skinny matrix product (K = cells, M = N = bins): a
// tile of GRAM_T cells of every field is staged in shared memory and each
// warp accumulates 4 x 4 blocks of (a, b) pairs in registers, so one shared
// load feeds four multiply-adds.  Values are complex (double2) or, when all
// operands are real meshes, plain doubles (VT).
// ---------------------------------------------------------------------

constexpr int GRAM_T = 64;        // cells per tile (two per lane)
constexpr int GRAM_WARPS = 8;
constexpr int GRAM_THREADS = GRAM_WARPS * 32;
constexpr int GRAM_B = 4;         // pair block edge

__device__ __forceinline__ void vt_zero(double& v) { v = 0.; }
__device__ __forceinline__ void vt_zero(double2& v) { v.x = 0.; v.y = 0.; }
__device__ __forceinline__ void vt_fma(double& acc, double a, double b) { acc += a * b; }
__device__ __forceinline__ void vt_fma(double2& acc, double2 a, double2 b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}
__device__ __forceinline__ double vt_shfl_down(double v, int o) {
  return __shfl_down_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ double2 vt_shfl_down(double2 v, int o) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, o), __shfl_down_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ void vt_add(double& a, double b) { a += b; }
__device__ __forceinline__ void vt_add(double2& a, double2 b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vt_store(double* p, double v) { p[0] = v; p[1] = 0.; }
__device__ __forceinline__ void vt_store(double* p, double2 v) { p[0] = v.x; p[1] = v.y; }

// Complex meshes A_a, B_b, G.
struct FieldLoader {
  typedef double2 VT;
  const double2* const* A;
  const double2* const* B;
  const double2* G;
  bool aligned16;   // every mesh 16-byte aligned: TMA bulk copies are usable
  int conj_b = 0;   // use conj(B) (B fields given as the mirror harmonic of the A fields)
  __device__ __forceinline__ double2 a(int ia, long long cell) const { return A[ia][cell]; }
  __device__ __forceinline__ double2 hb(int ib, long long cell) const {
    double2 b = B[ib][cell], g = G[cell];
    return make_double2(b.x * g.x - b.y * g.y, b.x * g.y + b.y * g.x);
  }
};

// Real meshes A_a, B_b, G (B_000-like statistics of a periodic box).
struct RealFieldLoader {
  typedef double VT;
  const double* const* A;
  const double* const* B;
  const double* G;
  bool aligned16;
  int conj_b = 0;   // no-op for real fields
  __device__ __forceinline__ double a(int ia, long long cell) const { return A[ia][cell]; }
  __device__ __forceinline__ double hb(int ib, long long cell) const { return B[ib][cell] * G[cell]; }
};

struct ShotLoader {
  typedef double2 VT;
  XView xi;
  GridDesc g;
  SjlView sja, sjb;
  const double* ka; const double* kb;
  YlmCoef ya_c, yb_c;
  __device__ __forceinline__ void rvec(long long cell, double& rx, double& ry, double& rz,
                                       double& r) const {
    const int k = (int)(cell % g.n[2]);
    const int j = (int)((cell / g.n[2]) % g.n[1]);
    const int i = (int)(cell / ((long long)g.n[2] * g.n[1]));
    // S/field.cpp:546-553: i*dr or (i-n)*dr.
    rx = __dmul_rn((double)signed_index(i, g.n[0]), g.dr[0]);
    ry = __dmul_rn((double)signed_index(j, g.n[1]), g.dr[1]);
    rz = __dmul_rn((double)signed_index(k, g.n[2]), g.dr[2]);
    r = vec3_norm_exact(rx, ry, rz);
  }
  __device__ __forceinline__ double2 a(int ia, long long cell) const {
    double rx, ry, rz, r; rvec(cell, rx, ry, rz, r);
    return make_double2(sjl_eval(sja, ka[ia] * r), 0.);
  }
  __device__ __forceinline__ double2 hb(int ib, long long cell) const {
    double rx, ry, rz, r; rvec(cell, rx, ry, rz, r);
    const double jb = sjl_eval(sjb, kb[ib] * r);
    cplx ya = ylm_eval(ya_c, rx, ry, rz);
    cplx yb = ylm_eval(yb_c, rx, ry, rz);
    cplx yy = cmul(ya, yb);
    double2 x = xload(xi, cell);
    cplx xv; xv.re = x.x; xv.im = x.y;
    cplx v = cmul(xv, yy);
    return make_double2(jb * v.re, jb * v.im);
  }
};

// Radially binned shot-noise mesh: for cubic cells |x| = dr sqrt(q) with the
// integer q = i^2 + j^2 + k^2 of the signed cell offset, so the N^3 cells
// collapse to <= 3 (n/2)^2 + 1 radii before any j_l is evaluated:
//   hist[q] = sum_{x : q(x) = q} y_a(xhat) y_b(xhat) xi(x).
// The eight cells (+-i, +-j, +-k) share q: they are loaded together (eight
// independent loads in flight per thread), summed in registers and sent as
// ONE RED.  TRIVIAL: y_a y_b = 1 (l_a = l_b = 0).
template <bool TRIVIAL>
__global__ void __launch_bounds__(256)
k_shot_radial_hist(XView xi, GridDesc g, int la, int ma, int lb,
                   int mb, double* __restrict__ hist) {
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const YlmCoef ya_c = ylm_coef(la, ma), yb_c = ylm_coef(lb, mb);
  for_each_cell(n0 / 2 + 1, n1 / 2 + 1, n2 / 2 + 1, [&](int ci, int cj, int ck, long long) {
    const int pi = ci ? n0 - ci : 0, pj = cj ? n1 - cj : 0, pk = ck ? n2 - ck : 0;
    int ii[2] = {ci, pi}, jj[2] = {cj, pj}, kk[2] = {ck, pk};
    double2 x[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int a = e >> 2, b = (e >> 1) & 1, c = e & 1;
      // A mirrored index equal to the original one is the same cell: skip it.
      const bool dup = (a && pi == ci) || (b && pj == cj) || (c && pk == ck);
      x[e] = make_double2(0., 0.);
      if (!dup) x[e] = xload(xi, ((long long)ii[a] * n1 + jj[b]) * n2 + kk[c]);
    }
    double re = 0., im = 0.;
    // y_lm of the eight mirror cells follow from the value at (ci, cj, ck) by parity
    // (ylm_mirror, bit-identical to direct evaluation): two evaluations per octant cell
    // instead of sixteen -- the pass was bound by their square roots and divisions.
    const bool by_parity = !TRIVIAL && !ya_c.generic && !yb_c.generic;
    cplx ya0, yb0;
    if (by_parity) {
      const double rx0 = (double)signed_index(ci, n0) * g.dr[0];
      const double ry0 = (double)signed_index(cj, n1) * g.dr[1];
      const double rz0 = (double)signed_index(ck, n2) * g.dr[2];
      ya0 = ylm_eval(ya_c, rx0, ry0, rz0); yb0 = ylm_eval(yb_c, rx0, ry0, rz0);
    }
#pragma unroll
    for (int e = 0; e < 8; e++) {
      if (TRIVIAL) { re += x[e].x; im += x[e].y; continue; }
      const int a = e >> 2, b = (e >> 1) & 1, c = e & 1;
      const bool dup = (a && pi == ci) || (b && pj == cj) || (c && pk == ck);
      if (dup) continue;
      cplx yy;
      if (by_parity) {
        yy = cmul(ylm_mirror(ya_c, ya0, a, b, c), ylm_mirror(yb_c, yb0, a, b, c));
      } else {
        const double rx = (double)signed_index(ii[a], n0) * g.dr[0];
        const double ry = (double)signed_index(jj[b], n1) * g.dr[1];
        const double rz = (double)signed_index(kk[c], n2) * g.dr[2];
        yy = cmul(ylm_eval(ya_c, rx, ry, rz), ylm_eval(yb_c, rx, ry, rz));
      }
      cplx xv; xv.re = x[e].x; xv.im = x[e].y;
      const cplx v = cmul(xv, yy);
      re += v.re; im += v.im;
    }
    const int si = signed_index(ci, n0), sj = signed_index(cj, n1), sk = signed_index(ck, n2);
    const long long q = (long long)si * si + (long long)sj * sj + (long long)sk * sk;
    atomicAdd(&hist[2 * q], re);
    if (xi.cplx || !TRIVIAL) atomicAdd(&hist[2 * q + 1], im);
  });
}

// The same histogram from the planes [x0, x0 + nx) of xi held by one rank of a distributed
// mesh (REAL, [nx][n1][n2]); the four cells (+-j, +-k) of a plane share q.
template <bool TRIVIAL>
__global__ void __launch_bounds__(256)
k_shot_radial_hist_slab(const double* __restrict__ xi, GridDesc g, int x0, int nx, int la, int ma,
                        int lb, int mb, double* __restrict__ hist) {
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const YlmCoef ya_c = ylm_coef(la, ma), yb_c = ylm_coef(lb, mb);
  for_each_cell(nx, n1 / 2 + 1, n2 / 2 + 1, [&](int xl, int cj, int ck, long long) {
    const int pj = cj ? n1 - cj : 0, pk = ck ? n2 - ck : 0;
    const int jj[2] = {cj, pj}, kk[2] = {ck, pk};
    const int si = signed_index(x0 + xl, n0);
    double re = 0., im = 0.;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int b = e >> 1, c = e & 1;
      if ((b && pj == cj) || (c && pk == ck)) continue;   // the same cell
      const double v = xi[((long long)xl * n1 + jj[b]) * n2 + kk[c]];
      if (TRIVIAL) { re += v; continue; }
      const double rx = (double)si * g.dr[0];
      const double ry = (double)signed_index(jj[b], n1) * g.dr[1];
      const double rz = (double)signed_index(kk[c], n2) * g.dr[2];
      const cplx yy = cmul(ylm_eval(ya_c, rx, ry, rz), ylm_eval(yb_c, rx, ry, rz));
      re += v * yy.re; im += v * yy.im;
    }
    const int sj = signed_index(cj, n1), sk = signed_index(ck, n2);
    const long long q = (long long)si * si + (long long)sj * sj + (long long)sk * sk;
    atomicAdd(&hist[2 * q], re);
    if (!TRIVIAL) atomicAdd(&hist[2 * q + 1], im);
  });
}

struct RadialLoader {
  typedef double2 VT;
  const double2* hist;
  double dr;
  SjlView sja, sjb;
  const double* ka; const double* kb;
  __device__ __forceinline__ double radius(long long q) const {
    // sqrt(q) dr reproduces the reference's |x| exactly on the axes and to an
    // ulp elsewhere (j_l is smooth: no discontinuous decision depends on it).
    return __dmul_rn(sqrt((double)q), dr);
  }
  __device__ __forceinline__ double2 a(int ia, long long q) const {
    return make_double2(sjl_eval(sja, ka[ia] * radius(q)), 0.);
  }
  __device__ __forceinline__ double2 hb(int ib, long long q) const {
    const double jb = sjl_eval(sjb, kb[ib] * radius(q));
    const double2 h = hist[q];
    return make_double2(jb * h.x, jb * h.y);
  }
};

// Pair blocks: block t covers a in [4 ta[t], +4) x b in [4 tb[t], +4) of the
// (padded) staged field lists.  Each warp owns up to TPW blocks.
//   sel_a / sel_b  staged field -> loader index (padded to multiples of 4)
//   partial        [gridDim.x][nblk][16] values (re, im)
template <class Loader, int TPW>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_gram(Loader ld, const int* __restrict__ sel_a, int na, const int* __restrict__ sel_b, int nb,
       long long ncells, const int* __restrict__ blk_a, const int* __restrict__ blk_b, int nblk,
       double* __restrict__ partial) {
  typedef typename Loader::VT VT;
  extern __shared__ double2 smem_raw[];
  VT* sA = reinterpret_cast<VT*>(smem_raw);   // [na][GRAM_T]
  VT* sB = sA + (size_t)na * GRAM_T;          // [nb][GRAM_T]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  VT acc[TPW][GRAM_B * GRAM_B];
  int a0[TPW], b0[TPW];
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
    a0[q] = (t < nblk) ? blk_a[t] * GRAM_B : 0;
    b0[q] = (t < nblk) ? blk_b[t] * GRAM_B : 0;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) vt_zero(acc[q][e]);
  }

  const long long ntiles = (ncells + GRAM_T - 1) / GRAM_T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * GRAM_T;
    __syncthreads();
    for (int idx = threadIdx.x; idx < na * GRAM_T; idx += GRAM_THREADS) {
      const int a = idx / GRAM_T, x = idx % GRAM_T;
      const long long cell = base + x;
      VT v; vt_zero(v);
      if (cell < ncells) v = ld.a(sel_a[a], cell);
      sA[idx] = v;
    }
    for (int idx = threadIdx.x; idx < nb * GRAM_T; idx += GRAM_THREADS) {
      const int b = idx / GRAM_T, x = idx % GRAM_T;
      const long long cell = base + x;
      VT v; vt_zero(v);
      if (cell < ncells) v = ld.hb(sel_b[b], cell);
      sB[idx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TPW; q++) {
      if (warp + q * GRAM_WARPS < nblk) {
        const VT* ra = sA + a0[q] * GRAM_T;
        const VT* rb = sB + b0[q] * GRAM_T;
#pragma unroll
        for (int h = 0; h < GRAM_T / 32; h++) {
          const int x = lane + 32 * h;
          VT va[GRAM_B], vb[GRAM_B];
#pragma unroll
          for (int e = 0; e < GRAM_B; e++) { va[e] = ra[e * GRAM_T + x]; vb[e] = rb[e * GRAM_T + x]; }
#pragma unroll
          for (int ea = 0; ea < GRAM_B; ea++)
#pragma unroll
            for (int eb = 0; eb < GRAM_B; eb++) vt_fma(acc[q][ea * GRAM_B + eb], va[ea], vb[eb]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) {
      VT v = acc[q][e];
      for (int o = 16; o > 0; o >>= 1) vt_add(v, vt_shfl_down(v, o));
      if (lane == 0 && t < nblk) {
        vt_store(partial + (((long long)blockIdx.x * nblk + t) * (GRAM_B * GRAM_B) + e) * 2, v);
      }
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" :: "n"(N));
}
__device__ __forceinline__ double vt_conj_if(double a, int) { return a; }
__device__ __forceinline__ double2 vt_conj_if(double2 a, int c) { if (c) a.y = -a.y; return a; }
__device__ __forceinline__ double vt_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 vt_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Same reduction for fields that are plain meshes in memory: raw tiles of
// A_a, B_b and G stream into a double-buffered shared-memory stage with
// cp.async (the next tile is in flight while the current one is reduced) and
// B_b * G is formed at use.
template <class VT, int TPW>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_gram_fields(const VT* const* __restrict__ A, const VT* const* __restrict__ B,
              const VT* __restrict__ G, const int* __restrict__ sel_a, int na,
              const int* __restrict__ sel_b, int nb, long long ncells,
              const int* __restrict__ blk_a, const int* __restrict__ blk_b, int nblk,
              double* __restrict__ partial, int conj_b) {
  constexpr int T = GRAM_T;
  constexpr int EPC = 16 / (int)sizeof(VT);   // elements per 16-byte chunk
  constexpr int CH = T / EPC;                 // chunks per staged row
  extern __shared__ double2 smem_raw[];
  const int rows = na + nb + 1;
  VT* buf0 = reinterpret_cast<VT*>(smem_raw);
  VT* buf1 = buf0 + (size_t)rows * T;
  const VT** s_ptr = reinterpret_cast<const VT**>(buf1 + (size_t)rows * T);
  for (int r = threadIdx.x; r < rows; r += GRAM_THREADS) {
    s_ptr[r] = r < na ? A[sel_a[r]] : (r < na + nb ? B[sel_b[r - na]] : G);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  VT acc[TPW][GRAM_B * GRAM_B];
  int a0[TPW], b0[TPW];
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
    a0[q] = (t < nblk) ? blk_a[t] * GRAM_B : 0;
    b0[q] = (t < nblk) ? blk_b[t] * GRAM_B : 0;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) vt_zero(acc[q][e]);
  }
  auto stage = [&](long long tile, VT* dst) {
    const long long base = tile * T;
    for (int idx = threadIdx.x; idx < rows * CH; idx += GRAM_THREADS) {
      const int r = idx / CH, ch = idx - r * CH;
      const long long cell = base + (long long)ch * EPC;
      VT* d = dst + (size_t)r * T + ch * EPC;
      const VT* src = s_ptr[r] + cell;
      if (cell + EPC <= ncells && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0) {
        cp_async16(d, src);
      } else {
#pragma unroll
        for (int e = 0; e < EPC; e++) {
          VT v; vt_zero(v);
          if (cell + e < ncells) v = src[e];
          d[e] = v;
        }
      }
    }
    cp_async_commit();
  };
  const long long ntiles = (ncells + T - 1) / T;
  if ((long long)blockIdx.x < ntiles) stage(blockIdx.x, buf0);
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    VT* cur = (it & 1) ? buf1 : buf0;
    VT* nxt = (it & 1) ? buf0 : buf1;
    const long long next = tile + gridDim.x;
    if (next < ntiles) stage(next, nxt); else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const VT* sA = cur;
    const VT* sB = (nb == 0) ? cur : cur + (size_t)na * T;   // nb == 0: B rows are the A rows
    const VT* sG = cur + (size_t)(na + nb) * T;
#pragma unroll
    for (int q = 0; q < TPW; q++) {
      if (warp + q * GRAM_WARPS < nblk) {
        const VT* ra = sA + a0[q] * T;
        const VT* rb = sB + b0[q] * T;
#pragma unroll
        for (int h = 0; h < T / 32; h++) {
          const int x = lane + 32 * h;
          const VT gx = sG[x];
          VT va[GRAM_B], vb[GRAM_B];
#pragma unroll
          for (int e = 0; e < GRAM_B; e++) { va[e] = ra[e * T + x]; vb[e] = vt_mul(vt_conj_if(rb[e * T + x], conj_b), gx); }
#pragma unroll
          for (int ea = 0; ea < GRAM_B; ea++)
#pragma unroll
            for (int eb = 0; eb < GRAM_B; eb++) vt_fma(acc[q][ea * GRAM_B + eb], va[ea], vb[eb]);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) {
      VT v = acc[q][e];
      for (int o = 16; o > 0; o >>= 1) vt_add(v, vt_shfl_down(v, o));
      if (lane == 0 && t < nblk) {
        vt_store(partial + (((long long)blockIdx.x * nblk + t) * (GRAM_B * GRAM_B) + e) * 2, v);
      }
    }
  }
}

// --- TMA (bulk async copy) + mbarrier primitives ----------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_LOOP:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE;\n"
    "bra WAIT_LOOP;\n"
    "DONE:\n"
    "}\n" :: "r"(a), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
// (SASS: UBLKCP).  dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes,
                                            unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
    :: "r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}

// The field reduction with the tile stage fed by the TMA engine: warp 0 issues
// one bulk copy per field row (64 cells) of the NEXT tile and arms that
// stage's mbarrier with the byte count; every warp waits on the barrier of the
// CURRENT stage, reduces its 4 x 4 pair blocks from it, and a block barrier
// hands the stage back.  No address arithmetic or staging instructions are
// spent by the compute warps.  Requires 16-byte aligned meshes.
template <class VT, int TPW, int T>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_gram_fields_tma(const VT* const* __restrict__ A, const VT* const* __restrict__ B,
                  const VT* __restrict__ G, const int* __restrict__ sel_a, int na,
                  const int* __restrict__ sel_b, int nb, long long ncells,
                  const int* __restrict__ blk_a, const int* __restrict__ blk_b, int nblk,
                  double* __restrict__ partial, int conj_b) {
  extern __shared__ __align__(128) double2 smem_raw[];
  const int rows = na + nb + 1;
  VT* buf0 = reinterpret_cast<VT*>(smem_raw);
  VT* buf1 = buf0 + (size_t)rows * T;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(buf1 + (size_t)rows * T);
  const VT** s_ptr = reinterpret_cast<const VT**>(bars + 2);
  for (int r = threadIdx.x; r < rows; r += GRAM_THREADS) {
    s_ptr[r] = r < na ? A[sel_a[r]] : (r < na + nb ? B[sel_b[r - na]] : G);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  VT acc[TPW][GRAM_B * GRAM_B];
  int a0[TPW], b0[TPW];
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
    a0[q] = (t < nblk) ? blk_a[t] * GRAM_B : 0;
    b0[q] = (t < nblk) ? blk_b[t] * GRAM_B : 0;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) vt_zero(acc[q][e]);
  }
  const long long ntiles = (ncells + T - 1) / T;
  auto issue = [&](long long tile, VT* dst, unsigned long long* bar) {   // warp 0 only
    const long long base = tile * T;
    const int valid = (int)min((long long)T, ncells - base);
    const unsigned row_bytes = (unsigned)(valid * sizeof(VT));
    if (lane == 0) mbar_arrive_expect_tx(bar, row_bytes * (unsigned)rows);
    __syncwarp();
    for (int r = lane; r < rows; r += 32) tma_load_1d(dst + (size_t)r * T, s_ptr[r] + base, row_bytes, bar);
  };
  if (warp == 0 && (long long)blockIdx.x < ntiles) issue(blockIdx.x, buf0, &bars[0]);
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    VT* cur = (it & 1) ? buf1 : buf0;
    const long long next = tile + gridDim.x;
    if (warp == 0 && next < ntiles) issue(next, (it & 1) ? buf0 : buf1, &bars[(it + 1) & 1]);
    mbar_wait(&bars[it & 1], (unsigned)((it >> 1) & 1));
    const int valid = (int)min((long long)T, ncells - tile * T);
    const VT* sA = cur;
    const VT* sB = (nb == 0) ? cur : cur + (size_t)na * T;   // nb == 0: B rows are the A rows
    const VT* sG = cur + (size_t)(na + nb) * T;
#pragma unroll
    for (int q = 0; q < TPW; q++) {
      if (warp + q * GRAM_WARPS < nblk) {
        const VT* ra = sA + a0[q] * T;
        const VT* rb = sB + b0[q] * T;
#pragma unroll
        for (int h = 0; h < T / 32; h++) {
          const int x = lane + 32 * h;
          if (x < valid) {
            const VT gx = sG[x];
            VT va[GRAM_B], vb[GRAM_B];
#pragma unroll
            for (int e = 0; e < GRAM_B; e++) { va[e] = ra[e * T + x]; vb[e] = vt_mul(vt_conj_if(rb[e * T + x], conj_b), gx); }
#pragma unroll
            for (int ea = 0; ea < GRAM_B; ea++)
#pragma unroll
              for (int eb = 0; eb < GRAM_B; eb++) vt_fma(acc[q][ea * GRAM_B + eb], va[ea], vb[eb]);
          }
        }
      }
    }
    __syncthreads();   // stage `cur` may be refilled from the next iteration on
  }
#pragma unroll
  for (int q = 0; q < TPW; q++) {
    const int t = warp + q * GRAM_WARPS;
#pragma unroll
    for (int e = 0; e < GRAM_B * GRAM_B; e++) {
      VT v = acc[q][e];
      for (int o = 16; o > 0; o >>= 1) vt_add(v, vt_shfl_down(v, o));
      if (lane == 0 && t < nblk) {
        vt_store(partial + (((long long)blockIdx.x * nblk + t) * (GRAM_B * GRAM_B) + e) * 2, v);
      }
    }
  }
}

template <class Loader> struct GramTraits {
  static constexpr bool kFields = false;
};
template <> struct GramTraits<FieldLoader> { static constexpr bool kFields = true; };
template <> struct GramTraits<RealFieldLoader> { static constexpr bool kFields = true; };

// Staged fields a launch may hold in shared memory (200 KB budget).
template <class Loader>
int gram_max_fields() {
  const size_t per_field = sizeof(typename Loader::VT) * GRAM_T;
  if (GramTraits<Loader>::kFields) return (int)((200 * 1024 - 4096) / (2 * per_field)) - 1;
  return (int)((200 * 1024) / per_field);
}

template <class Loader, int TPW>
int launch_gram_chunk(trvb_ctx* ctx, const Loader& ld, const int* d_sel_a, int na,
                      const int* d_sel_b, int nb, long long ncells, const int* d_blk_a,
                      const int* d_blk_b, int nblk, double* d_partial, int nblocks) {
  typedef typename Loader::VT VT;
  if constexpr (GramTraits<Loader>::kFields) {
    const size_t smem = sizeof(VT) * 2 * (size_t)(na + nb + 1) * GRAM_T + 16
      + sizeof(void*) * (size_t)(na + nb + 1);
    TRVB_REQUIRE(smem <= 200 * 1024, "gram reduce: %d + %d fields exceed the shared-memory stage", na, nb);
    if (ld.aligned16) {
      // Longest rows (cells per bulk copy) whose two stages fit in shared memory:
      // the TMA engine is fed per copy, so 2 KB rows beat 512 B rows.
      const size_t fixed = 16 + sizeof(void*) * (size_t)(na + nb + 1);
      auto stage_bytes = [&](int t) { return sizeof(VT) * 2 * (size_t)(na + nb + 1) * t + fixed; };
      auto launch_t = [&](auto tag) -> int {
        constexpr int TT = decltype(tag)::value;
        const size_t sm = stage_bytes(TT);
        const long long nt = (ncells + TT - 1) / TT;
        const int nbk = (int)std::min<long long>(nt, (long long)nblocks);
        TRVB_CUDA(cudaFuncSetAttribute(k_gram_fields_tma<VT, TPW, TT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        // Partials are laid out for `nblocks` blocks: idle ones must write zeros.
        if (nbk < nblocks) {
          TRVB_CUDA(cudaMemsetAsync(d_partial, 0, sizeof(double) * 2 * GRAM_B * GRAM_B * (size_t)nblocks * nblk,
                                    ctx->stream));
        }
        k_gram_fields_tma<VT, TPW, TT><<<nbk, GRAM_THREADS, sm, ctx->stream>>>(
          ld.A, ld.B, ld.G, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial,
          ld.conj_b);
        return 0;
      };
      int st = 0;
      if (stage_bytes(256) <= 200 * 1024) st = launch_t(std::integral_constant<int, 256>());
      else if (stage_bytes(128) <= 200 * 1024) st = launch_t(std::integral_constant<int, 128>());
      else st = launch_t(std::integral_constant<int, 64>());
      if (st) return st;
    } else {
      TRVB_CUDA(cudaFuncSetAttribute(k_gram_fields<VT, TPW>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_gram_fields<VT, TPW><<<nblocks, GRAM_THREADS, smem, ctx->stream>>>(
        ld.A, ld.B, ld.G, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial,
        ld.conj_b);
    }
  } else {
    const size_t smem = sizeof(VT) * (size_t)(na + nb) * GRAM_T;
    TRVB_REQUIRE(smem <= 200 * 1024, "gram reduce: %d + %d fields exceed the shared-memory tile", na, nb);
    TRVB_CUDA(cudaFuncSetAttribute(k_gram<Loader, TPW>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gram<Loader, TPW><<<nblocks, GRAM_THREADS, smem, ctx->stream>>>(
      ld, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial);
  }
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Runs the tiled reduction for `npairs` (ia, ib) pairs over `na_all` x
// `nb_all` loader fields and returns the complex sums in host array `out`
// (2 * npairs doubles).  Pairs are grouped into 4 x 4 blocks; launches cover
// up to GRAM_WARPS * 4 blocks and stage only the fields their blocks touch.
template <class Loader>
int run_gram(trvb_ctx* ctx, const Loader& ld, int na_all, int nb_all, long long ncells,
             const int* ia, const int* ib, int npairs, double* out, bool b_is_a = false) {
  TRVB_REQUIRE(na_all > 0 && nb_all > 0 && npairs > 0, "gram reduce: empty problem");
  constexpr int E = GRAM_B * GRAM_B;
  // Needed blocks in (a-block, b-block) order.
  const int nba = (na_all + GRAM_B - 1) / GRAM_B, nbb = (nb_all + GRAM_B - 1) / GRAM_B;
  std::vector<char> need((size_t)nba * nbb, 0);
  for (int p = 0; p < npairs; p++) need[(size_t)(ia[p] / GRAM_B) * nbb + ib[p] / GRAM_B] = 1;
  std::vector<std::pair<int, int> > blocks;
  for (int a = 0; a < nba; a++)
    for (int b = 0; b < nbb; b++) if (need[(size_t)a * nbb + b]) blocks.emplace_back(a, b);
  // Complex accumulators: 3 blocks per warp fit the register file without spills.
  constexpr int MAX_TPW = sizeof(typename Loader::VT) == sizeof(double) ? 4 : 3;
  const int max_blk = GRAM_WARPS * MAX_TPW;
  const long long ntiles = (ncells + GRAM_T - 1) / GRAM_T;
  const int nblocks = (int)std::min<long long>(ntiles, (long long)ctx->num_sms);

  // Chunk plan (host): per launch the staged field lists and block tables.
  struct Chunk { std::vector<int> sel_a, sel_b, blk_a, blk_b; size_t first; };
  std::vector<Chunk> chunks;
  const int max_fields = gram_max_fields<Loader>();
  for (size_t b0 = 0; b0 < blocks.size();) {
    Chunk c; c.first = b0;
    std::vector<int> slot_a(nba, -1), slot_b(nbb, -1);
    // When both sides are the same list of fields (b_is_a) a field is staged once:
    // the b-blocks take their rows from the a-list and the kernel sees nb == 0.
    std::vector<int>& slot_bb = b_is_a ? slot_a : slot_b;
    std::vector<int>& sel_bb = b_is_a ? c.sel_a : c.sel_b;
    size_t t = b0;
    for (; t < blocks.size() && (int)c.blk_a.size() < max_blk; t++) {
      const int a = blocks[t].first, b = blocks[t].second;
      const bool same_block = b_is_a && a == b;
      const int extra = (slot_a[a] < 0 ? GRAM_B : 0)
        + ((slot_bb[b] < 0 && !same_block) ? GRAM_B : 0);
      if (!c.blk_a.empty() && (int)(c.sel_a.size() + c.sel_b.size()) + extra > max_fields) break;
      if (slot_a[a] < 0) {
        slot_a[a] = (int)c.sel_a.size() / GRAM_B;
        for (int e = 0; e < GRAM_B; e++) c.sel_a.push_back(std::min(a * GRAM_B + e, na_all - 1));
      }
      if (slot_bb[b] < 0) {
        slot_bb[b] = (int)sel_bb.size() / GRAM_B;
        for (int e = 0; e < GRAM_B; e++) sel_bb.push_back(std::min(b * GRAM_B + e, nb_all - 1));
      }
      c.blk_a.push_back(slot_a[a]); c.blk_b.push_back(slot_bb[b]);
    }
    b0 = t;
    chunks.push_back(std::move(c));
  }
  // One scratch buffer: [int tables of all chunks][partials][block results].
  size_t nints = 0;
  for (const Chunk& c : chunks) nints += c.sel_a.size() + c.sel_b.size() + 2 * c.blk_a.size();
  const size_t bytes_tab = (sizeof(int) * nints + 255) / 256 * 256;
  const size_t bytes_partial = sizeof(double) * 2 * E * (size_t)nblocks * max_blk;
  const size_t bytes_res = sizeof(double) * 2 * E * blocks.size();
  double* scratch;
  int st = trvb_scratch(ctx, bytes_tab + bytes_partial + bytes_res + 512, &scratch);
  if (st) return st;
  int* d_tab = (int*)scratch;
  double* d_partial = (double*)((char*)scratch + bytes_tab);
  double* d_res = (double*)((char*)d_partial + bytes_partial);
  std::vector<int> h_tab; h_tab.reserve(nints);
  std::vector<size_t> off;
  for (const Chunk& c : chunks) {
    off.push_back(h_tab.size());
    h_tab.insert(h_tab.end(), c.sel_a.begin(), c.sel_a.end());
    h_tab.insert(h_tab.end(), c.sel_b.begin(), c.sel_b.end());
    h_tab.insert(h_tab.end(), c.blk_a.begin(), c.blk_a.end());
    h_tab.insert(h_tab.end(), c.blk_b.begin(), c.blk_b.end());
  }
  TRVB_CUDA(cudaMemcpyAsync(d_tab, h_tab.data(), sizeof(int) * h_tab.size(),
                            cudaMemcpyHostToDevice, ctx->stream));
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    const Chunk& c = chunks[ci];
    const int na = (int)c.sel_a.size(), nb = (int)c.sel_b.size(), nblk = (int)c.blk_a.size();
    const int* d_sel_a = d_tab + off[ci];
    const int* d_sel_b = d_sel_a + na;
    const int* d_blk_a = d_sel_b + nb;
    const int* d_blk_b = d_blk_a + nblk;
    if (nblk <= GRAM_WARPS) {
      st = launch_gram_chunk<Loader, 1>(ctx, ld, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial, nblocks);
    } else if (nblk <= GRAM_WARPS * 2) {
      st = launch_gram_chunk<Loader, 2>(ctx, ld, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial, nblocks);
    } else {
      st = launch_gram_chunk<Loader, MAX_TPW>(ctx, ld, d_sel_a, na, d_sel_b, nb, ncells, d_blk_a, d_blk_b, nblk, d_partial, nblocks);
    }
    if (st) return st;
    k_sum_cols<<<div_up(2 * E * nblk, 128), 128, 0, ctx->stream>>>(
      d_partial, nblocks, 2 * E * nblk, d_res + 2 * E * c.first);
    TRVB_LAUNCH_CHECK();
  }
  std::vector<double> h_res(2 * E * blocks.size());
  TRVB_CUDA(cudaMemcpyAsync(h_res.data(), d_res, bytes_res, cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int> block_of((size_t)nba * nbb, -1);
  for (size_t t = 0; t < blocks.size(); t++) block_of[(size_t)blocks[t].first * nbb + blocks[t].second] = (int)t;
  for (int p = 0; p < npairs; p++) {
    const int t = block_of[(size_t)(ia[p] / GRAM_B) * nbb + ib[p] / GRAM_B];
    const int e = (ia[p] % GRAM_B) * GRAM_B + ib[p] % GRAM_B;
    out[2 * p] = h_res[((size_t)t * E + e) * 2];
    out[2 * p + 1] = h_res[((size_t)t * E + e) * 2 + 1];
  }
  return 0;
}

// ---------------------------------------------------------------------
// Low-k cube iteration shared by shell statistics and binned 2-pt stats.
// ---------------------------------------------------------------------

struct Cube {
  int lo[3], cnt[3];
  long long total;
};

Cube cube_for(const GridDesc& g, double kmax) {
  Cube c;
  c.total = 1;
  for (int a = 0; a < 3; a++) {
    const int n = g.n[a];
    const int smin = -(n - n / 2), smax = n / 2 - 1;
    long long mc = (long long)std::floor(kmax / g.dk[a]) + 2;
    int lo = (int)std::max<long long>(smin, -mc);
    int hi = (int)std::min<long long>(smax, mc);
    if (n == 1) { lo = 0; hi = 0; }
    c.lo[a] = lo; c.cnt[a] = hi - lo + 1;
    c.total *= c.cnt[a];
  }
  return c;
}

struct BinRule {
  // fine == 0: mode in bin iff lo <= |k| < hi  (S/field.cpp:1826).
  // fine != 0: q = int(|k| / dsample) must satisfy qlo <= q < qhi, with the
  //            bounds precomputed on the host by the reference's own loop
  //            (S/field.cpp:2619, 2674-2676).
  int fine;
  double lo, hi;
  int qlo, qhi;
  double dsample;
  int nsample;
};

__device__ __forceinline__ bool in_bin(const BinRule& r, double v) {
  if (!r.fine) return r.lo <= v && v < r.hi;
  const int q = __double2int_rz(__ddiv_rn(v, r.dsample));
  return (0 <= q && q < r.nsample) && (r.qlo <= q && q < r.qhi);
}

// Cells of `cube` whose radius can fall in the bin: the rows (a, b) of the cube are dealt
// to the threads of blockIdx.x (1-D blocks), and on a row only the third coordinates with
// lo - slack <= |r| < hi + slack are visited (two short runs, one per sign), instead of
// the whole cube for every bin.  `f(s0, s1, s2)` receives signed indices and applies the
// exact bin rule itself, so the set of accepted cells is unchanged.
template <class F>
__device__ __forceinline__ void for_each_cell_near_shell(const Cube& cube, double d0, double d1,
                                                         double d2, const BinRule& rule, F f) {
  const double slack = (rule.fine ? rule.dsample : 0.) + 1.e-9 * rule.hi;
  const double rlo = fmax(rule.lo - slack, 0.), rhi = rule.hi + slack;
  const long long nrows = (long long)cube.cnt[0] * cube.cnt[1];
  const int cmin = cube.lo[2], cmax = cube.lo[2] + cube.cnt[2] - 1;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < nrows;
       row += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(row / cube.cnt[1]), b = (int)(row - (long long)a * cube.cnt[1]);
    const int s0 = cube.lo[0] + a, s1 = cube.lo[1] + b;
    const double x = s0 * d0, y = s1 * d1;
    const double rxy2 = x * x + y * y;
    const double zhi2 = rhi * rhi - rxy2;
    if (zhi2 < 0.) continue;
    const double zlo2 = rlo * rlo - rxy2;
    const int mhi = (int)(sqrt(zhi2) / d2) + 1;
    const int mlo = zlo2 > 0. ? max((int)(sqrt(zlo2) / d2) - 1, 0) : 0;
    for (int s2 = max(mlo, cmin); s2 <= min(mhi, cmax); s2++) f(s0, s1, s2);
    for (int s2 = max(-mhi, cmin); s2 <= min(-max(mlo, 1), cmax); s2++) f(s0, s1, s2);
  }
}

// blockIdx.y = bin.  partial layout: [bin][blockIdx.x][NQ].
template <int NQ>
__device__ __forceinline__ void store_partials(double (&v)[NQ], double* smem32,
                                               double* __restrict__ partial) {
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    double s = block_sum(v[q], smem32);
    if (threadIdx.x == 0) {
      partial[((long long)blockIdx.y * gridDim.x + blockIdx.x) * NQ + q] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
k_shell_stats(GridDesc g, Cube cube, const BinRule* __restrict__ rules,
              double* __restrict__ partial) {
  __shared__ double sm[32];
  const BinRule rule = rules[blockIdx.y];
  double v[2] = {0., 0.};
  for_each_cell_near_shell(cube, g.dk[0], g.dk[1], g.dk[2], rule, [&](int mi, int mj, int mk) {
    const double kmag = vec3_norm_exact(__dmul_rn((double)mi, g.dk[0]),
                                        __dmul_rn((double)mj, g.dk[1]),
                                        __dmul_rn((double)mk, g.dk[2]));
    if (in_bin(rule, kmag)) { v[0] += 1.; v[1] += kmag; }
  });
  store_partials<2>(v, sm, partial);
}

// Interlaced two-point statistics (S/field.cpp:2543-2552, 3504-3527): the mode
// is divided by the product of the two assignment windows,
// pow(sinc_x sinc_y sinc_z, order)^2, and the shot-noise aliasing function is the
// isotropic approximation cos^2(u/2)^order * C1_order(sin^2(u/2)), u = |pi m / n|.
__device__ __forceinline__ double window_sq(const Tables& tb, int order, int i, int j, int k) {
  const double wk = tb.sinc[0][i] * tb.sinc[1][j] * tb.sinc[2][k];
  double w = wk;
  for (int q = 1; q < order; q++) w *= wk;
  return w * w;
}

__device__ __forceinline__ double alias_iso(const GridDesc& g, int mi, int mj, int mk) {
  const double PI = 3.14159265358979323846;
  const double ux = PI * mi / double(g.n[0]);
  const double uy = PI * mj / double(g.n[1]);
  const double uz = PI * mk / double(g.n[2]);
  const double uhalf = sqrt(ux * ux + uy * uy + uz * uz) / 2.;
  const double s2h = sin(uhalf) * sin(uhalf), c2h = cos(uhalf) * cos(uhalf);
  double cc = 1.;
  for (int q = 0; q < g.order; q++) cc *= c2h;
  double cs = 1.;
  if (g.order == 2) cs = 1. - 2. / 3. * s2h;
  else if (g.order == 3) cs = 1. - s2h + 2. / 15. * s2h * s2h;
  else if (g.order == 4) cs = 1. - 4. / 3. * s2h + 2. / 5. * s2h * s2h - 4. / 315. * s2h * s2h * s2h;
  return cc * cs;
}

__global__ void __launch_bounds__(256)
k_twopt_fourier(KView fa, KView fb, GridDesc g, Tables tb, Cube cube,
                const BinRule* __restrict__ rules, double S_re, double S_im,
                int ell, int m, int interlaced, double* __restrict__ partial) {
  __shared__ double sm[32];
  const BinRule rule = rules[blockIdx.y];
  const YlmCoef yc = ylm_coef(ell, m);
  double v[6] = {0., 0., 0., 0., 0., 0.};
  for_each_cell_near_shell(cube, g.dk[0], g.dk[1], g.dk[2], rule, [&](int mi, int mj, int mk) {
    const double kx = __dmul_rn((double)mi, g.dk[0]);
    const double ky = __dmul_rn((double)mj, g.dk[1]);
    const double kz = __dmul_rn((double)mk, g.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    if (!in_bin(rule, kmag)) return;
    const int i = mi >= 0 ? mi : mi + g.n[0];
    const int j = mj >= 0 ? mj : mj + g.n[1];
    const int k = mk >= 0 ? mk : mk + g.n[2];
    const cplx a = kload(fa, i, j, k), b = kload(fb, i, j, k);
    // pk_mode = fa conj(fb) / C1 ; sn_mode = S C1 / C1  (S/field.cpp:2621-2633).
    double c1 = tb.alias[0][i] * tb.alias[1][j] * tb.alias[2][k];
    double win = c1;
    if (interlaced) { c1 = alias_iso(g, mi, mj, mk); win = window_sq(tb, g.order, i, j, k); }
    cplx pk; pk.re = (a.re * b.re + a.im * b.im) / win; pk.im = (a.im * b.re - a.re * b.im) / win;
    cplx sn; sn.re = (S_re * c1) / win; sn.im = (S_im * c1) / win;
    const cplx y = ylm_eval(yc, kx, ky, kz);
    pk = cmul(pk, y); sn = cmul(sn, y);
    v[0] += 1.; v[1] += kmag; v[2] += pk.re; v[3] += pk.im; v[4] += sn.re; v[5] += sn.im;
  });
  store_partials<6>(v, sm, partial);
}

// (fa conj(fb)/C1 - S C1/C1) / V on the full grid (S/field.cpp:3273-3298).
// `n2s` = stored extent of the last axis of dst: n2 (COMPLEX) or n2/2+1 (HALF).
template <bool INTERLACED>
__global__ void __launch_bounds__(256)
k_shot_spectrum(KView fa, KView fb, GridDesc g, Tables tb, double S_re, double S_im,
                int n2s, double2* __restrict__ dst) {
  const double inv_vol = 1. / g.vol;
  for_each_cell(g.n[0], g.n[1], n2s, [&](int i, int j, int k, long long t) {
    const cplx a = kload(fa, i, j, k), b = kload(fb, i, j, k);
    if constexpr (INTERLACED) {
      // (fa conj(fb) - S C1_iso) / (W_a W_b) / V, S/field.cpp:2760-2781.
      const int mi = i < g.n[0] / 2 ? i : i - g.n[0];   // S/field.cpp:538-544
      const int mj = j < g.n[1] / 2 ? j : j - g.n[1];
      const int mk = k < g.n[2] / 2 ? k : k - g.n[2];
      const double c1 = alias_iso(g, mi, mj, mk), win = window_sq(tb, g.order, i, j, k);
      const double re = (a.re * b.re + a.im * b.im) / win - (S_re * c1) / win;
      const double im = (a.im * b.re - a.re * b.im) / win - (S_im * c1) / win;
      dst[t] = make_double2(re * inv_vol, im * inv_vol);
      return;
    }
    // fa conj(fb)/C1 - S C1/C1, then /V, with tabulated per-axis reciprocals of
    // C1 instead of six fp64 divisions per mode (the pass was bound by the
    // divide sequence, not by HBM).
    const double rc1 = tb.ralias[0][i] * tb.ralias[1][j] * tb.ralias[2][k];
    const double re = (a.re * b.re + a.im * b.im) * rc1 - S_re;
    const double im = (a.im * b.re - a.re * b.im) * rc1 - S_im;
    dst[t] = make_double2(re * inv_vol, im * inv_vol);
  });
}

// Cells visited: the cube of signed offsets |i| dr <= r_max (bins are
// spherical shells, so the corners of the mesh never contribute).
__global__ void __launch_bounds__(256)
k_shot_3pcf_bin(XView xi, GridDesc g, Cube cube,
                const BinRule* __restrict__ rules, int la, int ma, int lb, int mb,
                double* __restrict__ partial) {
  __shared__ double sm[32];
  const BinRule rule = rules[blockIdx.y];
  const YlmCoef ya_c = ylm_coef(la, ma), yb_c = ylm_coef(lb, mb);
  double v[4] = {0., 0., 0., 0.};
  for_each_cell_near_shell(cube, g.dr[0], g.dr[1], g.dr[2], rule, [&](int si, int sj, int sk) {
    // S/field.cpp:546-553: i*dr or (i-n)*dr.
    const double rx = __dmul_rn((double)si, g.dr[0]);
    const double ry = __dmul_rn((double)sj, g.dr[1]);
    const double rz = __dmul_rn((double)sk, g.dr[2]);
    const double r = vec3_norm_exact(rx, ry, rz);
    if (!in_bin(rule, r)) return;
    const int i = si >= 0 ? si : si + g.n[0];
    const int j = sj >= 0 ? sj : sj + g.n[1];
    const int k = sk >= 0 ? sk : sk + g.n[2];
    const cplx ya = ylm_eval(ya_c, rx, ry, rz);
    const cplx yb = ylm_eval(yb_c, rx, ry, rz);
    const double2 x = xload(xi, ((long long)i * g.n[1] + j) * g.n[2] + k);
    cplx xv; xv.re = x.x; xv.im = x.y;
    const cplx val = cmul(xv, cmul(ya, yb));
    v[0] += 1.; v[1] += r; v[2] += val.re; v[3] += val.im;
  });
  store_partials<4>(v, sm, partial);
}

__global__ void __launch_bounds__(256)
k_sum_pow3(const double* __restrict__ p, long long n, int stride, int order,
           double* __restrict__ partial) {
  __shared__ double sm[32];
  double s = 0.;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = p[i * stride];
    double t = v * v;                       // order 2 (power spectrum) or 3 (bispectrum)
    for (int q = 2; q < order; q++) t *= v;
    s += t;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Host-side construction of bin rules; the fine-bin bounds follow the
// reference loop literally (S/field.cpp:2671-2682 / 3159-3169).
void make_rules(const double* edges, int nbins, int fine, double dsample, int nsample,
                std::vector<BinRule>& rules) {
  rules.resize(nbins);
  for (int b = 0; b < nbins; b++) {
    BinRule r;
    r.fine = fine; r.lo = edges[b]; r.hi = edges[b + 1];
    r.dsample = dsample; r.nsample = nsample;
    r.qlo = 0; r.qhi = 0;
    if (fine) {
      // Indices i with lo <= i*dsample < hi form the range [first(lo), first(hi))
      // because i*dsample is monotone in i; first(e) = min{i : i*dsample >= e}
      // is located by a short scan around e/dsample using the reference's own
      // comparison on i*dsample.
      auto first_ge = [&](double edge) {
        long long i = (long long)std::floor(edge / dsample) - 2;
        if (i < 0) i = 0;
        while (i < nsample && !((double)i * dsample >= edge)) i++;
        while (i > 0 && (double)(i - 1) * dsample >= edge) i--;
        return (int)std::min<long long>(i, nsample);
      };
      r.qlo = first_ge(r.lo);
      r.qhi = first_ge(r.hi);
      if (r.qhi < r.qlo) r.qhi = r.qlo;
    }
    rules[b] = r;
  }
}

// Launch a [blocks_x, nbins] binned kernel, reduce partials, copy to host.
template <int NQ, class Launch>
int run_binned(trvb_ctx* ctx, const std::vector<BinRule>& rules, long long nwork,
               Launch launch, std::vector<double>& host_out) {
  const int nbins = (int)rules.size();
  // `nwork` rows of the cube, one or two per thread
  const int bx = (int)std::max<long long>(1, std::min<long long>(div_up(nwork, 256),
                                          (long long)ctx->num_sms * 8));
  const size_t bytes_rules = (sizeof(BinRule) * nbins + 255) / 256 * 256;
  const size_t bytes_partial = sizeof(double) * (size_t)nbins * bx * NQ;
  const size_t bytes_out = sizeof(double) * (size_t)nbins * NQ;
  double* scratch;
  int st = trvb_scratch(ctx, bytes_rules + bytes_partial + bytes_out + 512, &scratch);
  if (st) return st;
  BinRule* d_rules = (BinRule*)scratch;
  double* d_partial = (double*)((char*)scratch + bytes_rules);
  double* d_out = (double*)((char*)d_partial + bytes_partial);
  TRVB_CUDA(cudaMemcpyAsync(d_rules, rules.data(), sizeof(BinRule) * nbins,
                            cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid(bx, nbins);
  launch(grid, d_rules, d_partial);
  TRVB_LAUNCH_CHECK();
  // partial is [bin][bx][NQ]; sum over bx for each (bin, q).
  k_sum_binned<<<nbins, 32, 0, ctx->stream>>>(d_partial, bx, NQ, d_out);
  TRVB_LAUNCH_CHECK();
  host_out.resize((size_t)nbins * NQ);
  TRVB_CUDA(cudaMemcpyAsync(host_out.data(), d_out, bytes_out, cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace

// =====================================================================
// Real-field Gram product on the FP64 tensor cores (DMMA, mma.sync m8n8k4)
// =====================================================================
//
// sum_x F_a(x) F_b(x) G(x) for all pairs is a skinny fp64 GEMM: C[a][b] += A[a][x] B'[x][b]
// with B'[x][b] = F_b(x) G(x) and x running over the cells.  The vector-FMA kernels above
// hold a private 4 x 4 pair block per lane and re-read eight operands from shared memory
// for every sixteen FMAs -- at 4.5 B per FMA the 128 B/clk of shared-memory bandwidth, not
// HBM, is what bounds them (C5: 29 ms for a 52 GB read).  With m8n8k4 the 8 x 8 block of C
// is spread over the warp (two doubles per lane), a fragment of eight fields x four cells
// is one 8-byte load per lane, and a warp that holds the fragments of all its field blocks
// issues every (a-block, b-block) product from registers: 6 loads for 15 DMMAs (3840 FMAs)
// on 40 fields, 0.4 B per FMA.  One warp owns ALL pair blocks; the eight warps of the CTA
// split the cells of the staged tile, so every field is read from HBM exactly once per
// launch and a launch covers up to 48 x 48 fields.
//
// Stage rows are padded by four doubles: the eight rows of a fragment then fall into
// distinct banks (2 wavefronts per 64-bit warp load, the minimum).

namespace {

constexpr int DM_WARPS = 8;                      // consumer warps: they split the cells of a tile
constexpr int DM_THREADS = (DM_WARPS + 1) * 32;  // + one producer warp feeding the TMA engine
constexpr int DM_PAD = 4;
constexpr int DM_MAXB = 6;        // 8-field blocks per launch, one list on both sides
constexpr int DM_MAXB_PAIR = 4;   // ... per side with two lists (or several groups)
constexpr int DM_MAX_STAGES = 6;
constexpr size_t DM_SMEM_LIMIT = 220 * 1024;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(a) : "memory");
}

// sel_a: 8 * nab field indices into A (padded by repetition), sel_b likewise into B
// (unused when SAME: the b-blocks are the a-blocks).  mask bit (8 i + j): block (i, j)
// wanted.  partial: [blockIdx][i * NB8 + j][64] (row-major 8 x 8 blocks).
//
// A ring of `nstages` tiles of T cells: the producer warp refills a stage as soon as the
// eight consumer warps have released it (`empty` barriers) and arms its `full` barrier
// with the byte count of the row copies; a consumer warp waits for `full`, takes its
// T / 8 cells through the tensor cores and releases the stage -- no block-wide barrier
// in the loop, several tiles in flight per SM.
template <int NB8, bool SAME, int T>
__global__ void __launch_bounds__(DM_THREADS, 1)
k_gram_dmma(const double* const* __restrict__ A, const double* const* __restrict__ B,
            const double* __restrict__ G, const int* __restrict__ sel_a, int nab,
            const int* __restrict__ sel_b, int nbb, long long ncells,
            unsigned long long mask, int nstages, double* __restrict__ partial) {
  constexpr int PITCH = T + DM_PAD;
  extern __shared__ __align__(128) double2 smem_raw[];
  const int na = 8 * nab, nb = SAME ? 0 : 8 * nbb;
  const int rows = na + nb + 1;
  const size_t stage_doubles = (size_t)rows * PITCH;
  double* ring = reinterpret_cast<double*>(smem_raw);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + stage_doubles * nstages);
  unsigned long long* empty = full + DM_MAX_STAGES;
  const double** s_ptr = reinterpret_cast<const double**>(empty + DM_MAX_STAGES);
  for (int r = threadIdx.x; r < rows; r += DM_THREADS) {
    s_ptr[r] = r < na ? A[sel_a[r]] : (r < na + nb ? B[sel_b[r - na]] : G);
  }
  if (threadIdx.x == 0) {
    for (int st = 0; st < nstages; st++) { mbar_init(&full[st], 1); mbar_init(&empty[st], DM_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntiles = (ncells + T - 1) / T;
  double acc[NB8 * NB8][2];
#pragma unroll
  for (int e = 0; e < NB8 * NB8; e++) { acc[e][0] = 0.; acc[e][1] = 0.; }

  if (warp == DM_WARPS) {
    // ---- producer ----
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
      const int st = n % nstages, round = n / nstages;
      if (round > 0) mbar_wait(&empty[st], (unsigned)((round - 1) & 1));
      const long long base = tile * T;
      const int valid = (int)min((long long)T, ncells - base);
      const unsigned row_bytes = (unsigned)(valid * sizeof(double));
      if (lane == 0) mbar_arrive_expect_tx(&full[st], row_bytes * (unsigned)rows);
      __syncwarp();
      double* dst = ring + stage_doubles * st;
      for (int r = lane; r < rows; r += 32) {
        tma_load_1d(dst + (size_t)r * PITCH, s_ptr[r] + base, row_bytes, &full[st]);
      }
    }
  } else {
    // ---- consumers ----
    const int frow = lane >> 2, fcol = lane & 3;
    constexpr int KSTEPS = T / (4 * DM_WARPS);
    int n = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, n++) {
      const int st = n % nstages, round = n / nstages;
      mbar_wait(&full[st], (unsigned)(round & 1));
      const double* cur = ring + stage_doubles * st;
      const int valid = (int)min((long long)T, ncells - tile * T);
      const double* sA = cur + (size_t)frow * PITCH;
      const double* sB = cur + (size_t)(na + frow) * PITCH;
      const double* sG = cur + (size_t)(na + nb) * PITCH;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ks++) {
        const int kk = (warp * KSTEPS + ks) * 4 + fcol;
        const bool live = kk < valid;
        const double g = live ? sG[kk] : 0.;
        double af[NB8], bf[NB8];
#pragma unroll
        for (int i = 0; i < NB8; i++) {
          af[i] = (live && i < nab) ? sA[(size_t)(8 * i) * PITCH + kk] : 0.;
        }
#pragma unroll
        for (int j = 0; j < NB8; j++) {
          if (SAME) bf[j] = af[j] * g;
          else bf[j] = ((live && j < nbb) ? sB[(size_t)(8 * j) * PITCH + kk] : 0.) * g;
        }
#pragma unroll
        for (int i = 0; i < NB8; i++) {
#pragma unroll
          for (int j = SAME ? i : 0; j < NB8; j++) {
            if ((mask >> (8 * i + j)) & 1ull) dmma884(acc[i * NB8 + j][0], acc[i * NB8 + j][1], af[i], bf[j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);   // this warp's reads of the stage are done
    }
  }
  __syncthreads();
  // Sum over the consumer warps in fixed order through the (now idle) ring.
  double* red = reinterpret_cast<double*>(smem_raw);   // [warp][NB8 * NB8][64]
  if (warp < DM_WARPS) {
    const int frow = lane >> 2, fcol = lane & 3;
#pragma unroll
    for (int e = 0; e < NB8 * NB8; e++) {
      double* d = red + ((size_t)warp * (NB8 * NB8) + e) * 64 + frow * 8 + 2 * fcol;
      d[0] = acc[e][0]; d[1] = acc[e][1];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NB8 * NB8 * 64; t += DM_THREADS) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < DM_WARPS; w++) v += red[(size_t)w * (NB8 * NB8 * 64) + t];
    partial[(size_t)blockIdx.x * (NB8 * NB8 * 64) + t] = v;
  }
}

template <int T>
size_t dmma_stage_bytes(int rows) { return sizeof(double) * (size_t)rows * (T + DM_PAD); }
inline size_t dmma_fixed_bytes(int rows) {
  return sizeof(unsigned long long) * 2 * DM_MAX_STAGES + sizeof(void*) * (size_t)rows;
}

template <int NB8, bool SAME, int T>
int launch_gram_dmma(trvb_ctx* ctx, const double* const* A, const double* const* B, const double* G,
                     const int* d_sel_a, int nab, const int* d_sel_b, int nbb, long long ncells,
                     unsigned long long mask, double* d_partial, int nblocks) {
  const int rows = 8 * nab + (SAME ? 0 : 8 * nbb) + 1;
  const size_t stage = dmma_stage_bytes<T>(rows), fixed = dmma_fixed_bytes(rows);
  int nstages = (int)std::min<size_t>(DM_MAX_STAGES, (DM_SMEM_LIMIT - fixed) / stage);
  {
    const char* env = getenv("TRV_GRAM_STAGES");
    if (env && atoi(env) >= 2) nstages = std::min(nstages, atoi(env));
  }
  TRVB_REQUIRE(nstages >= 2, "gram (dmma): %d staged rows exceed shared memory", rows);
  const size_t red = sizeof(double) * DM_WARPS * NB8 * NB8 * 64;
  const size_t sm = std::max(stage * nstages + fixed, red);
  TRVB_CUDA(cudaFuncSetAttribute(k_gram_dmma<NB8, SAME, T>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const long long nt = (ncells + T - 1) / T;
  const int nbk = (int)std::min<long long>(nt, (long long)nblocks);
  if (nbk < nblocks) {   // partials are laid out for `nblocks` blocks
    TRVB_CUDA(cudaMemsetAsync(d_partial, 0, sizeof(double) * NB8 * NB8 * 64 * (size_t)nblocks, ctx->stream));
  }
  k_gram_dmma<NB8, SAME, T><<<nbk, DM_THREADS, sm, ctx->stream>>>(
    A, B, G, d_sel_a, nab, d_sel_b, nbb, ncells, mask, nstages, d_partial);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Tile length: the longest rows (cells per bulk copy) of which two stages fit -- long rows
// feed the TMA engine better than a deeper ring of short ones (C5, 41 rows of 157 M cells:
// T = 320 / 256 with 2 stages 11.6 / 11.8 ms, T = 192 with 3 stages 11.9 ms, T = 128 with
// 2 / 3 / 6 stages 14.7 / 14.0 / 14.0 ms; profiles/r02_gram_dmma.txt).
template <int NB8, bool SAME>
int dispatch_gram_dmma_tile(trvb_ctx* ctx, const double* const* A, const double* const* B,
                            const double* G, const int* d_sel_a, int nab, const int* d_sel_b,
                            int nbb, long long ncells, unsigned long long mask, double* d_partial,
                            int nblocks) {
  const int rows = 8 * nab + (SAME ? 0 : 8 * nbb) + 1;
  const char* env = getenv("TRV_GRAM_TILE");
  const int want = env ? atoi(env) : 320;
  auto fits = [&](size_t stage) { return 2 * stage + dmma_fixed_bytes(rows) <= DM_SMEM_LIMIT; };
#define TRVB_DMMA_TILE(TT)                                                                      \
  if (want >= TT && fits(dmma_stage_bytes<TT>(rows)))                                           \
    return launch_gram_dmma<NB8, SAME, TT>(ctx, A, B, G, d_sel_a, nab, d_sel_b, nbb, ncells,    \
                                           mask, d_partial, nblocks);
  TRVB_DMMA_TILE(320)
  TRVB_DMMA_TILE(256)
  TRVB_DMMA_TILE(192)
#undef TRVB_DMMA_TILE
  return launch_gram_dmma<NB8, SAME, 128>(ctx, A, B, G, d_sel_a, nab, d_sel_b, nbb, ncells, mask,
                                          d_partial, nblocks);
}

template <int NB8>
int dispatch_gram_dmma(trvb_ctx* ctx, bool same, const double* const* A, const double* const* B,
                       const double* G, const int* d_sel_a, int nab, const int* d_sel_b, int nbb,
                       long long ncells, unsigned long long mask, double* d_partial, int nblocks) {
  if (same) {
    return dispatch_gram_dmma_tile<NB8, true>(ctx, A, B, G, d_sel_a, nab, d_sel_b, nbb, ncells,
                                              mask, d_partial, nblocks);
  }
  if constexpr (NB8 <= DM_MAXB_PAIR) {
    return dispatch_gram_dmma_tile<NB8, false>(ctx, A, B, G, d_sel_a, nab, d_sel_b, nbb, ncells,
                                               mask, d_partial, nblocks);
  } else {
    trvb_set_error("gram (dmma): %d blocks per side with two lists", NB8);
    return 2;
  }
}

// All `npairs` sums over real fields: groups of up to 48 fields per side and launch.
// d_A / d_B: device tables of field pointers; out: 2 * npairs doubles (imaginary parts 0).
int run_gram_dmma(trvb_ctx* ctx, const double* const* d_A, const double* const* d_B,
                  const double* G, int na_all, int nb_all, long long ncells, const int* ia,
                  const int* ib, int npairs, double* out, bool b_is_a) {
  const int nba = (na_all + 7) / 8, nbb_all = (nb_all + 7) / 8;
  // pair -> 8 x 8 block; with one list on both sides C is symmetric: upper blocks only
  std::vector<char> need((size_t)nba * nbb_all, 0);
  for (int p = 0; p < npairs; p++) {
    int a = ia[p] / 8, b = ib[p] / 8;
    if (b_is_a && a > b) std::swap(a, b);
    need[(size_t)a * nbb_all + b] = 1;
  }
  // one launch holds up to 48 fields when both sides are one list that fits, else groups
  // of 32 fields per side
  const int GB = (b_is_a && nba <= DM_MAXB) ? DM_MAXB : DM_MAXB_PAIR;
  const int nga = (nba + GB - 1) / GB, ngb = (nbb_all + GB - 1) / GB;
  struct Launch { int ga, gb, nab, nbb; bool same; unsigned long long mask; size_t res_off; int nb8; };
  std::vector<Launch> launches;
  size_t res_doubles = 0;
  for (int ga = 0; ga < nga; ga++) {
    for (int gb = 0; gb < ngb; gb++) {
      Launch L; L.ga = ga; L.gb = gb;
      L.nab = std::min(GB, nba - ga * GB);
      L.nbb = std::min(GB, nbb_all - gb * GB);
      L.same = b_is_a && ga == gb;
      L.mask = 0;
      for (int i = 0; i < L.nab; i++)
        for (int j = 0; j < L.nbb; j++)
          if (need[(size_t)(ga * GB + i) * nbb_all + gb * GB + j]) L.mask |= 1ull << (8 * i + j);
      if (!L.mask) continue;
      L.nb8 = std::max(L.nab, L.same ? L.nab : L.nbb);
      L.res_off = res_doubles;
      res_doubles += (size_t)L.nb8 * L.nb8 * 64;
      launches.push_back(L);
    }
  }
  const long long ntiles = (ncells + 127) / 128;
  const int nblocks = (int)std::min<long long>(ntiles, (long long)ctx->num_sms);
  const size_t sel_ints = 8 * (size_t)(nba + nbb_all);
  const size_t bytes_tab = (sizeof(int) * sel_ints + 255) / 256 * 256;
  const size_t bytes_partial = sizeof(double) * DM_MAXB * DM_MAXB * 64 * (size_t)nblocks;
  double* scratch;
  int st = trvb_scratch(ctx, bytes_tab + bytes_partial + sizeof(double) * res_doubles + 512, &scratch);
  if (st) return st;
  int* d_tab = (int*)scratch;
  double* d_partial = (double*)((char*)scratch + bytes_tab);
  double* d_res = d_partial + bytes_partial / sizeof(double);
  std::vector<int> h_tab(sel_ints);
  for (int r = 0; r < 8 * nba; r++) h_tab[r] = std::min(r, na_all - 1);
  for (int r = 0; r < 8 * nbb_all; r++) h_tab[8 * nba + r] = std::min(r, nb_all - 1);
  TRVB_CUDA(cudaMemcpyAsync(d_tab, h_tab.data(), sizeof(int) * h_tab.size(), cudaMemcpyHostToDevice,
                            ctx->stream));
  for (const Launch& L : launches) {
    const int* sa = d_tab + 8 * L.ga * GB;
    const int* sb = d_tab + 8 * nba + 8 * L.gb * GB;
#define TRVB_DMMA_CASE(N) case N: st = dispatch_gram_dmma<N>(ctx, L.same, d_A, d_B, G, sa, L.nab, sb, \
                                       L.nbb, ncells, L.mask, d_partial, nblocks); break;
    switch (L.nb8) {
      TRVB_DMMA_CASE(1) TRVB_DMMA_CASE(2) TRVB_DMMA_CASE(3)
      TRVB_DMMA_CASE(4) TRVB_DMMA_CASE(5) TRVB_DMMA_CASE(6)
      default: st = 2;
    }
#undef TRVB_DMMA_CASE
    if (st) return st;
    const int width = L.nb8 * L.nb8 * 64;
    k_sum_cols<<<div_up(width, 128), 128, 0, ctx->stream>>>(d_partial, nblocks, width, d_res + L.res_off);
    TRVB_LAUNCH_CHECK();
  }
  std::vector<double> h_res(res_doubles);
  TRVB_CUDA(cudaMemcpyAsync(h_res.data(), d_res, sizeof(double) * res_doubles, cudaMemcpyDeviceToHost,
                            ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int> launch_of((size_t)nga * ngb, -1);
  for (size_t t = 0; t < launches.size(); t++) launch_of[(size_t)launches[t].ga * ngb + launches[t].gb] = (int)t;
  for (int p = 0; p < npairs; p++) {
    int a = ia[p], b = ib[p];
    if (b_is_a && a / 8 > b / 8) std::swap(a, b);
    const Launch& L = launches[launch_of[(size_t)(a / 8 / GB) * ngb + b / 8 / GB]];
    const int i = a / 8 - L.ga * GB, j = b / 8 - L.gb * GB;
    out[2 * p] = h_res[L.res_off + ((size_t)(i * L.nb8 + j) * 64) + (a % 8) * 8 + b % 8];
    out[2 * p + 1] = 0.;
  }
  return 0;
}

}  // namespace

// =====================================================================
// C ABI
// =====================================================================

extern "C" int trvb_mesh_sum_pow3(trvb_ctx* ctx, trvb_mesh mesh, double* out) {
  return trvb_mesh_sum_pow(ctx, mesh, 3, out);
}

extern "C" int trvb_mesh_sum_pow(trvb_ctx* ctx, trvb_mesh mesh, int order, double* out) {
  TRVB_REQUIRE(ctx && mesh.data && out, "trvb_mesh_sum_pow: null argument");
  TRVB_REQUIRE(order >= 2, "trvb_mesh_sum_pow: order must be >= 2, got %d", order);
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_mesh_sum_pow: configuration-space layouts only");
  const int stride = mesh.layout == TRVB_COMPLEX ? 2 : 1;
  const int blocks = ctx->num_sms * 8;
  double* scratch;
  int st = trvb_scratch(ctx, sizeof(double) * (blocks + 8), &scratch);
  if (st) return st;
  k_sum_pow3<<<blocks, 256, 0, ctx->stream>>>((const double*)mesh.data, ctx->g.nmesh, stride, order, scratch);
  TRVB_LAUNCH_CHECK();
  k_sum_cols<<<1, 32, 0, ctx->stream>>>(scratch, blocks, 1, scratch + blocks);
  TRVB_LAUNCH_CHECK();
  TRVB_CUDA(cudaMemcpyAsync(out, scratch + blocks, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int trvb_gram_reduce(trvb_ctx* ctx, const void* const* A, int na,
                                const void* const* B, int nb, trvb_mesh G,
                                const int* ia, const int* ib, int npairs, int conj_b,
                                double* out) {
  TRVB_REQUIRE(ctx && A && B && G.data && ia && ib && out, "trvb_gram_reduce: null argument");
  TRVB_REQUIRE(G.layout == TRVB_COMPLEX || G.layout == TRVB_REAL,
               "trvb_gram_reduce: G must be a configuration-space mesh");
  for (int p = 0; p < npairs; p++) {
    TRVB_REQUIRE(ia[p] >= 0 && ia[p] < na && ib[p] >= 0 && ib[p] < nb,
                 "trvb_gram_reduce: pair %d = (%d, %d) out of range", p, ia[p], ib[p]);
  }
  // Pointer tables live in a small dedicated device buffer.
  const void** d_tab = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_tab, sizeof(void*) * (size_t)(na + nb)));
  TRVB_CUDA(cudaMemcpyAsync(d_tab, A, sizeof(void*) * na, cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaMemcpyAsync(d_tab + na, B, sizeof(void*) * nb, cudaMemcpyHostToDevice, ctx->stream));
  bool aligned = (reinterpret_cast<unsigned long long>(G.data) & 15ull) == 0;
  for (int i = 0; i < na; i++) aligned = aligned && (reinterpret_cast<unsigned long long>(A[i]) & 15ull) == 0;
  for (int i = 0; i < nb; i++) aligned = aligned && (reinterpret_cast<unsigned long long>(B[i]) & 15ull) == 0;
  // Real meshes with an odd cell count end in a half chunk: keep the cp.async path.
  if (G.layout == TRVB_REAL && (ctx->g.nmesh & 1)) aligned = false;
  const char* env_tma = getenv("TRV_GRAM_NO_TMA");
  if (env_tma && env_tma[0] == '1') aligned = false;
  // Both sides the same list of meshes (auto-correlation-like pair grids): stage once.
  bool b_is_a = na == nb;
  for (int i = 0; i < na && b_is_a; i++) b_is_a = A[i] == B[i];
  int st;
  if (G.layout == TRVB_REAL) {
    RealFieldLoader ld;
    ld.aligned16 = aligned;
    ld.A = (const double* const*)d_tab;
    ld.B = (const double* const*)(d_tab + na);
    ld.G = (const double*)G.data;
    const char* env_dmma = getenv("TRV_GRAM_NO_DMMA");
    if (aligned && !(env_dmma && env_dmma[0] == '1')) {
      st = run_gram_dmma(ctx, ld.A, ld.B, ld.G, na, nb, ctx->g.nmesh, ia, ib, npairs, out, b_is_a);
    } else {
      st = run_gram(ctx, ld, na, nb, ctx->g.nmesh, ia, ib, npairs, out, b_is_a);
    }
  } else {
    FieldLoader ld;
    ld.aligned16 = aligned;
    ld.conj_b = conj_b ? 1 : 0;
    ld.A = (const double2* const*)d_tab;
    ld.B = (const double2* const*)(d_tab + na);
    ld.G = (const double2*)G.data;
    st = run_gram(ctx, ld, na, nb, ctx->g.nmesh, ia, ib, npairs, out, b_is_a);
  }
  trvb_dev_free_raw(ctx, d_tab);
  return st;
}

namespace {

// The radial form of the shot-noise reduction as a real Gram product: rows
// F[f][q] = j_l(k_f r_q) for every distinct wavenumber and H[q] = Re hist[q], padded to an
// even number of radii; sum_q F_a F_b H for all pairs then runs on k_gram_dmma.  Empty
// radii (H = 0: most integers are not a small sum of three squares of mesh offsets ... or
// lie outside the mesh) are skipped: their products vanish whatever j_l is.
__global__ void __launch_bounds__(256)
k_radial_fields(const double2* __restrict__ hist, long long nq, long long pitch, double dr,
                SjlView sja, SjlView sjb, const double* __restrict__ ka, int na,
                const double* __restrict__ kb, int nb /* 0: the b rows are the a rows */,
                double* __restrict__ F /* [na + nb + 1][pitch], H last */) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < pitch;
       q += (long long)gridDim.x * blockDim.x) {
    const double h = q < nq ? hist[q].x : 0.;
    F[(long long)(na + nb) * pitch + q] = h;
    if (h == 0.) {
      for (int f = 0; f < na + nb; f++) F[(long long)f * pitch + q] = 0.;
      continue;
    }
    const double r = __dmul_rn(sqrt((double)q), dr);
    for (int f = 0; f < na; f++) F[(long long)f * pitch + q] = sjl_eval(sja, ka[f] * r);
    for (int f = 0; f < nb; f++) F[(long long)(na + f) * pitch + q] = sjl_eval(sjb, kb[f] * r);
  }
}

}  // namespace

namespace {
int shot_bispec_reduce_impl(trvb_ctx* ctx, trvb_mesh xi, int slab_x0, int slab_nx,
                            trvb_comm* comm, int la, int ma, int lb, int mb, const double* ka,
                            const double* kb, int npairs, double* out);
}

extern "C" int trvb_shot_bispec_reduce(trvb_ctx* ctx, trvb_mesh xi, int la, int ma,
                                       int lb, int mb, const double* ka, const double* kb,
                                       int npairs, double* out) {
  return shot_bispec_reduce_impl(ctx, xi, 0, 0, nullptr, la, ma, lb, mb, ka, kb, npairs, out);
}

extern "C" int trvb_shot_bispec_reduce_slab(trvb_ctx* ctx, const double* xi_planes, int x0, int nx,
                                            trvb_comm* comm, int la, int ma, int lb, int mb,
                                            const double* ka, const double* kb, int npairs,
                                            double* out) {
  TRVB_REQUIRE(ctx && comm && xi_planes && nx > 0 && x0 >= 0 && x0 + nx <= ctx->g.n[0],
               "trvb_shot_bispec_reduce_slab: bad argument");
  trvb_mesh xi; xi.data = const_cast<double*>(xi_planes); xi.layout = TRVB_REAL; xi.k0_add = 0.;
  return shot_bispec_reduce_impl(ctx, xi, x0, nx, comm, la, ma, lb, mb, ka, kb, npairs, out);
}

namespace {
// slab_nx > 0: `xi` holds the planes [slab_x0, slab_x0 + slab_nx) only (REAL) and the radial
// histogram is summed over the ranks of `comm` before the pair reduction.
int shot_bispec_reduce_impl(trvb_ctx* ctx, trvb_mesh xi, int slab_x0, int slab_nx,
                            trvb_comm* comm, int la, int ma, int lb, int mb, const double* ka,
                            const double* kb, int npairs, double* out) {
  TRVB_REQUIRE(ctx && xi.data && ka && kb && out && npairs > 0,
               "trvb_shot_bispec_reduce: bad argument");
  TRVB_REQUIRE(xi.layout == TRVB_COMPLEX || xi.layout == TRVB_REAL,
               "trvb_shot_bispec_reduce: xi must be a configuration-space mesh");
  XView xv; xv.p = (const double*)xi.data; xv.cplx = xi.layout == TRVB_COMPLEX;
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shot_bispec_reduce: root context only");
  auto ita = ctx->sjl.find(la), itb = ctx->sjl.find(lb);
  TRVB_REQUIRE(ita != ctx->sjl.end() && itb != ctx->sjl.end(),
               "trvb_shot_bispec_reduce: missing spline table (trvb_sjl_table)");
  // De-duplicate wavenumbers: a pair grid over N bins has only N distinct
  // values per side, so j_l is evaluated N times per cell, not npairs times.
  std::vector<double> ua, ub;
  std::vector<int> ia(npairs), ib(npairs);
  for (int p = 0; p < npairs; p++) {
    auto fa = std::find(ua.begin(), ua.end(), ka[p]);
    if (fa == ua.end()) { ua.push_back(ka[p]); ia[p] = (int)ua.size() - 1; }
    else ia[p] = (int)(fa - ua.begin());
    auto fb = std::find(ub.begin(), ub.end(), kb[p]);
    if (fb == ub.end()) { ub.push_back(kb[p]); ib[p] = (int)ub.size() - 1; }
    else ib[p] = (int)(fb - ub.begin());
  }
  double* d_k = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_k, sizeof(double) * (ua.size() + ub.size())));
  TRVB_CUDA(cudaMemcpyAsync(d_k, ua.data(), sizeof(double) * ua.size(), cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaMemcpyAsync(d_k + ua.size(), ub.data(), sizeof(double) * ub.size(), cudaMemcpyHostToDevice, ctx->stream));
  SjlView sja, sjb;
  sja.y = ita->second.d_y; sja.c = ita->second.d_c;
  sja.nsample = ita->second.nsample; sja.step = ita->second.step; sja.ell = la;
  sjb.y = itb->second.d_y; sjb.c = itb->second.d_c;
  sjb.nsample = itb->second.nsample; sjb.step = itb->second.step; sjb.ell = lb;
  const GridDesc& g = ctx->g;
  const char* env_direct = getenv("TRV_SHOT_DIRECT");
  const bool cubic = g.dr[0] == g.dr[1] && g.dr[1] == g.dr[2];
  const bool radial = cubic && !ctx->deterministic && !(env_direct && env_direct[0] == '1');
  TRVB_REQUIRE(slab_nx == 0 || radial, "trvb_shot_bispec_reduce_slab: needs cubic cells and the "
               "throughput mode (radial histogram)");
  int st;
  if (radial) {
    long long nq = 1;
    for (int a = 0; a < 3; a++) {
      const long long h = g.n[a] - g.n[a] / 2;   // largest |signed index|
      nq += h * h;
    }
    double* d_hist = nullptr;
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_hist, 2 * sizeof(double) * (size_t)nq));
    TRVB_CUDA(cudaMemsetAsync(d_hist, 0, 2 * sizeof(double) * (size_t)nq, ctx->stream));
    if (slab_nx > 0) {
      const RowLaunch rl = row_launch(ctx->num_sms, slab_nx, g.n[1] / 2 + 1, g.n[2] / 2 + 1);
      if (la == 0 && lb == 0) {
        k_shot_radial_hist_slab<true><<<rl.grid, rl.block, 0, ctx->stream>>>(
          xv.p, g, slab_x0, slab_nx, la, ma, lb, mb, d_hist);
      } else {
        k_shot_radial_hist_slab<false><<<rl.grid, rl.block, 0, ctx->stream>>>(
          xv.p, g, slab_x0, slab_nx, la, ma, lb, mb, d_hist);
      }
      TRVB_LAUNCH_CHECK();
      st = trvb_allreduce_device(ctx, comm, d_hist, 2 * nq);
      if (st) return st;
    } else {
      const RowLaunch rl = row_launch(ctx->num_sms, g.n[0] / 2 + 1, g.n[1] / 2 + 1, g.n[2] / 2 + 1);
      if (la == 0 && lb == 0) {
        k_shot_radial_hist<true><<<rl.grid, rl.block, 0, ctx->stream>>>(xv, g, la, ma, lb, mb, d_hist);
      } else {
        k_shot_radial_hist<false><<<rl.grid, rl.block, 0, ctx->stream>>>(xv, g, la, ma, lb, mb, d_hist);
      }
      TRVB_LAUNCH_CHECK();
    }
    const char* env_rd = getenv("TRV_SHOT_NO_DMMA");
    if (la == 0 && lb == 0 && !xv.cplx && !(env_rd && env_rd[0] == '1')) {
      // real histogram, real j_0: the pair sums are a real Gram product -> tensor cores
      const int na_u = (int)ua.size();
      const bool same = ua == ub;
      const int nb_u = same ? 0 : (int)ub.size();
      const long long pitch = (nq + 1) / 2 * 2;
      double* F = nullptr; const double** d_rows = nullptr;
      TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&F, sizeof(double) * (size_t)pitch * (na_u + nb_u + 1)));
      TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_rows, sizeof(void*) * (size_t)(na_u + nb_u)));
      std::vector<const double*> h_rows(na_u + nb_u);
      for (int f = 0; f < na_u + nb_u; f++) h_rows[f] = F + (size_t)f * pitch;
      TRVB_CUDA(cudaMemcpyAsync(d_rows, h_rows.data(), sizeof(void*) * h_rows.size(),
                                cudaMemcpyHostToDevice, ctx->stream));
      const int blocks = (int)std::min<long long>((pitch + 255) / 256, (long long)ctx->num_sms * 16);
      k_radial_fields<<<blocks, 256, 0, ctx->stream>>>((const double2*)d_hist, nq, pitch, g.dr[0],
                                                       sja, sjb, d_k, na_u, d_k + ua.size(), nb_u, F);
      TRVB_LAUNCH_CHECK();
      st = run_gram_dmma(ctx, d_rows, same ? d_rows : d_rows + na_u, F + (size_t)(na_u + nb_u) * pitch,
                         na_u, (int)ub.size(), pitch, ia.data(), ib.data(), npairs, out, same);
      trvb_dev_free_raw(ctx, F);
      trvb_dev_free_raw(ctx, d_rows);
    } else {
      RadialLoader ld;
      ld.hist = (const double2*)d_hist; ld.dr = g.dr[0];
      ld.sja = sja; ld.sjb = sjb; ld.ka = d_k; ld.kb = d_k + ua.size();
      st = run_gram(ctx, ld, (int)ua.size(), (int)ub.size(), nq, ia.data(), ib.data(), npairs, out);
    }
    trvb_dev_free_raw(ctx, d_hist);
  } else {
    ShotLoader ld;
    ld.xi = xv; ld.g = g;
    ld.sja = sja; ld.sjb = sjb;
    ld.ka = d_k; ld.kb = d_k + ua.size();
    ld.ya_c = ylm_coef(la, ma); ld.yb_c = ylm_coef(lb, mb);
    st = run_gram(ctx, ld, (int)ua.size(), (int)ub.size(), g.nmesh, ia.data(), ib.data(), npairs, out);
  }
  trvb_dev_free_raw(ctx, d_k);
  if (st) return st;
  for (int p = 0; p < 2 * npairs; p++) out[p] *= ctx->g.vol_cell;   // S/field.cpp:3393
  return 0;
}
}  // namespace

extern "C" int trvb_shell_stats(trvb_ctx* ctx, const double* edges, int nbins, int fine,
                                long long* nmodes, double* ksum) {
  TRVB_REQUIRE(ctx && edges && nmodes && ksum && nbins > 0, "trvb_shell_stats: bad argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shell_stats: root context only");
  std::vector<BinRule> rules;
  make_rules(edges, nbins, fine, 1.e-5, 1000000, rules);
  const GridDesc g = ctx->g;
  const Cube cube = cube_for(g, edges[nbins] + 2.e-5);
  std::vector<double> host;
  int st = run_binned<2>(ctx, rules, (long long)cube.cnt[0] * cube.cnt[1],
    [&](dim3 grid, const BinRule* d_rules, double* d_partial) {
      k_shell_stats<<<grid, 256, 0, ctx->stream>>>(g, cube, d_rules, d_partial);
    }, host);
  if (st) return st;
  for (int b = 0; b < nbins; b++) {
    nmodes[b] = (long long)llround(host[2 * b]);
    ksum[b] = host[2 * b + 1];
  }
  return 0;
}

extern "C" int trvb_twopt_fourier(trvb_ctx* ctx, trvb_mesh fa, trvb_mesh fb,
                                  const double S[2], int ell, int m, int interlaced,
                                  const double* edges, const double* centres, int nbins,
                                  long long* nmodes, double* k, double* pk, double* sn) {
  TRVB_REQUIRE(ctx && fa.data && fb.data && S && edges && centres && nmodes && k && pk && sn,
               "trvb_twopt_fourier: null argument");
  TRVB_REQUIRE(fa.layout != TRVB_REAL && fb.layout != TRVB_REAL,
               "trvb_twopt_fourier: Fourier-space meshes required");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_twopt_fourier: root context only");
  std::vector<BinRule> rules;
  make_rules(edges, nbins, 1, 1.e-5, 1000000, rules);
  const GridDesc g = ctx->g;
  const Cube cube = cube_for(g, edges[nbins] + 2.e-5);
  const KView va = kview_of(ctx, fa), vb = kview_of(ctx, fb);
  const Tables tb = tables_of(ctx);
  std::vector<double> host;
  int st = run_binned<6>(ctx, rules, (long long)cube.cnt[0] * cube.cnt[1],
    [&](dim3 grid, const BinRule* d_rules, double* d_partial) {
      k_twopt_fourier<<<grid, 256, 0, ctx->stream>>>(va, vb, g, tb, cube, d_rules,
                                                    S[0], S[1], ell, m, interlaced, d_partial);
    }, host);
  if (st) return st;
  for (int b = 0; b < nbins; b++) {
    const double* h = &host[6 * b];
    const long long nm = (long long)llround(h[0]);
    nmodes[b] = nm;
    if (nm != 0) {   // S/field.cpp:2684-2692
      k[b] = h[1] / double(nm);
      pk[2 * b] = h[2] / double(nm); pk[2 * b + 1] = h[3] / double(nm);
      sn[2 * b] = h[4] / double(nm); sn[2 * b + 1] = h[5] / double(nm);
    } else {
      k[b] = centres[b];
      pk[2 * b] = pk[2 * b + 1] = sn[2 * b] = sn[2 * b + 1] = 0.;
    }
  }
  return 0;
}

extern "C" int trvb_shot_xi(trvb_ctx* ctx, trvb_mesh fa, trvb_mesh fb, const double S[2],
                            int interlaced, trvb_mesh dst) {
  TRVB_REQUIRE(ctx && fa.data && fb.data && S && dst.data, "trvb_shot_xi: null argument");
  TRVB_REQUIRE(fa.layout != TRVB_REAL && fb.layout != TRVB_REAL,
               "trvb_shot_xi: Fourier-space inputs required");
  TRVB_REQUIRE(dst.layout == TRVB_COMPLEX
               || (dst.layout == TRVB_REAL && fa.layout == TRVB_HALF && fb.layout == TRVB_HALF
                   && S[1] == 0.),
               "trvb_shot_xi: dst must be COMPLEX (or REAL for two HALF inputs and real S)");
  TRVB_REQUIRE(dst.data != fa.data && dst.data != fb.data, "trvb_shot_xi: dst aliases an input");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shot_xi: root context only");
  const GridDesc& g = ctx->g;
  if (dst.layout == TRVB_REAL) {
    // Hermitian product: build the half spectrum in a temporary, Z2D into dst.
    trvb_mesh half; half.layout = TRVB_HALF; half.k0_add = 0.; half.data = nullptr;
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, &half.data, trvb_mesh_bytes(ctx, TRVB_HALF)));
    const RowLaunch rl = row_launch(ctx->num_sms, g.n[0], g.n[1], g.nh);
    if (interlaced) {
      k_shot_spectrum<true><<<rl.grid, rl.block, 0, ctx->stream>>>(
        kview_of(ctx, fa), kview_of(ctx, fb), g, tables_of(ctx), S[0], S[1], g.nh,
        (double2*)half.data);
    } else {
      k_shot_spectrum<false><<<rl.grid, rl.block, 0, ctx->stream>>>(
        kview_of(ctx, fa), kview_of(ctx, fb), g, tables_of(ctx), S[0], S[1], g.nh,
        (double2*)half.data);
    }
    TRVB_LAUNCH_CHECK();
    int st = trvb_fft_inverse(ctx, half, dst);
    trvb_dev_free_raw(ctx, half.data);
    return st;
  }
  const RowLaunch rl = row_launch(ctx->num_sms, g.n[0], g.n[1], g.n[2]);
  if (interlaced) {
    k_shot_spectrum<true><<<rl.grid, rl.block, 0, ctx->stream>>>(
      kview_of(ctx, fa), kview_of(ctx, fb), g, tables_of(ctx), S[0], S[1], g.n[2],
      (double2*)dst.data);
  } else {
    k_shot_spectrum<false><<<rl.grid, rl.block, 0, ctx->stream>>>(
      kview_of(ctx, fa), kview_of(ctx, fb), g, tables_of(ctx), S[0], S[1], g.n[2],
      (double2*)dst.data);
  }
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(ctx, dst, dst);
}

namespace {

// Per-bin {count, sum r, sum Re, sum Im} of y_{la ma}(r) y_{lb mb}(r) xi(r) over the
// cells whose separation falls in the bin by the reference's sampled-histogram rule
// (index int(r / dsample) < nsample, S/field.cpp:2862-2921 and 3095-3169).
int binned_xi_sums(trvb_ctx* ctx, trvb_mesh xi, int la, int ma, int lb, int mb,
                   const double* edges, int nbins, double dsample, int nsample,
                   std::vector<double>& host) {
  std::vector<BinRule> rules;
  make_rules(edges, nbins, 1, dsample, nsample, rules);
  const GridDesc g = ctx->g;
  XView d_xi; d_xi.p = (const double*)xi.data; d_xi.cplx = xi.layout == TRVB_COMPLEX;
  // Real-space analogue of cube_for: signed offsets with |i| dr <= r_max + slack.
  Cube cube; cube.total = 1;
  for (int a = 0; a < 3; a++) {
    const int n = g.n[a];
    const int smin = -(n - n / 2), smax = n / 2 - 1;
    const long long mc = (long long)std::floor((edges[nbins] + 2.) / g.dr[a]) + 2;
    int lo = (int)std::max<long long>(smin, -mc), hi = (int)std::min<long long>(smax, mc);
    if (n == 1) { lo = 0; hi = 0; }
    cube.lo[a] = lo; cube.cnt[a] = hi - lo + 1; cube.total *= cube.cnt[a];
  }
  return run_binned<4>(ctx, rules, (long long)cube.cnt[0] * cube.cnt[1],
    [&](dim3 grid, const BinRule* d_rules, double* d_partial) {
      k_shot_3pcf_bin<<<grid, 256, 0, ctx->stream>>>(d_xi, g, cube, d_rules, la, ma, lb, mb, d_partial);
    }, host);
}

}  // namespace

extern "C" int trvb_twopt_config_bin(trvb_ctx* ctx, trvb_mesh xi, int ell, int m,
                                     const double* edges, const double* centres, int nbins,
                                     long long* npairs, double* r, double* xi_out) {
  TRVB_REQUIRE(ctx && xi.data && edges && centres && npairs && r && xi_out,
               "trvb_twopt_config_bin: null argument");
  TRVB_REQUIRE(xi.layout == TRVB_COMPLEX || xi.layout == TRVB_REAL,
               "trvb_twopt_config_bin: xi must be a configuration-space mesh");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_twopt_config_bin: root context only");
  std::vector<double> host;
  // dr_sample = 0.1, n_sample = 1e6 (S/field.cpp:2846-2847); y_00 = 1 exactly.
  int st = binned_xi_sums(ctx, xi, ell, m, 0, 0, edges, nbins, 1.e-1, 1000000, host);
  if (st) return st;
  for (int b = 0; b < nbins; b++) {
    const double* h = &host[4 * b];
    const long long np = (long long)llround(h[0]);
    npairs[b] = np;
    if (np != 0) {   // S/field.cpp:2923-2931
      r[b] = h[1] / double(np);
      xi_out[2 * b] = h[2] / double(np); xi_out[2 * b + 1] = h[3] / double(np);
    } else {
      r[b] = centres[b];
      xi_out[2 * b] = 0.; xi_out[2 * b + 1] = 0.;
    }
  }
  return 0;
}

extern "C" int trvb_shot_3pcf_bin(trvb_ctx* ctx, trvb_mesh xi, int la, int ma, int lb,
                                  int mb, const double* edges, const double* centres,
                                  int nbins, double parity, long long* npairs, double* r,
                                  double* xi_out) {
  TRVB_REQUIRE(ctx && xi.data && edges && centres && npairs && r && xi_out,
               "trvb_shot_3pcf_bin: null argument");
  TRVB_REQUIRE(xi.layout == TRVB_COMPLEX || xi.layout == TRVB_REAL,
               "trvb_shot_3pcf_bin: xi must be a configuration-space mesh");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shot_3pcf_bin: root context only");
  const GridDesc g = ctx->g;
  std::vector<double> host;
  // dr_sample = 1, n_sample = 1e5 (S/field.cpp:3095-3096).
  int st = binned_xi_sums(ctx, xi, la, ma, lb, mb, edges, nbins, 1., 100000, host);
  if (st) return st;
  const double norm_factors = 1 / g.vol_cell * parity;   // S/field.cpp:3181-3182
  for (int b = 0; b < nbins; b++) {
    const double* h = &host[4 * b];
    const long long np = (long long)llround(h[0]);
    npairs[b] = np;
    if (np != 0) {
      r[b] = h[1] / double(np);
      double re = h[2] / double(np), im = h[3] / double(np);   // S/field.cpp:3173
      re *= norm_factors / double(np);                         // S/field.cpp:3186 (F5d)
      im *= norm_factors / double(np);
      xi_out[2 * b] = re; xi_out[2 * b + 1] = im;
    } else {
      r[b] = centres[b];
      xi_out[2 * b] = 0.; xi_out[2 * b + 1] = 0.;
    }
  }
  return 0;
}
