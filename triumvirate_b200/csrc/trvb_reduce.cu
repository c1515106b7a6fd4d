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

namespace {

struct Tables {
  const double* sinc[3];
  const double* alias[3];
};

Tables tables_of(const trvb_ctx* ctx) {
  const trvb_ctx* root = ctx->parent ? ctx->parent : ctx;
  Tables t;
  for (int a = 0; a < 3; a++) { t.sinc[a] = root->d_sinc[a]; t.alias[a] = root->d_alias[a]; }
  return t;
}

KView kview_of(const trvb_ctx* ctx, trvb_mesh m) {
  KView v;
  v.p = (const double2*)m.data; v.layout = m.layout;
  v.n0 = ctx->g.n[0]; v.n1 = ctx->g.n[1]; v.n2 = ctx->g.n[2]; v.nh = ctx->g.nh;
  return v;
}

__global__ void k_sum_cols(const double* __restrict__ partial, int nblocks, int width,
                           double* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= width) return;
  double s = 0.;
  for (int b = 0; b < nblocks; b++) s += partial[(long long)b * width + t];
  out[t] = s;
}

// ---------------------------------------------------------------------
// Tiled all-pairs ("Gram") reduction.
// ---------------------------------------------------------------------

constexpr int GRAM_T = 64;        // cells per tile (two per lane)
constexpr int GRAM_WARPS = 8;
constexpr int GRAM_THREADS = GRAM_WARPS * 32;

struct FieldLoader {
  const double2* const* A;
  const double2* const* B;
  const double2* G;
  __device__ __forceinline__ double2 a(int ia, long long cell) const { return A[ia][cell]; }
  __device__ __forceinline__ double2 hb(int ib, long long cell) const {
    double2 b = B[ib][cell], g = G[cell];
    return make_double2(b.x * g.x - b.y * g.y, b.x * g.y + b.y * g.x);
  }
};

struct ShotLoader {
  const double2* xi;
  GridDesc g;
  SjlView sja, sjb;
  const double* ka; const double* kb;
  int la, ma, lb, mb;
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
    cplx ya = ylm_reduced(la, ma, rx, ry, rz);
    cplx yb = ylm_reduced(lb, mb, rx, ry, rz);
    cplx yy = cmul(ya, yb);
    double2 x = xi[cell];
    cplx xv; xv.re = x.x; xv.im = x.y;
    cplx v = cmul(xv, yy);
    return make_double2(jb * v.re, jb * v.im);
  }
};

// Radially binned shot-noise mesh: for cubic cells |x| = dr sqrt(q) with the
// integer q = i^2 + j^2 + k^2 of the signed cell offset, so the N^3 cells
// collapse to <= 3 (n/2)^2 + 1 radii before any j_l is evaluated:
//   hist[q] = sum_{x : q(x) = q} y_a(xhat) y_b(xhat) xi(x).
__global__ void __launch_bounds__(256)
k_shot_radial_hist(const double2* __restrict__ xi, GridDesc g, int la, int ma, int lb,
                   int mb, double* __restrict__ hist) {
  const bool trivial = (la == 0 && lb == 0);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = signed_index((int)(t % g.n[2]), g.n[2]);
    const int j = signed_index((int)((t / g.n[2]) % g.n[1]), g.n[1]);
    const int i = signed_index((int)(t / ((long long)g.n[2] * g.n[1])), g.n[0]);
    const long long q = (long long)i * i + (long long)j * j + (long long)k * k;
    const double2 x = xi[t];
    double re = x.x, im = x.y;
    if (!trivial) {
      const double rx = (double)i * g.dr[0], ry = (double)j * g.dr[1], rz = (double)k * g.dr[2];
      const cplx yy = cmul(ylm_reduced(la, ma, rx, ry, rz), ylm_reduced(lb, mb, rx, ry, rz));
      cplx xv; xv.re = re; xv.im = im;
      const cplx v = cmul(xv, yy);
      re = v.re; im = v.im;
    }
    atomicAdd(&hist[2 * q], re);
    atomicAdd(&hist[2 * q + 1], im);
  }
}

struct RadialLoader {
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

// Each warp owns up to PPW pairs; each lane owns two cells of the tile and
// keeps PPW complex accumulators in registers.
template <class Loader, int PPW>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_gram(Loader ld, int na, int nb, long long ncells, const int* __restrict__ pair_ia,
       const int* __restrict__ pair_ib, int npairs, double* __restrict__ partial) {
  extern __shared__ double2 smem[];
  double2* sA = smem;                         // [na][GRAM_T]
  double2* sB = smem + (size_t)na * GRAM_T;   // [nb][GRAM_T]
  __shared__ short s_ia[GRAM_WARPS * PPW], s_ib[GRAM_WARPS * PPW];
  for (int p = threadIdx.x; p < GRAM_WARPS * PPW; p += GRAM_THREADS) {
    s_ia[p] = (p < npairs) ? (short)pair_ia[p] : (short)0;
    s_ib[p] = (p < npairs) ? (short)pair_ib[p] : (short)0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc_re[PPW], acc_im[PPW];
#pragma unroll
  for (int q = 0; q < PPW; q++) { acc_re[q] = 0.; acc_im[q] = 0.; }

  const long long ntiles = (ncells + GRAM_T - 1) / GRAM_T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * GRAM_T;
    __syncthreads();
    for (int idx = threadIdx.x; idx < na * GRAM_T; idx += GRAM_THREADS) {
      const int a = idx / GRAM_T, x = idx % GRAM_T;
      const long long cell = base + x;
      sA[idx] = (cell < ncells) ? ld.a(a, cell) : make_double2(0., 0.);
    }
    for (int idx = threadIdx.x; idx < nb * GRAM_T; idx += GRAM_THREADS) {
      const int b = idx / GRAM_T, x = idx % GRAM_T;
      const long long cell = base + x;
      sB[idx] = (cell < ncells) ? ld.hb(b, cell) : make_double2(0., 0.);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PPW; q++) {
      const int p = warp + q * GRAM_WARPS;
      if (p < npairs) {
        const double2* ra = sA + (int)s_ia[p] * GRAM_T;
        const double2* rb = sB + (int)s_ib[p] * GRAM_T;
        const double2 a0 = ra[lane], b0 = rb[lane];
        const double2 a1 = ra[lane + 32], b1 = rb[lane + 32];
        acc_re[q] += a0.x * b0.x - a0.y * b0.y;
        acc_im[q] += a0.x * b0.y + a0.y * b0.x;
        acc_re[q] += a1.x * b1.x - a1.y * b1.y;
        acc_im[q] += a1.x * b1.y + a1.y * b1.x;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PPW; q++) {
    double re = acc_re[q], im = acc_im[q];
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_down_sync(0xffffffffu, re, o);
      im += __shfl_down_sync(0xffffffffu, im, o);
    }
    const int p = warp + q * GRAM_WARPS;
    if (lane == 0 && p < npairs) {
      partial[((long long)blockIdx.x * npairs + p) * 2] = re;
      partial[((long long)blockIdx.x * npairs + p) * 2 + 1] = im;
    }
  }
}

template <class Loader, int PPW>
int launch_gram_chunk(trvb_ctx* ctx, const Loader& ld, int na, int nb, long long ncells,
                      const int* d_ia, const int* d_ib, int npairs, double* d_partial,
                      int nblocks) {
  const size_t smem = sizeof(double2) * (size_t)(na + nb) * GRAM_T;
  TRVB_REQUIRE(smem <= 200 * 1024, "gram reduce: %d + %d fields exceed the shared-memory tile", na, nb);
  TRVB_CUDA(cudaFuncSetAttribute(k_gram<Loader, PPW>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gram<Loader, PPW><<<nblocks, GRAM_THREADS, smem, ctx->stream>>>(
    ld, na, nb, ncells, d_ia, d_ib, npairs, d_partial);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Runs the tiled reduction over all pairs (chunks of <= 256 pairs) and
// returns the complex sums in host array `out` (2 * npairs doubles).
template <class Loader>
int run_gram(trvb_ctx* ctx, const Loader& ld, int na, int nb, long long ncells,
             const int* ia, const int* ib, int npairs, double* out) {
  TRVB_REQUIRE(na > 0 && nb > 0 && npairs > 0, "gram reduce: empty problem");
  TRVB_REQUIRE(na < 32768 && nb < 32768, "gram reduce: too many fields");
  const long long ntiles = (ncells + GRAM_T - 1) / GRAM_T;
  const int nblocks = (int)std::min<long long>(ntiles, (long long)ctx->num_sms);
  const int max_chunk = GRAM_WARPS * 32;
  // Scratch: pair tables (ints) + partials + result.
  const size_t bytes_pairs = sizeof(int) * 2 * (size_t)npairs;
  const size_t pad_pairs = (bytes_pairs + 255) / 256 * 256;
  const size_t bytes_partial = sizeof(double) * 2 * (size_t)nblocks * max_chunk;
  const size_t bytes_out = sizeof(double) * 2 * (size_t)npairs;
  double* scratch;
  int st = trvb_scratch(ctx, pad_pairs + bytes_partial + bytes_out + 512, &scratch);
  if (st) return st;
  int* d_ia = (int*)scratch;
  int* d_ib = d_ia + npairs;
  double* d_partial = (double*)((char*)scratch + pad_pairs);
  double* d_out = (double*)((char*)d_partial + bytes_partial);
  TRVB_CUDA(cudaMemcpyAsync(d_ia, ia, sizeof(int) * npairs, cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaMemcpyAsync(d_ib, ib, sizeof(int) * npairs, cudaMemcpyHostToDevice, ctx->stream));
  for (int p0 = 0; p0 < npairs; p0 += max_chunk) {
    const int np = std::min(max_chunk, npairs - p0);
    if (np <= GRAM_WARPS * 4) {
      st = launch_gram_chunk<Loader, 4>(ctx, ld, na, nb, ncells, d_ia + p0, d_ib + p0, np, d_partial, nblocks);
    } else if (np <= GRAM_WARPS * 16) {
      st = launch_gram_chunk<Loader, 16>(ctx, ld, na, nb, ncells, d_ia + p0, d_ib + p0, np, d_partial, nblocks);
    } else {
      st = launch_gram_chunk<Loader, 32>(ctx, ld, na, nb, ncells, d_ia + p0, d_ib + p0, np, d_partial, nblocks);
    }
    if (st) return st;
    k_sum_cols<<<div_up(2 * np, 128), 128, 0, ctx->stream>>>(d_partial, nblocks, 2 * np, d_out + 2 * p0);
    TRVB_LAUNCH_CHECK();
  }
  TRVB_CUDA(cudaMemcpyAsync(out, d_out, bytes_out, cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
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
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cube.total;
       t += (long long)gridDim.x * blockDim.x) {
    const int mk = cube.lo[2] + (int)(t % cube.cnt[2]);
    const int mj = cube.lo[1] + (int)((t / cube.cnt[2]) % cube.cnt[1]);
    const int mi = cube.lo[0] + (int)(t / ((long long)cube.cnt[2] * cube.cnt[1]));
    const double kmag = vec3_norm_exact(__dmul_rn((double)mi, g.dk[0]),
                                        __dmul_rn((double)mj, g.dk[1]),
                                        __dmul_rn((double)mk, g.dk[2]));
    if (in_bin(rule, kmag)) { v[0] += 1.; v[1] += kmag; }
  }
  store_partials<2>(v, sm, partial);
}

__global__ void __launch_bounds__(256)
k_twopt_fourier(KView fa, KView fb, GridDesc g, Tables tb, Cube cube,
                const BinRule* __restrict__ rules, double S_re, double S_im,
                int ell, int m, double* __restrict__ partial) {
  __shared__ double sm[32];
  const BinRule rule = rules[blockIdx.y];
  double v[6] = {0., 0., 0., 0., 0., 0.};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cube.total;
       t += (long long)gridDim.x * blockDim.x) {
    const int mk = cube.lo[2] + (int)(t % cube.cnt[2]);
    const int mj = cube.lo[1] + (int)((t / cube.cnt[2]) % cube.cnt[1]);
    const int mi = cube.lo[0] + (int)(t / ((long long)cube.cnt[2] * cube.cnt[1]));
    const double kx = __dmul_rn((double)mi, g.dk[0]);
    const double ky = __dmul_rn((double)mj, g.dk[1]);
    const double kz = __dmul_rn((double)mk, g.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    if (!in_bin(rule, kmag)) continue;
    const int i = mi >= 0 ? mi : mi + g.n[0];
    const int j = mj >= 0 ? mj : mj + g.n[1];
    const int k = mk >= 0 ? mk : mk + g.n[2];
    const cplx a = kload(fa, i, j, k), b = kload(fb, i, j, k);
    // pk_mode = fa conj(fb) / C1 ; sn_mode = S C1 / C1  (S/field.cpp:2621-2633).
    const double c1 = tb.alias[0][i] * tb.alias[1][j] * tb.alias[2][k];
    cplx pk; pk.re = (a.re * b.re + a.im * b.im) / c1; pk.im = (a.im * b.re - a.re * b.im) / c1;
    cplx sn; sn.re = (S_re * c1) / c1; sn.im = (S_im * c1) / c1;
    const cplx y = ylm_reduced(ell, m, kx, ky, kz);
    pk = cmul(pk, y); sn = cmul(sn, y);
    v[0] += 1.; v[1] += kmag; v[2] += pk.re; v[3] += pk.im; v[4] += sn.re; v[5] += sn.im;
  }
  store_partials<6>(v, sm, partial);
}

// (fa conj(fb)/C1 - S C1/C1) / V on the full grid (S/field.cpp:3273-3298).
__global__ void __launch_bounds__(256)
k_shot_spectrum(KView fa, KView fb, GridDesc g, Tables tb, double S_re, double S_im,
                double2* __restrict__ dst) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % g.n[2]);
    const int j = (int)((t / g.n[2]) % g.n[1]);
    const int i = (int)(t / ((long long)g.n[2] * g.n[1]));
    const cplx a = kload(fa, i, j, k), b = kload(fb, i, j, k);
    const double c1 = tb.alias[0][i] * tb.alias[1][j] * tb.alias[2][k];
    double re = (a.re * b.re + a.im * b.im) / c1 - (S_re * c1) / c1;
    double im = (a.im * b.re - a.re * b.im) / c1 - (S_im * c1) / c1;
    dst[t] = make_double2(re / g.vol, im / g.vol);
  }
}

__global__ void __launch_bounds__(256)
k_shot_3pcf_bin(const double2* __restrict__ xi, GridDesc g,
                const BinRule* __restrict__ rules, int la, int ma, int lb, int mb,
                double* __restrict__ partial) {
  __shared__ double sm[32];
  const BinRule rule = rules[blockIdx.y];
  double v[4] = {0., 0., 0., 0.};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % g.n[2]);
    const int j = (int)((t / g.n[2]) % g.n[1]);
    const int i = (int)(t / ((long long)g.n[2] * g.n[1]));
    const double rx = __dmul_rn((double)signed_index(i, g.n[0]), g.dr[0]);
    const double ry = __dmul_rn((double)signed_index(j, g.n[1]), g.dr[1]);
    const double rz = __dmul_rn((double)signed_index(k, g.n[2]), g.dr[2]);
    const double r = vec3_norm_exact(rx, ry, rz);
    if (!in_bin(rule, r)) continue;
    const cplx ya = ylm_reduced(la, ma, rx, ry, rz);
    const cplx yb = ylm_reduced(lb, mb, rx, ry, rz);
    const double2 x = xi[t];
    cplx xv; xv.re = x.x; xv.im = x.y;
    const cplx val = cmul(xv, cmul(ya, yb));
    v[0] += 1.; v[1] += r; v[2] += val.re; v[3] += val.im;
  }
  store_partials<4>(v, sm, partial);
}

__global__ void __launch_bounds__(256)
k_sum_pow3(const double* __restrict__ p, long long n, int stride,
           double* __restrict__ partial) {
  __shared__ double sm[32];
  double s = 0.;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = p[i * stride];
    s += v * v * v;
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
  const int bx = (int)std::max<long long>(1, std::min<long long>(div_up(nwork, 256 * 4),
                                          (long long)ctx->num_sms * 4));
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
  for (int b = 0; b < nbins; b++) {
    k_sum_cols<<<1, 32, 0, ctx->stream>>>(d_partial + (size_t)b * bx * NQ, bx, NQ, d_out + (size_t)b * NQ);
    g_trvb_launches++;
  }
  TRVB_CUDA(cudaGetLastError());
  host_out.resize((size_t)nbins * NQ);
  TRVB_CUDA(cudaMemcpyAsync(host_out.data(), d_out, bytes_out, cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace

// =====================================================================
// C ABI
// =====================================================================

extern "C" int trvb_mesh_sum_pow3(trvb_ctx* ctx, trvb_mesh mesh, double* out) {
  TRVB_REQUIRE(ctx && mesh.data && out, "trvb_mesh_sum_pow3: null argument");
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_mesh_sum_pow3: configuration-space layouts only");
  const int stride = mesh.layout == TRVB_COMPLEX ? 2 : 1;
  const int blocks = ctx->num_sms * 8;
  double* scratch;
  int st = trvb_scratch(ctx, sizeof(double) * (blocks + 8), &scratch);
  if (st) return st;
  k_sum_pow3<<<blocks, 256, 0, ctx->stream>>>((const double*)mesh.data, ctx->g.nmesh, stride, scratch);
  TRVB_LAUNCH_CHECK();
  k_sum_cols<<<1, 32, 0, ctx->stream>>>(scratch, blocks, 1, scratch + blocks);
  TRVB_LAUNCH_CHECK();
  TRVB_CUDA(cudaMemcpyAsync(out, scratch + blocks, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int trvb_gram_reduce(trvb_ctx* ctx, const void* const* A, int na,
                                const void* const* B, int nb, trvb_mesh G,
                                const int* ia, const int* ib, int npairs, double* out) {
  TRVB_REQUIRE(ctx && A && B && G.data && ia && ib && out, "trvb_gram_reduce: null argument");
  TRVB_REQUIRE(G.layout == TRVB_COMPLEX, "trvb_gram_reduce: G must be a COMPLEX mesh");
  for (int p = 0; p < npairs; p++) {
    TRVB_REQUIRE(ia[p] >= 0 && ia[p] < na && ib[p] >= 0 && ib[p] < nb,
                 "trvb_gram_reduce: pair %d = (%d, %d) out of range", p, ia[p], ib[p]);
  }
  // Pointer tables live in a small dedicated device buffer.
  const void** d_tab = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_tab, sizeof(void*) * (size_t)(na + nb)));
  TRVB_CUDA(cudaMemcpyAsync(d_tab, A, sizeof(void*) * na, cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaMemcpyAsync(d_tab + na, B, sizeof(void*) * nb, cudaMemcpyHostToDevice, ctx->stream));
  FieldLoader ld;
  ld.A = (const double2* const*)d_tab;
  ld.B = (const double2* const*)(d_tab + na);
  ld.G = (const double2*)G.data;
  int st = run_gram(ctx, ld, na, nb, ctx->g.nmesh, ia, ib, npairs, out);
  cudaStreamSynchronize(ctx->stream);
  trvb_dev_free_raw(ctx, d_tab);
  return st;
}

extern "C" int trvb_shot_bispec_reduce(trvb_ctx* ctx, trvb_mesh xi, int la, int ma,
                                       int lb, int mb, const double* ka, const double* kb,
                                       int npairs, double* out) {
  TRVB_REQUIRE(ctx && xi.data && ka && kb && out && npairs > 0,
               "trvb_shot_bispec_reduce: bad argument");
  TRVB_REQUIRE(xi.layout == TRVB_COMPLEX, "trvb_shot_bispec_reduce: xi must be COMPLEX");
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
    const int blocks = (int)std::min<long long>(div_up(g.nmesh, 256), (long long)ctx->num_sms * 16);
    k_shot_radial_hist<<<blocks, 256, 0, ctx->stream>>>((const double2*)xi.data, g, la, ma, lb, mb, d_hist);
    TRVB_LAUNCH_CHECK();
    RadialLoader ld;
    ld.hist = (const double2*)d_hist; ld.dr = g.dr[0];
    ld.sja = sja; ld.sjb = sjb; ld.ka = d_k; ld.kb = d_k + ua.size();
    st = run_gram(ctx, ld, (int)ua.size(), (int)ub.size(), nq, ia.data(), ib.data(), npairs, out);
    trvb_dev_free_raw(ctx, d_hist);
  } else {
    ShotLoader ld;
    ld.xi = (const double2*)xi.data; ld.g = g;
    ld.sja = sja; ld.sjb = sjb;
    ld.ka = d_k; ld.kb = d_k + ua.size();
    ld.la = la; ld.ma = ma; ld.lb = lb; ld.mb = mb;
    st = run_gram(ctx, ld, (int)ua.size(), (int)ub.size(), g.nmesh, ia.data(), ib.data(), npairs, out);
  }
  trvb_dev_free_raw(ctx, d_k);
  if (st) return st;
  for (int p = 0; p < 2 * npairs; p++) out[p] *= ctx->g.vol_cell;   // S/field.cpp:3393
  return 0;
}

extern "C" int trvb_shell_stats(trvb_ctx* ctx, const double* edges, int nbins, int fine,
                                long long* nmodes, double* ksum) {
  TRVB_REQUIRE(ctx && edges && nmodes && ksum && nbins > 0, "trvb_shell_stats: bad argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shell_stats: root context only");
  std::vector<BinRule> rules;
  make_rules(edges, nbins, fine, 1.e-5, 1000000, rules);
  const GridDesc g = ctx->g;
  const Cube cube = cube_for(g, edges[nbins] + 2.e-5);
  std::vector<double> host;
  int st = run_binned<2>(ctx, rules, cube.total,
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
                                  const double S[2], int ell, int m, const double* edges,
                                  const double* centres, int nbins, long long* nmodes,
                                  double* k, double* pk, double* sn) {
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
  int st = run_binned<6>(ctx, rules, cube.total,
    [&](dim3 grid, const BinRule* d_rules, double* d_partial) {
      k_twopt_fourier<<<grid, 256, 0, ctx->stream>>>(va, vb, g, tb, cube, d_rules,
                                                    S[0], S[1], ell, m, d_partial);
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
                            trvb_mesh dst) {
  TRVB_REQUIRE(ctx && fa.data && fb.data && S && dst.data, "trvb_shot_xi: null argument");
  TRVB_REQUIRE(fa.layout != TRVB_REAL && fb.layout != TRVB_REAL && dst.layout == TRVB_COMPLEX,
               "trvb_shot_xi: Fourier-space inputs and a COMPLEX output required");
  TRVB_REQUIRE(dst.data != fa.data && dst.data != fb.data, "trvb_shot_xi: dst aliases an input");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shot_xi: root context only");
  const int blocks = (int)std::min<long long>(div_up(ctx->g.nmesh, 256), (long long)ctx->num_sms * 32);
  k_shot_spectrum<<<blocks, 256, 0, ctx->stream>>>(
    kview_of(ctx, fa), kview_of(ctx, fb), ctx->g, tables_of(ctx), S[0], S[1], (double2*)dst.data);
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(ctx, dst, dst);
}

extern "C" int trvb_shot_3pcf_bin(trvb_ctx* ctx, trvb_mesh xi, int la, int ma, int lb,
                                  int mb, const double* edges, const double* centres,
                                  int nbins, double parity, long long* npairs, double* r,
                                  double* xi_out) {
  TRVB_REQUIRE(ctx && xi.data && edges && centres && npairs && r && xi_out,
               "trvb_shot_3pcf_bin: null argument");
  TRVB_REQUIRE(xi.layout == TRVB_COMPLEX, "trvb_shot_3pcf_bin: xi must be COMPLEX");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_shot_3pcf_bin: root context only");
  std::vector<BinRule> rules;
  make_rules(edges, nbins, 1, 1., 100000, rules);   // S/field.cpp:3095-3096
  const GridDesc g = ctx->g;
  const double2* d_xi = (const double2*)xi.data;
  std::vector<double> host;
  int st = run_binned<4>(ctx, rules, g.nmesh,
    [&](dim3 grid, const BinRule* d_rules, double* d_partial) {
      k_shot_3pcf_bin<<<grid, 256, 0, ctx->stream>>>(d_xi, g, d_rules, la, ma, lb, mb, d_partial);
    }, host);
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
