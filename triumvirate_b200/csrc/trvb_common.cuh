// trvb_common.cuh -- shared state and device helpers for libtrvb.so
// (sm_100a; CUDA 12.9).  See include/trvb.h for the C-ABI this implements.
#ifndef TRVB_COMMON_CUH_
#define TRVB_COMMON_CUH_

#include <cuda_runtime.h>

#include <atomic>
#include <cufft.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "trvb.h"

// ---------------------------------------------------------------------
// Error handling: every C-ABI entry returns an int and stores a message.
// ---------------------------------------------------------------------

void trvb_set_error(const char* fmt, ...);
extern std::atomic<long long> g_trvb_launches;    // hand-written kernels launched
extern std::atomic<long long> g_trvb_fft_execs;   // cuFFT executions (library launches)

#define TRVB_CUDA(call)                                                    \
  do {                                                                     \
    cudaError_t err__ = (call);                                            \
    if (err__ != cudaSuccess) {                                            \
      trvb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,         \
                     cudaGetErrorString(err__));                           \
      return 100 + (int)err__;                                             \
    }                                                                      \
  } while (0)

#define TRVB_CUFFT(call)                                                   \
  do {                                                                     \
    cufftResult err__ = (call);                                            \
    if (err__ != CUFFT_SUCCESS) {                                          \
      trvb_set_error("%s:%d: %s -> cufftResult %d", __FILE__, __LINE__,    \
                     #call, (int)err__);                                   \
      return 1000 + (int)err__;                                            \
    }                                                                      \
  } while (0)

#define TRVB_REQUIRE(cond, ...)                                            \
  do {                                                                     \
    if (!(cond)) {                                                         \
      trvb_set_error(__VA_ARGS__);                                         \
      return 2;                                                            \
    }                                                                      \
  } while (0)

#define TRVB_LAUNCH_CHECK()                                                \
  do {                                                                     \
    g_trvb_launches++;                                                     \
    TRVB_CUDA(cudaGetLastError());                                         \
  } while (0)

// ---------------------------------------------------------------------
// Grid description passed by value to kernels.
// ---------------------------------------------------------------------

struct GridDesc {
  int n[3];            // cells per axis
  int nh;              // n[2]/2 + 1 (half-spectrum extent)
  long long nmesh;     // n0*n1*n2
  double L[3];         // box size
  double dr[3];        // L/n      (S/field.cpp:353-355)
  double dk[3];        // 2 pi / L (S/field.cpp:358-360)
  double vol;          // L0 L1 L2
  double vol_cell;     // vol / nmesh (S/field.cpp:363-364)
  int order;           // assignment order 1..4
};

struct SjlTable {
  double* d_y = nullptr;
  double* d_c = nullptr;
  std::vector<double> h_y, h_c;   // host copies: skip re-uploading an identical table
  int nsample = 0;
  double step = 0.;
};

struct trvb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;          // sub-grid context with a stream of its own
  cudaEvent_t fork_event = nullptr; // orders that stream against the parent's
  GridDesc g;
  trvb_ctx* parent = nullptr;       // non-null for a sub-grid context
  // Sub-grid contexts handed out by trvb_subgrid_create, keyed by extents;
  // owned by this (root) context so that their cuFFT plans persist.
  std::map<std::vector<int>, trvb_ctx*> subgrids;
  // Per-axis tables on the PARENT grid, indexed by storage index i < n[ax]:
  //   sinc[ax][i]  = sin(u)/u, u = pi*m/n  (m signed)   S/field.cpp:1138-1145
  //   alias[ax][i] = per-axis factor of C1(k)           S/field.cpp:3479-3502
  double* d_sinc[3] = {nullptr, nullptr, nullptr};
  double* d_alias[3] = {nullptr, nullptr, nullptr};
  double* d_ralias[3] = {nullptr, nullptr, nullptr};   // 1 / alias
  // cuFFT plans, created lazily.
  cufftHandle plan_z2z = 0, plan_d2z = 0, plan_z2d = 0;
  bool has_z2z = false, has_d2z = false, has_z2d = false;
  // Batched out-of-place/in-place plans keyed by (cufftType, batch).
  std::map<std::pair<int, int>, cufftHandle> batch_plans;
  // Batched 1-D plans of the slab transforms keyed by {type, length, batch}.
  std::map<std::vector<long long>, cufftHandle> line_plans;
  std::map<std::vector<long long>, size_t> line_plan_work;   // their work-area sizes
  // Spherical-Bessel spline tables keyed by ell.
  std::map<int, SjlTable> sjl;
  // Scratch for two-stage reductions.
  double* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  int num_sms = 148;
  int deterministic = 0;   // 1: reductions avoid atomics (bit-reproducible)
  bool borrowed_stream = false;   // root context that runs on another context's stream
  // Low-|k| storage registered by the distributed mesh phase (root contexts): a HALF mesh at
  // this address holds the modes of this grid on the extents lowk_dims (KView::p0).
  const void* lowk_ptr = nullptr;
  int lowk_dims[3] = {0, 0, 0};
  struct trvb_dmesh* dmesh = nullptr;   // trvb_dmesh_get: owned by the context
};


struct trvb_cat {
  trvb_ctx* owner = nullptr;   // context that owns the allocations (device)
  long long n = 0;
  double* x = nullptr; double* y = nullptr; double* z = nullptr;
  bool borrowed_xyz = false;   // x, y, z are the caller's device arrays (not freed here)
  double* w = nullptr;         // nullptr -> unit weights
  double* los = nullptr;       // SoA: lx[n], ly[n], lz[n] or nullptr
  double* cw = nullptr;        // custom complex weights (interleaved) or nullptr
  // Cell-sorted permutation cache (valid for the grid it was built for) and
  // the catalogue columns gathered into that order (coalesced kernel reads).
  int* order = nullptr;        // particle ids sorted by sort key
  bool order_valid = false;    // false when the last sort did not need (and skipped) it
  bool chunked = false;        // throughput order built chunk by chunk: no global key offsets
  double4* s4 = nullptr;       // {x, y, z, w} per particle: one 32-byte sector
  double* slos = nullptr; double* scw = nullptr;
  bool scw_valid = false;
  int* cell_start = nullptr;   // deterministic mode: nmesh+1 offsets
  int sort_n[3] = {0, 0, 0};
  double sort_L[3] = {0., 0., 0.};
  int sort_shifted = -1;
  int sort_kind = -1;          // 0 tile-sorted (throughput), 1 cell-sorted
  int sort_order = 0;
  // Owner-sorted copies for the tile-owned assignment (trvb_assign_own.cuh): one record
  // per (particle, task it touches), counting-sorted by (task, first plane touched).
  double4* own_rec = nullptr;  // {s_x, s_y, s_z, w}
  int* own_pk = nullptr;       // relative bases, plane key, TSC branches
  int* own_src = nullptr;      // catalogue index of the copy (only with los / custom weights)
  int* own_offsets = nullptr;  // nkeys + 1
  int* own_irr = nullptr;      // catalogue indices of the irregular particles
  long long own_total = 0;
  int own_nirr = 0;
  int own_n[3] = {0, 0, 0};
  double own_L[3] = {0., 0., 0.};
  int own_shifted = -1;
  int own_order = 0;
};

int trvb_scratch(trvb_ctx* ctx, size_t bytes, double** out);
// Device memory comes from a process-wide caching arena (trvb_ctx.cu): blocks
// are cudaMalloc'ed once, handed back to a size-keyed free list on release and
// reused without touching the driver (meshes of a few GB are allocated and
// released many times per estimator call; the driver's own pool re-maps
// physical memory when block sizes vary, which costs milliseconds per GB).
// Reuse is stream-ordered: a block released on stream s is reusable at once
// by work enqueued later on s; another stream first synchronises s.
cudaError_t trvb_arena_alloc(int device, cudaStream_t stream, void** p, size_t bytes);
cudaError_t trvb_arena_free(int device, cudaStream_t stream, void* p);
size_t trvb_arena_cached_bytes(int device);
void trvb_arena_trim(int device);   // cudaFree every cached (unused) block
void trvb_arena_retire_stream(int device, cudaStream_t stream);   // stream synced + about to die
inline cudaError_t trvb_dev_alloc_raw(trvb_ctx* ctx, void** p, size_t bytes) {
  return trvb_arena_alloc(ctx->device, ctx->stream, p, bytes);
}
inline cudaError_t trvb_dev_free_raw(trvb_ctx* ctx, void* p) {
  return p ? trvb_arena_free(ctx->device, ctx->stream, p) : cudaSuccess;
}

// exp(-2 pi i t / N), t < N, on ctx's device (evaluated in long double on the host; one
// table per (device, N) for the life of the process).  csrc/trvb_xpass.cu.
int trvb_twiddle_table(trvb_ctx* ctx, int N, const double2** out);

// ---------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------

__host__ __device__ inline int signed_index(int i, int n) {
  // S/field.cpp:540-544: i < n/2 ? i : i - n.
  return (i < n / 2) ? i : i - n;
}

struct cplx {
  double re, im;
};

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  cplx r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}

__device__ __forceinline__ cplx cconj(cplx a) {
  cplx r; r.re = a.re; r.im = -a.im; return r;
}

// |v| evaluated exactly as std::sqrt(v0*v0 + v1*v1 + v2*v2) on x86-64
// without FMA contraction (S/maths.cpp:57-59).  Shell membership depends
// on the last bit of this value, hence the explicit _rn intrinsics.
__device__ __forceinline__ double vec3_norm_exact(double a, double b, double c) {
  double s = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)),
                       __dmul_rn(c, c));
  return __dsqrt_rn(s);
}

// Mesh accessor for Fourier-space meshes in COMPLEX or HALF layout.
struct KView {
  const double2* p;
  int layout;
  int n0, n1, n2, nh;   // extents of the STORED array
  double add0;   // trvb_mesh::k0_add
  // Low-|k| storage (multi-GPU runs with a distributed mesh): the array holds, in HALF
  // layout on the extents above, only the modes of a (p0, p1, p2) grid that lie strictly
  // inside the stored extents' Nyquist frequencies; callers keep indexing by the big grid.
  int p0 = 0, p1 = 0, p2 = 0;   // 0: the stored array IS the grid
};

// Kernel-side view of a Fourier-space mesh of `ctx`'s grid.
inline KView kview_of(const trvb_ctx* ctx, trvb_mesh m) {
  KView v;
  v.p = (const double2*)m.data; v.layout = m.layout;
  v.n0 = ctx->g.n[0]; v.n1 = ctx->g.n[1]; v.n2 = ctx->g.n[2]; v.nh = ctx->g.nh;
  v.add0 = m.k0_add;
  if (m.data != nullptr && m.data == ctx->lowk_ptr && m.layout == TRVB_HALF) {
    v.p0 = v.n0; v.p1 = v.n1; v.p2 = v.n2;
    v.n0 = ctx->lowk_dims[0]; v.n1 = ctx->lowk_dims[1]; v.n2 = ctx->lowk_dims[2];
    v.nh = v.n2 / 2 + 1;
  }
  return v;
}

// Configuration-space mesh read as complex values (REAL meshes have Im = 0).
struct XView {
  const double* p;
  int cplx;
};
__device__ __forceinline__ double2 xload(const XView& v, long long cell) {
  if (v.cplx) return reinterpret_cast<const double2*>(v.p)[cell];
  return make_double2(v.p[cell], 0.);
}

__device__ __forceinline__ cplx kload(const KView& v, int i, int j, int k) {
  cplx r;
  if ((i | j | k) == 0) {
    double2 t = v.p[0];
    r.re = t.x + v.add0; r.im = t.y;
    return r;
  }
  if (v.p0) {
    // (i, j, k) index the big grid; the array holds its low-|k| modes only.
    int mi = (i < v.p0 / 2) ? i : i - v.p0;
    int mj = (j < v.p1 / 2) ? j : j - v.p1;
    int mk = (k < v.p2 / 2) ? k : k - v.p2;
    r.re = 0.; r.im = 0.;
    if (2 * abs(mi) >= v.n0 || 2 * abs(mj) >= v.n1 || 2 * abs(mk) >= v.n2) return r;
    const bool conj = mk < 0;
    if (conj) { mi = -mi; mj = -mj; mk = -mk; }
    const int is = mi >= 0 ? mi : mi + v.n0, js = mj >= 0 ? mj : mj + v.n1;
    double2 t = v.p[((long long)is * v.n1 + js) * v.nh + mk];
    r.re = t.x; r.im = conj ? -t.y : t.y;
    return r;
  }
  if (v.layout == TRVB_COMPLEX) {
    double2 t = v.p[((long long)i * v.n1 + j) * v.n2 + k];
    r.re = t.x; r.im = t.y;
  } else {
    if (k < v.nh) {
      double2 t = v.p[((long long)i * v.n1 + j) * v.nh + k];
      r.re = t.x; r.im = t.y;
    } else {
      int ic = i ? v.n0 - i : 0, jc = j ? v.n1 - j : 0, kc = v.n2 - k;
      double2 t = v.p[((long long)ic * v.n1 + jc) * v.nh + kc];
      r.re = t.x; r.im = -t.y;
    }
  }
  return r;
}

// Reduced spherical harmonic exactly as S/maths.cpp:171-220:
//   y_lm = sqrt(4pi/(2l+1)) (-1)^((m-|m|)/2) conj( e^{i m phi} Nlm P_l^|m|(mu) )
// with Nlm P_l^|m| the GSL sphPlm (Condon-Shortley phase included),
// y = 1 for l = m = 0, y = 0 for |r| < 1e-9, phi = acos(x/r_xy)
// (2pi - phi if y < 0; 0 if r_xy < 1e-9).
__device__ inline cplx ylm_reduced(int ell, int m, double x, double y, double z) {
  cplx out;
  if (ell == 0 && m == 0) { out.re = 1.; out.im = 0.; return out; }
  const double eps = 1.e-9;
  const double PI = 3.14159265358979323846;
  double r2 = x * x + y * y + z * z;
  double r = sqrt(r2);
  if (fabs(r) < eps) { out.re = 0.; out.im = 0.; return out; }
  double mu = z / r;
  double rxy = sqrt(x * x + y * y);
  double phi = 0.;
  if (fabs(rxy) >= eps) {
    double c = x / rxy;
    c = fmin(1., fmax(-1., c));
    phi = acos(c);
    if (y < 0.) phi = -phi + 2. * PI;
  }
  const int am = (m < 0) ? -m : m;
  // Normalised associated Legendre function sqrt((2l+1)/(4pi) (l-m)!/(l+m)!) P_l^m.
  double somx2 = sqrt((1. - mu) * (1. + mu));
  double pmm = sqrt(1. / (4. * PI));
  for (int i = 1; i <= am; i++) {
    pmm *= -sqrt((2. * i + 1.) / (2. * i)) * somx2;
  }
  double plm;
  if (ell == am) {
    plm = pmm;
  } else {
    double pmmp1 = mu * sqrt(2. * am + 3.) * pmm;
    if (ell == am + 1) {
      plm = pmmp1;
    } else {
      double pll = 0.;
      for (int ll = am + 2; ll <= ell; ll++) {
        double a = sqrt((4. * ll * ll - 1.) / ((double)ll * ll - (double)am * am));
        double b = sqrt((((double)ll - 1.) * (ll - 1.) - (double)am * am)
                        / (4. * (ll - 1.) * (ll - 1.) - 1.));
        pll = a * (mu * pmmp1 - b * pmm);
        pmm = pmmp1; pmmp1 = pll;
      }
      plm = pll;
    }
  }
  double sn, cs;
  sincos((double)m * phi, &sn, &cs);
  // conj(e^{i m phi} P) = (cos, -sin) P; parity (-1)^((m-|m|)/2).
  double par = (((m - am) / 2) % 2 != 0) ? -1. : 1.;
  double norm = sqrt(4. * PI / (2. * ell + 1.)) * par * plm;
  out.re = norm * cs;
  out.im = -norm * sn;
  return out;
}

// The same function with everything that depends on (l, m) only hoisted into
// a coefficient set: per evaluation 3 square roots and 2 divisions remain, and
// e^{i m phi} comes from (x + i y)/r_xy by repeated multiplication instead of
// acos + sincos (identical in exact arithmetic: phi = acos(x/r_xy), reflected
// for y < 0, means cos phi = x/r_xy and sin phi = y/r_xy).  Agreement with
// ylm_reduced is at the 1e-15 level.
constexpr int YLM_NREC = 8;   // recursion steps held in registers: l - |m| - 1 <= 8
struct YlmCoef {
  int ell, m, am, nrec;
  bool generic;               // l beyond the register budget: use ylm_reduced
  double pmm0;                // sqrt(1/4pi) prod_i -sqrt((2i+1)/(2i)), i = 1..|m|
  double c1;                  // sqrt(2|m| + 3)
  double a[YLM_NREC], b[YLM_NREC];
  double norm;                // sqrt(4pi/(2l+1)) (-1)^((m-|m|)/2)
};

__host__ __device__ inline YlmCoef ylm_coef(int ell, int m) {
  YlmCoef y;
  const double PI = 3.14159265358979323846;
  y.ell = ell; y.m = m; y.am = (m < 0) ? -m : m;
  y.nrec = ell - y.am - 1; if (y.nrec < 0) y.nrec = 0;
  y.generic = y.nrec > YLM_NREC;
  y.pmm0 = sqrt(1. / (4. * PI));
  for (int i = 1; i <= y.am; i++) y.pmm0 *= -sqrt((2. * i + 1.) / (2. * i));
  y.c1 = sqrt(2. * y.am + 3.);
  for (int t = 0; t < YLM_NREC; t++) {
    const int ll = y.am + 2 + t;
    y.a[t] = sqrt((4. * ll * ll - 1.) / ((double)ll * ll - (double)y.am * y.am));
    y.b[t] = sqrt((((double)ll - 1.) * (ll - 1.) - (double)y.am * y.am)
                  / (4. * (ll - 1.) * (ll - 1.) - 1.));
  }
  const double par = (((m - y.am) / 2) % 2 != 0) ? -1. : 1.;
  y.norm = sqrt(4. * PI / (2. * ell + 1.)) * par;
  return y;
}

__device__ __forceinline__ cplx ylm_eval(const YlmCoef& c, double x, double y, double z) {
  cplx out;
  if (c.ell == 0) { out.re = 1.; out.im = 0.; return out; }
  if (c.generic) return ylm_reduced(c.ell, c.m, x, y, z);
  const double eps = 1.e-9;
  const double rxy2 = x * x + y * y;
  const double r = sqrt(rxy2 + z * z);
  if (r < eps) { out.re = 0.; out.im = 0.; return out; }
  const double mu = z / r;
  const double rxy = sqrt(rxy2);
  double cs = 1., sn = 0.;   // phi = 0 when r_xy < eps
  if (rxy >= eps) { const double ir = 1. / rxy; cs = x * ir; sn = y * ir; }
  // (cs + i sn)^|m|
  double er = 1., ei = 0.;
  for (int i = 0; i < c.am; i++) { const double t = er * cs - ei * sn; ei = er * sn + ei * cs; er = t; }
  const double somx2 = sqrt((1. - mu) * (1. + mu));
  double pmm = c.pmm0;
  for (int i = 0; i < c.am; i++) pmm *= somx2;
  double plm = pmm;
  if (c.ell > c.am) {
    double pmmp1 = mu * c.c1 * pmm;
    plm = pmmp1;
#pragma unroll
    for (int t = 0; t < YLM_NREC; t++) {
      if (t < c.nrec) {
        const double pll = c.a[t] * (mu * pmmp1 - c.b[t] * pmm);
        pmm = pmmp1; pmmp1 = pll; plm = pll;
      }
    }
  }
  // conj(e^{i m phi}) P: e^{i m phi} = (er, sign(m) ei).
  const double v = c.norm * plm;
  out.re = v * er;
  out.im = (c.m < 0) ? v * ei : -v * ei;
  return out;
}

// Spherical Bessel j_l(x) evaluated directly (x >= split region, or any x for
// l = 0): upward recurrence, stable for x >= l (S/maths.cpp:368-371 uses
// gsl_sf_bessel_jl there).
__device__ inline double sjl_direct(int ell, double x) {
  if (x == 0.) return (ell == 0) ? 1. : 0.;
  double s, c;
  sincos(x, &s, &c);
  double j0 = s / x;
  if (ell == 0) return j0;
  double j1 = (s / x - c) / x;
  if (ell == 1) return j1;
  double jm = j0, jc = j1;
  for (int n = 1; n < ell; n++) {
    double jn = (2. * n + 1.) / x * jc - jm;
    jm = jc; jc = jn;
  }
  return jc;
}

struct SjlView {
  const double* y;
  const double* c;
  int nsample;
  double step;
  int ell;
};

// SphericalBesselCalculator::eval (S/maths.cpp:368-375): natural cubic spline
// (GSL cspline evaluation formula) below `split`, direct evaluation above.
__device__ __forceinline__ double sjl_eval(const SjlView& t, double x) {
  const double split = t.step * (t.nsample - 1);
  if (x >= split) return sjl_direct(t.ell, x);
  int i = (int)(x / t.step);
  if (i > t.nsample - 2) i = t.nsample - 2;
  if (i < 0) i = 0;
  // Match the bisection result x_i <= x < x_{i+1} with x_i = step * i.
  while (i < t.nsample - 2 && t.step * (i + 1) <= x) i++;
  while (i > 0 && t.step * i > x) i--;
  const double x_lo = t.step * i, x_hi = t.step * (i + 1);
  const double y_lo = __ldg(t.y + i), y_hi = __ldg(t.y + i + 1);
  const double c_i = __ldg(t.c + i), c_ip1 = __ldg(t.c + i + 1);
  const double dx = x_hi - x_lo;
  const double dy = y_hi - y_lo;
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  const double delx = x - x_lo;
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

// y_lm at the mirror image (+-x, +-y, +-z) of a point from its value y0 there, by the
// parity rules of the factors of ylm_eval: P_l^|m|(-mu) = (-1)^(l-|m|) P_l^|m|(mu),
// (-cs + i sn)^|m| = (-1)^|m| conj((cs + i sn)^|m|), (cs - i sn)^|m| = conj(...).
// Sign flips commute with every rounding in ylm_eval, so the result is bit-identical to
// evaluating at the mirrored point.  Not valid for the generic (large l) evaluator.
__device__ __forceinline__ cplx ylm_mirror(const YlmCoef& c, const cplx& y0, bool fx, bool fy,
                                           bool fz) {
  const double par_am = (c.am & 1) ? -1. : 1.;
  const double par_l = ((c.ell - c.am) & 1) ? -1. : 1.;
  const double sz = fz ? par_l : 1.;
  cplx out;
  out.re = y0.re * ((fx ? par_am : 1.) * sz);
  out.im = y0.im * ((fx ? -par_am : 1.) * (fy ? -1. : 1.) * sz);
  return out;
}

// The same interpolant with the three divisions of the evaluation formula replaced
// by multiplications with precomputed reciprocals (1/step, 1/3): equal to sjl_eval
// to a few ulp, for passes that evaluate it once per mesh cell and bin.
__device__ __forceinline__ double sjl_eval_fast(const SjlView& t, double x, double inv_step) {
  const double split = t.step * (t.nsample - 1);
  if (x >= split) return sjl_direct(t.ell, x);
  int i = (int)(x * inv_step);
  if (i > t.nsample - 2) i = t.nsample - 2;
  if (i < 0) i = 0;
  while (i < t.nsample - 2 && t.step * (i + 1) <= x) i++;
  while (i > 0 && t.step * i > x) i--;
  const double x_lo = t.step * i, x_hi = t.step * (i + 1);
  const double y_lo = __ldg(t.y + i), y_hi = __ldg(t.y + i + 1);
  const double c_i = __ldg(t.c + i), c_ip1 = __ldg(t.c + i + 1);
  const double dx = x_hi - x_lo;
  const double third = 1. / 3.;
  const double b_i = (y_hi - y_lo) * inv_step - dx * (c_ip1 + 2.0 * c_i) * third;
  const double d_i = (c_ip1 - c_i) * inv_step * third;
  const double delx = x - x_lo;
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

// Block-wide sum of one double; result valid in thread 0.  blockDim.x must
// be a multiple of 32 and <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  double r = 0.;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? smem32[lane] : 0.;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  return r;
}

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Row-wise traversal of an [n0][n1][n2s] mesh without 64-bit divisions:
// rows (i, j) are dealt to (blockIdx.x, threadIdx.y), the contiguous last
// axis to threadIdx.x.  f(i, j, k, flat_index).
template <class F>
__device__ __forceinline__ void for_each_cell(int n0, int n1, int n2s, F f) {
  const int rows = n0 * n1;
  for (int row = blockIdx.x * blockDim.y + threadIdx.y; row < rows;
       row += gridDim.x * blockDim.y) {
    const int i = row / n1, j = row - i * n1;
    const long long base = (long long)row * n2s;
    for (int k = threadIdx.x; k < n2s; k += blockDim.x) f(i, j, k, base + k);
  }
}

struct RowLaunch { dim3 grid, block; };
inline RowLaunch row_launch(int num_sms, int n0, int n1, int n2s) {
  int bx = ((n2s < 256 ? n2s : 256) + 31) / 32 * 32;
  int by = 256 / bx; if (by < 1) by = 1;
  const long long rows = (long long)n0 * n1;
  long long gx = (rows + by - 1) / by;
  const long long cap = (long long)num_sms * 32;
  if (gx > cap) gx = cap;
  RowLaunch r; r.grid = dim3((unsigned)gx, 1, 1); r.block = dim3(bx, by, 1);
  return r;
}

#endif  // TRVB_COMMON_CUH_
