// trvb_fourier.cu -- transforms and Fourier-space field construction.
//
// cuFFT is the only library call (forward/inverse 3-D FFT, Z2Z / D2Z / Z2D);
// everything around it -- scaling, mean subtraction, interlacing, window
// compensation, shell filtering with y_lm weights, spherical-Bessel weighting
// -- is a hand-written vectorised kernel adjacent to the transform.
// Replaces S/field.cpp:1496-1720 (transforms), 1764-1785 (compensation),
// 1792-1906 (band-limited y_lm-weighted IFFT), 1908-2010 (j_l-weighted IFFT).
#include "trvb_common.cuh"
#include "trvb_zpass.cuh"

#include <algorithm>
#include <map>
#include <mutex>

namespace {

__global__ void k_scale(double* __restrict__ p, long long n, double s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) p[i] *= s;
}

__global__ void k_add_const(double* __restrict__ p, long long n, int stride, double c) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) p[i * stride] += c;
}

__global__ void k_axpby(double* __restrict__ dst, const double* __restrict__ src,
                        long long n, double a, double b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) dst[i] = a * dst[i] + b * src[i];
}

__global__ void k_add_zero_mode(double* p, double add_re) { p[0] += add_re; }

// Number of stored complex elements along the last axis.
__host__ __device__ inline int kdim2(const GridDesc& g, int layout) {
  return layout == TRVB_HALF ? g.nh : g.n[2];
}

// f = (f + e^{+i pi (mx+my+mz)} f_s) / 2 with m = i/n or i/n - 1
// (S/field.cpp:1618-1653).
__global__ void __launch_bounds__(256)
k_interlace(double2* __restrict__ f, const double2* __restrict__ fs,
            GridDesc g, int layout) {
  for_each_cell(g.n[0], g.n[1], kdim2(g, layout), [&](int i, int j, int k, long long t) {
    double m0 = (i < g.n[0] / 2) ? double(i) / g.n[0] : double(i) / g.n[0] - 1;
    double m1 = (j < g.n[1] / 2) ? double(j) / g.n[1] : double(j) / g.n[1] - 1;
    double m2 = (k < g.n[2] / 2) ? double(k) / g.n[2] : double(k) / g.n[2] - 1;
    double arg = 3.14159265358979323846 * (m0 + m1 + m2);
    double sn, cs;
    sincos(arg, &sn, &cs);
    double2 a = f[t], b = fs[t];
    a.x += cs * b.x - sn * b.y;
    a.y += sn * b.x + cs * b.y;
    a.x /= 2.; a.y /= 2.;
    f[t] = a;
  });
}

struct Tables {
  const double* sinc[3];
  const double* alias[3];
};

__device__ __forceinline__ double window_at(const Tables& t, int order, int i, int j, int k) {
  // pow(wk_x * wk_y * wk_z, order), S/field.cpp:1147-1149.
  double wk = t.sinc[0][i] * t.sinc[1][j] * t.sinc[2][k];
  double w = wk;
  for (int q = 1; q < order; q++) w *= wk;
  return w;
}

__global__ void __launch_bounds__(256)
k_compensate(double2* __restrict__ f, GridDesc g, int layout, Tables tb) {
  for_each_cell(g.n[0], g.n[1], kdim2(g, layout), [&](int i, int j, int k, long long t) {
    const double w = window_at(tb, g.order, i, j, k);
    double2 a = f[t];
    a.x /= w; a.y /= w;
    f[t] = a;
  });
}

// f(x) *= |x|^(-power) for |x| >= 1e-6, x the signed cell offset vector
// (S/field.cpp:1727-1762, 546-553); `stride` = doubles per cell.
__global__ void __launch_bounds__(256)
k_pow_law(double* __restrict__ f, GridDesc g, int stride, int power) {
  for_each_cell(g.n[0], g.n[1], g.n[2], [&](int i, int j, int k, long long t) {
    const double rx = __dmul_rn((double)signed_index(i, g.n[0]), g.dr[0]);
    const double ry = __dmul_rn((double)signed_index(j, g.n[1]), g.dr[1]);
    const double rz = __dmul_rn((double)signed_index(k, g.n[2]), g.dr[2]);
    const double r = vec3_norm_exact(rx, ry, rz);
    if (r < 1.e-6) return;
    const double w = pow(r, (double)(-power));
    for (int c = 0; c < stride; c++) f[t * stride + c] *= w;
  });
}

// One Fourier mode of the filtered spectrum (S/field.cpp:1815-1847):
// y_lm(khat) src(k) / W(k) * amp for the signed mode (mi, mj, mk).
__device__ __forceinline__ double2 shell_mode(const KView& src, const GridDesc& gp,
                                              const Tables& tb, const YlmCoef& yc,
                                              int mi, int mj, int mk, double kx, double ky,
                                              double kz, double amp) {
  const int ip = mi >= 0 ? mi : mi + gp.n[0];
  const int jp = mj >= 0 ? mj : mj + gp.n[1];
  const int kp = mk >= 0 ? mk : mk + gp.n[2];
  cplx fk = kload(src, ip, jp, kp);
  const double rw = 1. / window_at(tb, gp.order, ip, jp, kp);
  fk.re *= rw; fk.im *= rw;
  const cplx y = ylm_eval(yc, kx, ky, kz);
  const cplx v = cmul(y, fk);
  return make_double2(v.re * amp, v.im * amp);
}

// Dense form: every cell of the destination (sub-)grid spectrum is written.
//   gs = sub grid, gp = parent grid (tables and `src` live on the parent);
//   n2s = stored extent of dst's last axis: gs.n[2] (COMPLEX) or gs.nh (HALF).
__global__ void __launch_bounds__(256)
k_shell_spectrum(KView src, GridDesc gp, GridDesc gs, Tables tb, int ell, int m,
                 double klo, double khi, int use_shell, double amp, int n2s,
                 double2* __restrict__ dst) {
  const bool same = gs.n[0] == gp.n[0] && gs.n[1] == gp.n[1] && gs.n[2] == gp.n[2];
  const YlmCoef yc = ylm_coef(ell, m);
  for_each_cell(gs.n[0], gs.n[1], n2s, [&](int is, int js, int ks, long long t) {
    const int mi = signed_index(is, gs.n[0]), mj = signed_index(js, gs.n[1]);
    const int mk = signed_index(ks, gs.n[2]);
    // On a coarser grid only modes strictly inside its Nyquist frequency are
    // representable on both grids without ambiguity.
    const bool rep = same
      || ((2 * abs(mi) < gs.n[0]) && (2 * abs(mj) < gs.n[1]) && (2 * abs(mk) < gs.n[2]));
    double2 out = make_double2(0., 0.);
    if (rep) {
      // kv = i * dk (S/field.cpp:555-562), |k| without contraction.
      const double kx = __dmul_rn((double)mi, gp.dk[0]);
      const double ky = __dmul_rn((double)mj, gp.dk[1]);
      const double kz = __dmul_rn((double)mk, gp.dk[2]);
      const double kmag = vec3_norm_exact(kx, ky, kz);
      if (!use_shell || (klo <= kmag && kmag < khi)) {
        out = shell_mode(src, gp, tb, yc, mi, mj, mk, kx, ky, kz, amp);
      }
    }
    dst[t] = out;
  });
}

// Sparse batched form: the destination slab (nbins consecutive spectra on the
// sub grid) has been zero-filled; only the modes of the low-|k| cube are
// visited and each is written into the spectrum of every bin that holds it.
struct ShellBatch {
  const double* klo; const double* khi; const double* amp;   // device, [nbins]
  int nbins;
};

__global__ void __launch_bounds__(256)
k_shell_scatter(KView src, GridDesc gp, GridDesc gs, Tables tb, int ell, int m,
                int lo0, int lo1, int lo2, int c0, int c1, int c2,
                ShellBatch sb, int n2s, long long bin_stride, double2* __restrict__ dst) {
  const YlmCoef yc = ylm_coef(ell, m);
  for_each_cell(c0, c1, c2, [&](int a, int b, int c, long long) {
    const int mi = lo0 + a, mj = lo1 + b, mk = lo2 + c;
    // HALF destination: mk >= 0 stored, plus the Nyquist plane (signed -n/2).
    if (n2s != gs.n[2] && mk < 0 && 2 * (-mk) != gs.n[2]) return;
    const double kx = __dmul_rn((double)mi, gp.dk[0]);
    const double ky = __dmul_rn((double)mj, gp.dk[1]);
    const double kz = __dmul_rn((double)mk, gp.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    bool loaded = false;
    double2 base = make_double2(0., 0.);
    for (int q = 0; q < sb.nbins; q++) {
      if (!(sb.klo[q] <= kmag && kmag < sb.khi[q])) continue;
      if (!loaded) {
        base = shell_mode(src, gp, tb, yc, mi, mj, mk, kx, ky, kz, 1.);
        loaded = true;
      }
      const int is = mi >= 0 ? mi : mi + gs.n[0];
      const int js = mj >= 0 ? mj : mj + gs.n[1];
      const int ks = mk >= 0 ? mk : mk + gs.n[2];
      const double amp = sb.amp[q];
      dst[q * bin_stride + ((long long)is * gs.n[1] + js) * n2s + ks] =
        make_double2(base.x * amp, base.y * amp);
    }
  });
}

// j_l(|k| r) y_lm(khat) src(k)/W(k) * amp on the full grid (S/field.cpp:1936-1961).
__global__ void __launch_bounds__(256)
k_sjl_spectrum(KView src, GridDesc g, Tables tb, SjlView sj, int ell, int m,
               double r, double amp, double2* __restrict__ dst) {
  const YlmCoef yc = ylm_coef(ell, m);
  for_each_cell(g.n[0], g.n[1], g.n[2], [&](int i, int j, int k, long long t) {
    const double kx = __dmul_rn((double)signed_index(i, g.n[0]), g.dk[0]);
    const double ky = __dmul_rn((double)signed_index(j, g.n[1]), g.dk[1]);
    const double kz = __dmul_rn((double)signed_index(k, g.n[2]), g.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    cplx fk = kload(src, i, j, k);
    const double rw = 1. / window_at(tb, g.order, i, j, k);
    fk.re *= rw; fk.im *= rw;
    cplx y = ylm_eval(yc, kx, ky, kz);
    cplx v = cmul(y, fk);
    const double jl = sjl_eval(sj, __dmul_rn(kmag, r));
    dst[t] = make_double2(jl * v.re * amp, jl * v.im * amp);
  });
}

// The bin-independent part of k_sjl_spectrum, evaluated once per (l, m):
// A(k) = amp y_lm(khat) src(k) / W(k) and |k|.
__global__ void __launch_bounds__(256)
k_sjl_prepare(KView src, GridDesc g, Tables tb, int ell, int m, double amp,
              double2* __restrict__ A, double* __restrict__ kmag_out) {
  const YlmCoef yc = ylm_coef(ell, m);
  for_each_cell(g.n[0], g.n[1], g.n[2], [&](int i, int j, int k, long long t) {
    const double kx = __dmul_rn((double)signed_index(i, g.n[0]), g.dk[0]);
    const double ky = __dmul_rn((double)signed_index(j, g.n[1]), g.dk[1]);
    const double kz = __dmul_rn((double)signed_index(k, g.n[2]), g.dk[2]);
    cplx fk = kload(src, i, j, k);
    const double rw = 1. / window_at(tb, g.order, i, j, k);
    fk.re *= rw; fk.im *= rw;
    const cplx v = cmul(ylm_eval(yc, kx, ky, kz), fk);
    A[t] = make_double2(v.re * amp, v.im * amp);
    kmag_out[t] = vec3_norm_exact(kx, ky, kz);
  });
}

// NB spectra j_l(|k| r_q) A(k) per pass over A and |k| (24 B read, 16 B written per
// bin and mode; no division, no square root).
constexpr int SJL_NB = 4;
struct SjlBatch { double r[SJL_NB]; double2* dst[SJL_NB]; int count; };

__global__ void __launch_bounds__(256)
k_sjl_apply(const double2* __restrict__ A, const double* __restrict__ kmag, long long nmesh,
            SjlView sj, SjlBatch b) {
  const double inv_step = 1. / sj.step;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const double2 a = A[t];
    const double kk = kmag[t];
#pragma unroll
    for (int q = 0; q < SJL_NB; q++) {
      if (q < b.count) {
        const double jl = sjl_eval_fast(sj, __dmul_rn(kk, b.r[q]), inv_step);
        b.dst[q][t] = make_double2(jl * a.x, jl * a.y);
      }
    }
  }
}

// Plans with cuFFT's own work areas allocate behind the arena's back: when that fails the
// arena's cached blocks go back to the driver and the plan is tried once more (the arena
// does the same for its own allocations).
template <typename Make>
cufftResult plan_with_retry(trvb_ctx* ctx, Make make) {
  cufftResult rc = make();
  if (rc == CUFFT_ALLOC_FAILED || rc == CUFFT_INTERNAL_ERROR) {
    cudaGetLastError();
    // batched plans of other batch sizes keep their work areas: they are rebuilt on demand
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->batch_plans) cufftDestroy(kv.second);
    ctx->batch_plans.clear();
    trvb_arena_trim(ctx->device);
    rc = make();
  }
  return rc;
}

// Batched plan over `batch` consecutive meshes of ctx's grid.
int get_batch_plan(trvb_ctx* ctx, cufftType type, int batch, cufftHandle* out) {
  auto key = std::make_pair((int)type, batch);
  auto it = ctx->batch_plans.find(key);
  if (it == ctx->batch_plans.end()) {
    const GridDesc& g = ctx->g;
    int dims[3] = {g.n[0], g.n[1], g.n[2]};
    const long long nfull = g.nmesh, nhalf = (long long)g.n[0] * g.n[1] * g.nh;
    TRVB_REQUIRE(nfull < 2147483647LL, "batched FFT: grid too large for a batched plan");
    cufftHandle plan;
    if (type == CUFFT_Z2Z) {
      TRVB_CUFFT(plan_with_retry(ctx, [&] {
        return cufftPlanMany(&plan, 3, dims, nullptr, 1, (int)nfull, nullptr, 1, (int)nfull,
                             CUFFT_Z2Z, batch); }));
    } else {   // Z2D, out of place: HALF spectra -> REAL meshes
      TRVB_CUFFT(plan_with_retry(ctx, [&] {
        return cufftPlanMany(&plan, 3, dims, nullptr, 1, (int)nhalf, nullptr, 1, (int)nfull,
                             CUFFT_Z2D, batch); }));
    }
    TRVB_CUFFT(cufftSetStream(plan, ctx->stream));
    it = ctx->batch_plans.emplace(key, plan).first;
  }
  *out = it->second;
  return 0;
}

int get_plan(trvb_ctx* ctx, cufftType type, cufftHandle* out) {
  cufftHandle* slot; bool* has;
  if (type == CUFFT_Z2Z) { slot = &ctx->plan_z2z; has = &ctx->has_z2z; }
  else if (type == CUFFT_D2Z) { slot = &ctx->plan_d2z; has = &ctx->has_d2z; }
  else { slot = &ctx->plan_z2d; has = &ctx->has_z2d; }
  if (!*has) {
    TRVB_CUFFT(plan_with_retry(ctx, [&] {
      return cufftPlan3d(slot, ctx->g.n[0], ctx->g.n[1], ctx->g.n[2], type); }));
    TRVB_CUFFT(cufftSetStream(*slot, ctx->stream));
    *has = true;
  }
  *out = *slot;
  return 0;
}

Tables tables_of(const trvb_ctx* ctx) {
  const trvb_ctx* root = ctx->parent ? ctx->parent : ctx;
  Tables t;
  for (int a = 0; a < 3; a++) { t.sinc[a] = root->d_sinc[a]; t.alias[a] = root->d_alias[a]; }
  return t;
}


inline int grid_for(const trvb_ctx* ctx, long long n, int threads) {
  return (int)std::min<long long>(div_up(n, threads), (long long)ctx->num_sms * 32);
}

}  // namespace

extern "C" int trvb_mesh_add_const(trvb_ctx* ctx, trvb_mesh mesh, double c) {
  TRVB_REQUIRE(ctx && mesh.data, "trvb_mesh_add_const: null argument");
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_mesh_add_const: configuration-space layouts only");
  const int stride = mesh.layout == TRVB_COMPLEX ? 2 : 1;
  k_add_const<<<grid_for(ctx, ctx->g.nmesh, 256), 256, 0, ctx->stream>>>(
    (double*)mesh.data, ctx->g.nmesh, stride, c);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_mesh_axpby(trvb_ctx* ctx, trvb_mesh dst, double a, trvb_mesh src,
                               double b) {
  TRVB_REQUIRE(ctx && dst.data && src.data, "trvb_mesh_axpby: null argument");
  TRVB_REQUIRE(dst.layout == src.layout, "trvb_mesh_axpby: layouts differ");
  const long long n = (long long)(trvb_mesh_bytes(ctx, dst.layout) / sizeof(double));
  k_axpby<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(
    (double*)dst.data, (const double*)src.data, n, a, b);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_mesh_pow_law(trvb_ctx* ctx, trvb_mesh mesh, int power) {
  TRVB_REQUIRE(ctx && mesh.data, "trvb_mesh_pow_law: null argument");
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_mesh_pow_law: configuration-space layouts only");
  const int stride = mesh.layout == TRVB_COMPLEX ? 2 : 1;
  const RowLaunch rl = row_launch(ctx->num_sms, ctx->g.n[0], ctx->g.n[1], ctx->g.n[2]);
  k_pow_law<<<rl.grid, rl.block, 0, ctx->stream>>>((double*)mesh.data, ctx->g, stride, power);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_fft_forward(trvb_ctx* ctx, trvb_mesh src, trvb_mesh dst,
                                double prescale) {
  TRVB_REQUIRE(ctx && src.data && dst.data, "trvb_fft_forward: null argument");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  if (prescale != 1.) {
    const long long n = (long long)(trvb_mesh_bytes(ctx, src.layout) / sizeof(double));
    k_scale<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>((double*)src.data, n, prescale);
    TRVB_LAUNCH_CHECK();
  }
  cufftHandle plan;
  if (src.layout == TRVB_REAL) {
    TRVB_REQUIRE(dst.layout == TRVB_HALF && dst.data != src.data,
                 "trvb_fft_forward: REAL source needs a distinct HALF destination");
    int st = get_plan(ctx, CUFFT_D2Z, &plan); if (st) return st;
    TRVB_CUFFT(cufftExecD2Z(plan, (cufftDoubleReal*)src.data, (cufftDoubleComplex*)dst.data));
  } else {
    TRVB_REQUIRE(src.layout == TRVB_COMPLEX && dst.layout == TRVB_COMPLEX,
                 "trvb_fft_forward: COMPLEX source needs a COMPLEX destination");
    int st = get_plan(ctx, CUFFT_Z2Z, &plan); if (st) return st;
    TRVB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)src.data,
                            (cufftDoubleComplex*)dst.data, CUFFT_FORWARD));
  }
  g_trvb_fft_execs++;
  return 0;
}

extern "C" int trvb_fft_inverse(trvb_ctx* ctx, trvb_mesh src, trvb_mesh dst) {
  TRVB_REQUIRE(ctx && src.data && dst.data, "trvb_fft_inverse: null argument");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  cufftHandle plan;
  if (src.layout == TRVB_HALF) {
    TRVB_REQUIRE(dst.layout == TRVB_REAL && dst.data != src.data,
                 "trvb_fft_inverse: HALF source needs a distinct REAL destination");
    int st = get_plan(ctx, CUFFT_Z2D, &plan); if (st) return st;
    TRVB_CUFFT(cufftExecZ2D(plan, (cufftDoubleComplex*)src.data, (cufftDoubleReal*)dst.data));
  } else {
    TRVB_REQUIRE(src.layout == TRVB_COMPLEX && dst.layout == TRVB_COMPLEX,
                 "trvb_fft_inverse: COMPLEX source needs a COMPLEX destination");
    int st = get_plan(ctx, CUFFT_Z2Z, &plan); if (st) return st;
    TRVB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)src.data,
                            (cufftDoubleComplex*)dst.data, CUFFT_INVERSE));
  }
  g_trvb_fft_execs++;
  return 0;
}

extern "C" int trvb_kmesh_add_zero_mode(trvb_ctx* ctx, trvb_mesh kmesh, double add_re) {
  TRVB_REQUIRE(ctx && kmesh.data && kmesh.layout != TRVB_REAL,
               "trvb_kmesh_add_zero_mode: needs a Fourier-space mesh");
  k_add_zero_mode<<<1, 1, 0, ctx->stream>>>((double*)kmesh.data, add_re);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_interlace_combine(trvb_ctx* ctx, trvb_mesh kmesh, trvb_mesh kmesh_s) {
  TRVB_REQUIRE(ctx && kmesh.data && kmesh_s.data, "trvb_interlace_combine: null argument");
  TRVB_REQUIRE(kmesh.layout == kmesh_s.layout && kmesh.layout != TRVB_REAL,
               "trvb_interlace_combine: both meshes must share a Fourier layout");
  const RowLaunch rl = row_launch(ctx->num_sms, ctx->g.n[0], ctx->g.n[1], kdim2(ctx->g, kmesh.layout));
  k_interlace<<<rl.grid, rl.block, 0, ctx->stream>>>(
    (double2*)kmesh.data, (const double2*)kmesh_s.data, ctx->g, kmesh.layout);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_compensate(trvb_ctx* ctx, trvb_mesh kmesh) {
  TRVB_REQUIRE(ctx && kmesh.data && kmesh.layout != TRVB_REAL,
               "trvb_compensate: needs a Fourier-space mesh");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_compensate: root context only");
  const RowLaunch rl = row_launch(ctx->num_sms, ctx->g.n[0], ctx->g.n[1], kdim2(ctx->g, kmesh.layout));
  k_compensate<<<rl.grid, rl.block, 0, ctx->stream>>>(
    (double2*)kmesh.data, ctx->g, kmesh.layout, tables_of(ctx));
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_shell_ifft(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell,
                               int m, double klo, double khi, double amp,
                               trvb_mesh dst) {
  TRVB_REQUIRE(ctx && sub && src.data && dst.data, "trvb_shell_ifft: null argument");
  TRVB_REQUIRE(ctx->parent == nullptr && (sub == ctx || sub->parent == ctx),
               "trvb_shell_ifft: `sub` must be `ctx` or a sub-grid of it");
  TRVB_REQUIRE(src.layout != TRVB_REAL, "trvb_shell_ifft: src must be a Fourier mesh");
  // A REAL destination is allowed when the filtered spectrum is Hermitian:
  // spectrum of a real field (HALF) times a real, even-parity y_l0.
  const bool herm = src.layout == TRVB_HALF && m == 0 && (ell % 2) == 0;
  TRVB_REQUIRE(dst.layout == TRVB_COMPLEX || (dst.layout == TRVB_REAL && herm),
               "trvb_shell_ifft: dst must be COMPLEX (REAL only for a HALF source, m = 0, even l)");
  TRVB_REQUIRE(abs(m) <= ell && ell >= 0, "trvb_shell_ifft: bad (l, m) = (%d, %d)", ell, m);
  TRVB_REQUIRE(dst.data != src.data, "trvb_shell_ifft: src and dst must differ");
  const int use_shell = !(klo < 0. && khi < 0.);
  const GridDesc& gs = sub->g;
  if (dst.layout == TRVB_REAL) {
    trvb_mesh half; half.layout = TRVB_HALF; half.k0_add = 0.; half.data = nullptr;
    TRVB_CUDA(trvb_dev_alloc_raw(sub, &half.data, trvb_mesh_bytes(sub, TRVB_HALF)));
    const RowLaunch rl = row_launch(ctx->num_sms, gs.n[0], gs.n[1], gs.nh);
    k_shell_spectrum<<<rl.grid, rl.block, 0, sub->stream>>>(
      kview_of(ctx, src), ctx->g, gs, tables_of(ctx), ell, m, klo, khi, use_shell, amp,
      gs.nh, (double2*)half.data);
    TRVB_LAUNCH_CHECK();
    int st = trvb_fft_inverse(sub, half, dst);
    trvb_dev_free_raw(sub, half.data);
    return st;
  }
  const RowLaunch rl = row_launch(ctx->num_sms, gs.n[0], gs.n[1], gs.n[2]);
  k_shell_spectrum<<<rl.grid, rl.block, 0, sub->stream>>>(
    kview_of(ctx, src), ctx->g, gs, tables_of(ctx), ell, m, klo, khi, use_shell, amp,
    gs.n[2], (double2*)dst.data);
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(sub, dst, dst);
}

extern "C" int trvb_shell_ifft_batch(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell,
                                     int m, const double* klo, const double* khi,
                                     const double* amp, int nbins, void* dst,
                                     int dst_layout) {
  TRVB_REQUIRE(ctx && sub && src.data && dst && klo && khi && amp && nbins > 0,
               "trvb_shell_ifft_batch: bad argument");
  TRVB_REQUIRE(ctx->parent == nullptr && (sub == ctx || sub->parent == ctx),
               "trvb_shell_ifft_batch: `sub` must be `ctx` or a sub-grid of it");
  TRVB_REQUIRE(src.layout != TRVB_REAL, "trvb_shell_ifft_batch: src must be a Fourier mesh");
  const bool herm = src.layout == TRVB_HALF && m == 0 && (ell % 2) == 0;
  TRVB_REQUIRE(dst_layout == TRVB_COMPLEX || (dst_layout == TRVB_REAL && herm),
               "trvb_shell_ifft_batch: dst must be COMPLEX (REAL only for a HALF source, m = 0, even l)");
  TRVB_REQUIRE(abs(m) <= ell && ell >= 0, "trvb_shell_ifft_batch: bad (l, m) = (%d, %d)", ell, m);
  const GridDesc& gp = ctx->g;
  const GridDesc& gs = sub->g;
  const bool real_out = dst_layout == TRVB_REAL;
  const int n2s = real_out ? gs.nh : gs.n[2];
  const long long bin_stride = (long long)gs.n[0] * gs.n[1] * n2s;   // complex elements
  const size_t dst_stride = trvb_mesh_bytes(sub, dst_layout);
  // Sub-batches bound the transient half-spectrum slab (and cuFFT's work
  // area) to a few GiB whatever the number of bins and the grid size.
  const size_t spec_bin_bytes = sizeof(double2) * (size_t)bin_stride;
  int maxb = (int)std::max<size_t>(1, ((size_t)6 << 30) / spec_bin_bytes);
  maxb = std::min(maxb, nbins);
  void* spec_tmp = nullptr;
  if (real_out) TRVB_CUDA(trvb_dev_alloc_raw(sub, &spec_tmp, spec_bin_bytes * maxb));
  // Low-|k| cube that holds every shell; clipped to what the sub grid represents.
  double kmax = 0.;
  for (int q = 0; q < nbins; q++) kmax = std::max(kmax, khi[q]);
  int lo[3], cnt[3];
  for (int a = 0; a < 3; a++) {
    const long long mc = (long long)std::floor(kmax / gp.dk[a]) + 1;
    // Signed indices present on the destination grid: the full range when it is
    // the parent grid itself, else strictly inside its Nyquist frequency.
    const long long smin = (sub == ctx) ? -(long long)(gs.n[a] - gs.n[a] / 2) : -(long long)((gs.n[a] - 1) / 2);
    const long long smax = (sub == ctx) ? (long long)(gs.n[a] / 2 - 1) : (long long)((gs.n[a] - 1) / 2);
    int lo_a = (int)std::max<long long>(-mc, smin);
    int hi_a = (int)std::min<long long>(mc, smax);
    if (gs.n[a] == 1) { lo_a = 0; hi_a = 0; }
    lo[a] = lo_a; cnt[a] = hi_a - lo_a + 1;
  }
  double* d_par = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(sub, (void**)&d_par, sizeof(double) * 3 * (size_t)nbins));
  std::vector<double> h_par(3 * (size_t)nbins);
  for (int q = 0; q < nbins; q++) {
    h_par[q] = klo[q]; h_par[nbins + q] = khi[q]; h_par[2 * nbins + q] = amp[q];
  }
  // Pageable source: the copy is staged before cudaMemcpyAsync returns.
  TRVB_CUDA(cudaMemcpyAsync(d_par, h_par.data(), sizeof(double) * h_par.size(),
                            cudaMemcpyHostToDevice, sub->stream));
  const RowLaunch rl = row_launch(ctx->num_sms, cnt[0], cnt[1], cnt[2]);
  int st = 0;
  for (int q0 = 0; q0 < nbins && st == 0; q0 += maxb) {
    const int nq = std::min(maxb, nbins - q0);
    void* out = static_cast<char*>(dst) + dst_stride * (size_t)q0;
    void* spec = real_out ? spec_tmp : out;
    TRVB_CUDA(cudaMemsetAsync(spec, 0, spec_bin_bytes * nq, sub->stream));
    ShellBatch sb; sb.klo = d_par + q0; sb.khi = d_par + nbins + q0; sb.amp = d_par + 2 * nbins + q0;
    sb.nbins = nq;
    k_shell_scatter<<<rl.grid, rl.block, 0, sub->stream>>>(
      kview_of(ctx, src), gp, gs, tables_of(ctx), ell, m, lo[0], lo[1], lo[2], cnt[0], cnt[1],
      cnt[2], sb, n2s, bin_stride, (double2*)spec);
    TRVB_LAUNCH_CHECK();
    cufftHandle plan;
    st = get_batch_plan(sub, real_out ? CUFFT_Z2D : CUFFT_Z2Z, nq, &plan);
    if (st) break;
    if (real_out) {
      TRVB_CUFFT(cufftExecZ2D(plan, (cufftDoubleComplex*)spec, (cufftDoubleReal*)out));
    } else {
      TRVB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)out, (cufftDoubleComplex*)out,
                              CUFFT_INVERSE));
    }
    g_trvb_fft_execs++;
  }
  trvb_dev_free_raw(sub, d_par);
  if (real_out) trvb_dev_free_raw(sub, spec_tmp);
  return st;
}

// ---------------------------------------------------------------------
// Pruned (and x-slab) form of the batched shell transform.
// ---------------------------------------------------------------------
//
// A shell [klo, khi) only has modes with |k_a| <= khi on every axis: on the sub-grid
// (n_s >= 4 m_cut + 2 per axis) at most a (2 m + 1)^2 (m + 1) corner block of the
// n_s^2 (n_s/2 + 1) half-spectrum is non-zero.  The 3-D inverse transform is taken one
// axis at a time and every pass only touches the lines that can be non-zero:
//   x: (2m+1)(m+1) lines  ->  y: n_x (m+1) lines  ->  z: n_x n_y lines (c2r),
// 6 % + 24 % + 100 % of the line transforms of a dense 3-D FFT at m = n_s/4, less for the
// inner shells (the extents follow the largest khi of each sub-batch of shells).  With
// x-slabs [x0, x0 + nx) the y and z passes run on the slab's planes only.

namespace {

// The filtered modes of the low-|k| cube, evaluated ONCE per call: value y_lm src / W x amp
// and the shell (index into the call's bin list, -1: none) of every mode (k_y, k_z >= 0,
// k_x), laid out [c = k_z][b = k_y][a = k_x]: the x-lines are contiguous.
__global__ void __launch_bounds__(256)
k_lowk_modes(KView src, GridDesc gp, Tables tb, int ell, int m, int lo0, int lo1,
             int K0, int K1, int K2, ShellBatch sb, double2* __restrict__ val,
             int* __restrict__ shell) {
  const YlmCoef yc = ylm_coef(ell, m);
  const long long total = (long long)K0 * K1 * K2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t % K0), b = (int)((t / K0) % K1), c = (int)(t / ((long long)K0 * K1));
    const int mi = lo0 + a, mj = lo1 + b, mk = c;
    const double kx = __dmul_rn((double)mi, gp.dk[0]);
    const double ky = __dmul_rn((double)mj, gp.dk[1]);
    const double kz = __dmul_rn((double)mk, gp.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    int q = -1;
    for (int i = 0; i < sb.nbins; i++) {
      const double lo = sb.klo[i], hi = sb.khi[i];
      if ((lo < 0. && hi < 0.) || (lo <= kmag && kmag < hi)) { q = i; break; }
    }
    shell[t] = q;
    if (q >= 0) {
      const double2 v = shell_mode(src, gp, tb, yc, mi, mj, mk, kx, ky, kz, 1.);
      const double amp = sb.amp[q];
      val[t] = make_double2(v.x * amp, v.y * amp);
    } else {
      val[t] = make_double2(0., 0.);
    }
  }
}

// Extents of the non-zero corner block of one sub-batch of shells (signed mode indices
// -mc[a] .. mc[a]; k_z >= 0 only) inside the call's cube (-gc[a] .. gc[a]).
struct PruneDims {
  int mc[3];    // sub-batch
  int gc[3];    // whole call: the layout of val / shell
};

// Pass 1 input: zero-padded x-lines A[q][c][b][x] of the nq shells of the sub-batch,
// c = k_z in [0, mc2], b = k_y + mc1 in [0, 2 mc1], x = the slot of k_x on the n0-line.
__global__ void __launch_bounds__(256)
k_shell_xlines(const double2* __restrict__ val, const int* __restrict__ shell, PruneDims pd,
               int q0, int nq, int n0, double2* __restrict__ A) {
  const int K1 = 2 * pd.mc[1] + 1, K2 = pd.mc[2] + 1;
  const int G0 = 2 * pd.gc[0] + 1, G1 = 2 * pd.gc[1] + 1;
  const long long lines = (long long)K1 * K2;
  const long long total = lines * n0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % n0);
    const long long col = t / n0;
    const int b = (int)(col % K1), c = (int)(col / K1);
    int mi = 0; bool live = true;
    if (x <= pd.mc[0]) mi = x;
    else if (x >= n0 - pd.mc[0]) mi = x - n0;
    else live = false;
    int q = -1;
    double2 v = make_double2(0., 0.);
    if (live) {
      const int mj = b - pd.mc[1];
      const long long g = ((long long)c * G1 + (mj + pd.gc[1])) * G0 + (mi + pd.gc[0]);
      q = shell[g] - q0;
      if (q >= 0 && q < nq) v = val[g];
    }
    for (int i = 0; i < nq; i++) {
      A[((long long)i * lines + col) * n0 + x] = (i == q) ? v : make_double2(0., 0.);
    }
  }
}

// Pass 2 input: B[q][xi][c][y] = A[q][c][b(y)][x0 + xi], zero where y is the slot of no k_y
// of the block.  One 32 x 32 tile (x by y) per step through shared memory.
__global__ void __launch_bounds__(256)
k_shell_ylines(const double2* __restrict__ A, PruneDims pd, int nq, int n0, int n1, int x0,
               int nx, double2* __restrict__ B) {
  __shared__ double2 tile[32][33];
  const int K1 = 2 * pd.mc[1] + 1, K2 = pd.mc[2] + 1;
  const int tiles_x = (nx + 31) / 32, tiles_y = (n1 + 31) / 32;
  const long long ntile = (long long)nq * K2 * tiles_x * tiles_y;
  for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int ty = (int)(t % tiles_y);
    const int tx = (int)((t / tiles_y) % tiles_x);
    const long long qc = t / ((long long)tiles_y * tiles_x);
    const int c = (int)(qc % K2), q = (int)(qc / K2);
    // rows of the tile: 32 y-slots; columns: 32 planes
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int y = ty * 32 + i, xi = tx * 32 + threadIdx.x;
      int b = -1;
      if (y < n1) {
        if (y <= pd.mc[1]) b = y + pd.mc[1];
        else if (y >= n1 - pd.mc[1]) b = y - n1 + pd.mc[1];
      }
      tile[i][threadIdx.x] = (b >= 0 && xi < nx)
        ? A[(((long long)q * K2 + c) * K1 + b) * n0 + x0 + xi] : make_double2(0., 0.);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int xi = tx * 32 + i, y = ty * 32 + threadIdx.x;
      if (xi < nx && y < n1) {
        B[((((long long)q * nx + xi) * K2) + c) * n1 + y] = tile[threadIdx.x][i];
      }
    }
    __syncthreads();
  }
}

// Pass 3 input: E[r][y][kz] = B[r][kz][y] for kz < K2, zero for K2 <= kz < kw, for every
// (shell, plane) r: the rows the c2r transform reads.  Entries kz >= kw are zero already
// (E is cleared once per call and the blocks of the sub-batches only grow).
__global__ void __launch_bounds__(256)
k_shell_zlines(const double2* __restrict__ B, int K2, int kw, int n1, int nh, long long nrows,
               double2* __restrict__ E) {
  __shared__ double2 tile[32][33];
  const int tiles_y = (n1 + 31) / 32, tiles_k = (kw + 31) / 32;
  const long long ntile = nrows * tiles_y * tiles_k;
  for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int tk = (int)(t % tiles_k);
    const int ty = (int)((t / tiles_k) % tiles_y);
    const long long r = t / ((long long)tiles_k * tiles_y);
    const double2* src = B + r * (long long)K2 * n1;
    double2* dst = E + r * (long long)n1 * nh;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int kz = tk * 32 + i, y = ty * 32 + threadIdx.x;
      tile[i][threadIdx.x] = (kz < K2 && y < n1) ? src[(long long)kz * n1 + y] : make_double2(0., 0.);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int y = ty * 32 + i, kz = tk * 32 + threadIdx.x;
      if (y < n1 && kz < kw) dst[(long long)y * nh + kz] = tile[threadIdx.x][i];
    }
    __syncthreads();
  }
}

// Pass 3, hand-written: pruned-input complex-to-real transform along z straight from
// B[r][kz][y] to the REAL slab (csrc/trvb_zpass.cuh).  One CTA = 2 LP adjacent lines of one
// (shell, plane) r; twiddles are read through L1 (an 8.6 KB table at N = 540), which leaves
// the shared memory to three resident tiles.
template <int N, int LP, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_shell_zpass(const double2* __restrict__ B, int K2, int n1, int tiles_y,
              const double2* __restrict__ tw, double* __restrict__ out) {
  using namespace xpass;
  extern __shared__ __align__(16) unsigned char zp_smem[];
  double2* tile = reinterpret_cast<double2*>(zp_smem);
  const int tid = threadIdx.x;
  const long long r = blockIdx.x / tiles_y;
  const int y0 = (int)(blockIdx.x % tiles_y) * 2 * LP;
  constexpr int NS = Radix<N>::NS;
  zstage_load<N, LP, NT>(tid, B + r * (long long)K2 * n1, K2, n1, y0, tile);
  __syncthreads();
  if constexpr (NS >= 4) { zstage<N, LP, NT, 3>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 3) { zstage<N, LP, NT, 2>(tid, tile, tw); __syncthreads(); }
  zstage<N, LP, NT, 1>(tid, tile, tw);
  __syncthreads();
  zstage_store<N, LP, NT>(tid, tile, tw, n1, y0, out + r * (long long)n1 * N);
}

template <int N>
int launch_zpass(trvb_ctx* sub, const double2* B, int K2, int n1, long long nrows, double* out) {
  // Small tiles, many CTAs: the kernel is latency-bound at 12 warps per SM (LP = 8 at
  // N = 540: long-scoreboard stalls, 2.1 TB/s; profiles/r02_ncu_k_shell_zpass540_lp8.txt).
  constexpr int LP = N <= 288 ? 8 : 4, NT = 128;
  constexpr size_t smem = sizeof(double2) * (size_t)LP * xpass::zp_pitch<N, LP>();
  constexpr int MINB = smem * 5 <= 220 * 1024 ? 5 : 4;
  const double2* tw = nullptr;
  int st = trvb_twiddle_table(sub, N, &tw);
  if (st) return st;
  static std::mutex attr_mutex;
  static std::map<int, bool> attr_done;   // per device
  {
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (!attr_done[sub->device]) {
      TRVB_CUDA(cudaFuncSetAttribute(k_shell_zpass<N, LP, NT, MINB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done[sub->device] = true;
    }
  }
  const int tiles_y = (n1 + 2 * LP - 1) / (2 * LP);
  const long long tiles = nrows * tiles_y;
  TRVB_REQUIRE(tiles < 2147483647LL, "pruned transform: too many z-pass tiles");
  k_shell_zpass<N, LP, NT, MINB><<<(unsigned)tiles, NT, smem, sub->stream>>>(B, K2, n1, tiles_y, tw, out);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Pass 2, hand-written: pruned-input complex-to-complex transform along y from A[q][c][b][x]
// to B[q][xi][c][y] (csrc/trvb_zpass.cuh).  One CTA = XT adjacent planes of one (q, c).
template <int N, int XT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_shell_ypass(const double2* __restrict__ A, int K1, int K2, int mc1, int n0, int x0, int nx,
              int tiles_x, const double2* __restrict__ tw, double2* __restrict__ B) {
  using namespace xpass;
  extern __shared__ __align__(16) unsigned char zp_smem[];
  double2* tile = reinterpret_cast<double2*>(zp_smem);
  const int tid = threadIdx.x;
  const long long qc = blockIdx.x / tiles_x;            // q K2 + c
  const int xi0 = (int)(blockIdx.x % tiles_x) * XT;
  const long long q = qc / K2;
  const int c = (int)(qc - q * K2);
  constexpr int NS = Radix<N>::NS;
  ystage_load<N, XT, NT>(tid, A + qc * (long long)K1 * n0, K1, mc1, n0, x0 + xi0, x0 + nx, tile);
  __syncthreads();
  if constexpr (NS >= 4) { zstage<N, XT, NT, 3>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 3) { zstage<N, XT, NT, 2>(tid, tile, tw); __syncthreads(); }
  zstage<N, XT, NT, 1>(tid, tile, tw);
  __syncthreads();
  ystage_store<N, XT, NT>(tid, tile, tw, xi0, nx, (long long)K2 * N,
                          B + (q * nx * K2 + c) * (long long)N);
}

template <int N>
int launch_ypass(trvb_ctx* sub, const double2* A, int nq, int K1, int K2, int mc1, int n0,
                 int x0, int nx, double2* B) {
  constexpr int XT = N <= 288 ? 8 : 4, NT = 128;
  constexpr size_t smem = sizeof(double2) * (size_t)XT * xpass::zp_pitch<N, XT>();
  constexpr int MINB = smem * 5 <= 220 * 1024 ? 5 : 4;
  const double2* tw = nullptr;
  int st = trvb_twiddle_table(sub, N, &tw);
  if (st) return st;
  static std::mutex attr_mutex;
  static std::map<int, bool> attr_done;   // per device
  {
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (!attr_done[sub->device]) {
      TRVB_CUDA(cudaFuncSetAttribute(k_shell_ypass<N, XT, NT, MINB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done[sub->device] = true;
    }
  }
  const int tiles_x = (nx + XT - 1) / XT;
  const long long tiles = (long long)nq * K2 * tiles_x;
  TRVB_REQUIRE(tiles < 2147483647LL, "pruned transform: too many y-pass tiles");
  k_shell_ypass<N, XT, NT, MINB><<<(unsigned)tiles, NT, smem, sub->stream>>>(
    A, K1, K2, mc1, n0, x0, nx, tiles_x, tw, B);
  TRVB_LAUNCH_CHECK();
  return 0;
}

#define TRVB_ZPASS_LENGTHS(X) \
  X(64) X(72) X(96) X(108) X(128) X(144) X(160) X(180) X(192) X(216) X(240) X(256) X(270) \
  X(288) X(320) X(360) X(384) X(432) X(480) X(512) X(540) X(576) X(600) X(640) X(720)

bool zpass_supported(int n2) {
  switch (n2) {
#define X(n) case n:
    TRVB_ZPASS_LENGTHS(X)
#undef X
      return true;
    default: return false;
  }
}

int run_zpass(trvb_ctx* sub, int n2, const double2* B, int K2, int n1, long long nrows,
              double* out) {
  switch (n2) {
#define X(n) case n: return launch_zpass<n>(sub, B, K2, n1, nrows, out);
    TRVB_ZPASS_LENGTHS(X)
#undef X
    default: break;
  }
  TRVB_REQUIRE(false, "pruned transform: no hand-written z pass for length %d", n2);
}

int run_ypass(trvb_ctx* sub, int n1, const double2* A, int nq, int K1, int K2, int mc1, int n0,
              int x0, int nx, double2* B) {
  switch (n1) {
#define X(n) case n: return launch_ypass<n>(sub, A, nq, K1, K2, mc1, n0, x0, nx, B);
    TRVB_ZPASS_LENGTHS(X)
#undef X
    default: break;
  }
  TRVB_REQUIRE(false, "pruned transform: no hand-written y pass for length %d", n1);
}

// Batched 1-D plans WITHOUT their own work areas (several batch sizes are alive at once and
// cuFFT's automatic areas would add up): the area comes from the arena for the duration of
// one execution.
int get_line_plan(trvb_ctx* ctx, cufftType type, int n, long long batch, cufftHandle* out,
                  size_t* work_bytes) {
  const std::vector<long long> key = {(long long)type, (long long)n, batch};
  auto it = ctx->line_plans.find(key);
  if (it == ctx->line_plans.end()) {
    TRVB_REQUIRE(batch < 2147483647LL, "pruned transform: batch too large");
    if (ctx->line_plans.size() >= 192) {   // binning after binning in one process: start over
      TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
      for (auto& kv : ctx->line_plans) cufftDestroy(kv.second);
      ctx->line_plans.clear(); ctx->line_plan_work.clear();
    }
    cufftHandle plan;
    TRVB_CUFFT(cufftCreate(&plan));
    TRVB_CUFFT(cufftSetAutoAllocation(plan, 0));
    int len[1] = {n};
    size_t ws = 0;
    cufftResult rc;
    if (type == CUFFT_Z2Z) {
      rc = cufftMakePlanMany(plan, 1, len, nullptr, 1, n, nullptr, 1, n, CUFFT_Z2Z, (int)batch, &ws);
    } else {
      int inembed[1] = {n / 2 + 1}, onembed[1] = {n};
      rc = cufftMakePlanMany(plan, 1, len, inembed, 1, n / 2 + 1, onembed, 1, n, CUFFT_Z2D,
                             (int)batch, &ws);
    }
    if (rc != CUFFT_SUCCESS) { cufftDestroy(plan); TRVB_CUFFT(rc); }
    TRVB_CUFFT(cufftSetStream(plan, ctx->stream));
    it = ctx->line_plans.emplace(key, plan).first;
    ctx->line_plan_work[key] = ws;
  }
  *out = it->second;
  *work_bytes = ctx->line_plan_work[key];
  return 0;
}

// Arena blocks of one call, returned on every exit.
struct ArenaBlocks {
  trvb_ctx* ctx;
  std::vector<void*> blocks;
  cudaError_t get(void** p, size_t bytes) {
    cudaError_t e = trvb_dev_alloc_raw(ctx, p, bytes);
    if (e == cudaSuccess) blocks.push_back(*p);
    return e;
  }
  ~ArenaBlocks() { for (void* b : blocks) trvb_dev_free_raw(ctx, b); }
};

}  // namespace

extern "C" int trvb_shell_zpass_supported(int n2) { return zpass_supported(n2) ? 1 : 0; }

extern "C" int trvb_shell_slab_batch(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh src, int ell,
                                     int m, const double* klo, const double* khi,
                                     const double* amp, int nbins, int x0, int nx, void* dst) {
  TRVB_REQUIRE(ctx && sub && src.data && dst && klo && khi && amp && nbins > 0,
               "trvb_shell_slab_batch: bad argument");
  TRVB_REQUIRE(ctx->parent == nullptr && sub->parent == ctx,
               "trvb_shell_slab_batch: `sub` must be a sub-grid of `ctx`");
  TRVB_REQUIRE(src.layout == TRVB_HALF && m == 0 && (ell % 2) == 0 && ell >= 0,
               "trvb_shell_slab_batch: needs a HALF source, m = 0 and even l (real shell fields)");
  const GridDesc& gp = ctx->g;
  const GridDesc& gs = sub->g;
  TRVB_REQUIRE(x0 >= 0 && nx > 0 && x0 + nx <= gs.n[0], "trvb_shell_slab_batch: planes [%d, %d) "
               "outside the %d planes of the sub-grid", x0, x0 + nx, gs.n[0]);
  const int n0 = gs.n[0], n1 = gs.n[1], n2 = gs.n[2], nh = gs.nh;
  // Largest signed mode index per axis that a shell reaching `kmax` can hold, strictly
  // inside the sub-grid's Nyquist frequency.
  auto extent_of = [&](double kmax, bool unbounded, int mc[3]) {
    for (int a = 0; a < 3; a++) {
      const long long smax = (long long)((gs.n[a] - 1) / 2);
      long long v = unbounded ? smax : (long long)std::floor(kmax / gp.dk[a]) + 1;
      mc[a] = (int)std::min(v, smax);
    }
  };
  PruneDims pd;
  {
    double kmax = 0.; bool unbounded = false;
    for (int q = 0; q < nbins; q++) {
      if (klo[q] < 0. && khi[q] < 0.) unbounded = true;
      kmax = std::max(kmax, khi[q]);
    }
    extent_of(kmax, unbounded, pd.gc);
  }
  const int G0 = 2 * pd.gc[0] + 1, G1 = 2 * pd.gc[1] + 1, G2 = pd.gc[2] + 1;

  ArenaBlocks arena{sub, {}};
  double* d_par = nullptr;
  TRVB_CUDA(arena.get((void**)&d_par, sizeof(double) * 3 * (size_t)nbins));
  std::vector<double> h_par(3 * (size_t)nbins);
  for (int q = 0; q < nbins; q++) {
    h_par[q] = klo[q]; h_par[nbins + q] = khi[q]; h_par[2 * nbins + q] = amp[q];
  }
  // Pageable source: the copy is staged before cudaMemcpyAsync returns.
  TRVB_CUDA(cudaMemcpyAsync(d_par, h_par.data(), sizeof(double) * h_par.size(),
                            cudaMemcpyHostToDevice, sub->stream));

  // The modes of the low-|k| cube, once for the whole call.
  const long long nmodes = (long long)G0 * G1 * G2;
  double2* d_val = nullptr; int* d_shell = nullptr;
  TRVB_CUDA(arena.get((void**)&d_val, sizeof(double2) * (size_t)nmodes));
  TRVB_CUDA(arena.get((void**)&d_shell, sizeof(int) * (size_t)nmodes));
  {
    ShellBatch all; all.klo = d_par; all.khi = d_par + nbins; all.amp = d_par + 2 * nbins;
    all.nbins = nbins;
    const int blocks = (int)std::min<long long>((nmodes + 255) / 256, (long long)ctx->num_sms * 16);
    k_lowk_modes<<<blocks, 256, 0, sub->stream>>>(kview_of(ctx, src), gp, tables_of(ctx), ell, m,
                                                  -pd.gc[0], -pd.gc[1], G0, G1, G2, all, d_val,
                                                  d_shell);
    TRVB_LAUNCH_CHECK();
  }

  // Sub-batches of consecutive shells: they bound the transient arrays to ~6 GiB, and each
  // takes the extents of its own largest shell (TRV_SHELL_GROUPS = least number of them).
  // The z pass: hand-written pruned c2r (k_shell_zpass) for the lengths it is built for,
  // else zero-padded lines + cuFFT (TRV_NO_ZPASS=1 forces the latter).
  bool own_z = zpass_supported(n2), own_y = zpass_supported(n1);
  {
    const char* env = getenv("TRV_NO_ZPASS");
    if (env && env[0] == '1') own_z = own_y = false;
    env = getenv("TRV_NO_YPASS");
    if (env && env[0] == '1') own_y = false;
  }
  const size_t a_bin = sizeof(double2) * (size_t)G1 * G2 * n0;
  const size_t b_bin = sizeof(double2) * (size_t)nx * G2 * n1;
  const size_t e_bin = own_z ? 0 : sizeof(double2) * (size_t)nx * n1 * nh;
  int maxb = (int)std::max<size_t>(1, ((size_t)6 << 30) / (a_bin + b_bin + e_bin));
  {
    // (C5 shares of 8, 40 shells: 4 groups 13.0 ms, 10 groups 12.0 ms)
    const char* env = getenv("TRV_SHELL_GROUPS");
    const int groups = std::max(1, env ? atoi(env) : std::min(10, std::max(4, nbins / 4)));
    maxb = std::min(maxb, (nbins + groups - 1) / groups);
  }
  maxb = std::max(1, std::min(maxb, nbins));
  double2* A = nullptr; double2* B = nullptr; double2* E = nullptr;
  TRVB_CUDA(arena.get((void**)&A, a_bin * maxb));
  TRVB_CUDA(arena.get((void**)&B, b_bin * maxb));
  if (!own_z) {
    TRVB_CUDA(arena.get((void**)&E, e_bin * maxb));
    TRVB_CUDA(cudaMemsetAsync(E, 0, e_bin * maxb, sub->stream));
  }
  int e_written = 0;   // kz extent of E that may hold non-zero values
  const size_t out_bin = sizeof(double) * (size_t)nx * n1 * n2;
  auto exec_with_area = [&](cufftHandle plan, size_t ws, auto run) -> int {
    void* area = nullptr;
    if (ws) TRVB_CUDA(trvb_dev_alloc_raw(sub, &area, ws));
    if (ws) TRVB_CUFFT(cufftSetWorkArea(plan, area));
    int rc = run();
    if (ws) trvb_dev_free_raw(sub, area);   // stream-ordered reuse
    return rc;
  };
  const long long cap = (long long)ctx->num_sms * 16;
  for (int q0 = 0; q0 < nbins; q0 += maxb) {
    const int nq = std::min(maxb, nbins - q0);
    {
      double kmax = 0.; bool unbounded = false;
      for (int q = q0; q < q0 + nq; q++) {
        if (klo[q] < 0. && khi[q] < 0.) unbounded = true;
        kmax = std::max(kmax, khi[q]);
      }
      extent_of(kmax, unbounded, pd.mc);
    }
    const int K1 = 2 * pd.mc[1] + 1, K2 = pd.mc[2] + 1;
    const long long lines_x = (long long)nq * K1 * K2;
    // -- x --
    {
      const long long work = (long long)K1 * K2 * n0;
      k_shell_xlines<<<(int)std::min((work + 255) / 256, cap), 256, 0, sub->stream>>>(
        d_val, d_shell, pd, q0, nq, n0, A);
      TRVB_LAUNCH_CHECK();
    }
    cufftHandle plan; size_t ws = 0;
    int st = get_line_plan(sub, CUFFT_Z2Z, n0, lines_x, &plan, &ws);
    if (st) return st;
    st = exec_with_area(plan, ws, [&]() -> int {
      TRVB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)A, (cufftDoubleComplex*)A, CUFFT_INVERSE));
      return 0;
    });
    if (st) return st;
    g_trvb_fft_execs++;
    // -- y --
    if (own_y) {
      st = run_ypass(sub, n1, A, nq, K1, K2, pd.mc[1], n0, x0, nx, B);
      if (st) return st;
    } else {
      {
        const long long ntile = (long long)nq * K2 * ((nx + 31) / 32) * ((n1 + 31) / 32);
        k_shell_ylines<<<(int)std::min(ntile, cap), dim3(32, 8), 0, sub->stream>>>(
          A, pd, nq, n0, n1, x0, nx, B);
        TRVB_LAUNCH_CHECK();
      }
      st = get_line_plan(sub, CUFFT_Z2Z, n1, (long long)nq * nx * K2, &plan, &ws);
      if (st) return st;
      st = exec_with_area(plan, ws, [&]() -> int {
        TRVB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)B, (cufftDoubleComplex*)B, CUFFT_INVERSE));
        return 0;
      });
      if (st) return st;
      g_trvb_fft_execs++;
    }
    // -- z --
    const long long nrows = (long long)nq * nx;
    if (own_z) {
      double* out = (double*)((char*)dst + out_bin * (size_t)q0);
      st = run_zpass(sub, n2, B, K2, n1, nrows, out);
      if (st) return st;
      continue;
    }
    const int kw = std::min(nh, std::max(K2, e_written));
    e_written = kw;
    {
      const long long ntile = nrows * ((n1 + 31) / 32) * ((kw + 31) / 32);
      k_shell_zlines<<<(int)std::min(ntile, cap), dim3(32, 8), 0, sub->stream>>>(
        B, K2, kw, n1, nh, nrows, E);
      TRVB_LAUNCH_CHECK();
    }
    st = get_line_plan(sub, CUFFT_Z2D, n2, nrows * n1, &plan, &ws);
    if (st) return st;
    double* out = (double*)((char*)dst + out_bin * (size_t)q0);
    st = exec_with_area(plan, ws, [&]() -> int {
      TRVB_CUFFT(cufftExecZ2D(plan, (cufftDoubleComplex*)E, (cufftDoubleReal*)out));
      return 0;
    });
    if (st) return st;
    g_trvb_fft_execs++;
  }
  return 0;
}

extern "C" int trvb_sjl_ifft_batch(trvb_ctx* ctx, trvb_mesh src, int ell, int m,
                                   const double* r, double amp, int nbins, void* dst) {
  TRVB_REQUIRE(ctx && src.data && r && dst && nbins > 0, "trvb_sjl_ifft_batch: bad argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_sjl_ifft_batch: root context only");
  TRVB_REQUIRE(src.layout != TRVB_REAL, "trvb_sjl_ifft_batch: src must be a Fourier mesh");
  auto it = ctx->sjl.find(ell);
  TRVB_REQUIRE(it != ctx->sjl.end(), "trvb_sjl_ifft_batch: no spline table for ell = %d "
               "(call trvb_sjl_table first)", ell);
  SjlView sj;
  sj.y = it->second.d_y; sj.c = it->second.d_c;
  sj.nsample = it->second.nsample; sj.step = it->second.step; sj.ell = ell;
  const GridDesc& g = ctx->g;
  const size_t mesh_bytes = trvb_mesh_bytes(ctx, TRVB_COMPLEX);
  double2* A = nullptr; double* kmag = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&A, mesh_bytes));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&kmag, mesh_bytes / 2));
  const RowLaunch rl = row_launch(ctx->num_sms, g.n[0], g.n[1], g.n[2]);
  k_sjl_prepare<<<rl.grid, rl.block, 0, ctx->stream>>>(
    kview_of(ctx, src), g, tables_of(ctx), ell, m, amp, A, kmag);
  TRVB_LAUNCH_CHECK();
  const int blocks = (int)std::min<long long>((g.nmesh + 255) / 256, (long long)ctx->num_sms * 16);
  int st = 0;
  for (int q0 = 0; q0 < nbins && st == 0; q0 += SJL_NB) {
    SjlBatch b;
    b.count = std::min(SJL_NB, nbins - q0);
    for (int q = 0; q < SJL_NB; q++) {
      const int qq = q0 + std::min(q, b.count - 1);
      b.r[q] = r[qq];
      b.dst[q] = (double2*)((char*)dst + mesh_bytes * (size_t)qq);
    }
    k_sjl_apply<<<blocks, 256, 0, ctx->stream>>>(A, kmag, g.nmesh, sj, b);
    TRVB_LAUNCH_CHECK();
    for (int q = 0; q < b.count && st == 0; q++) {
      trvb_mesh out; out.data = b.dst[q]; out.layout = TRVB_COMPLEX; out.k0_add = 0.;
      st = trvb_fft_inverse(ctx, out, out);
    }
  }
  trvb_dev_free_raw(ctx, A);      // stream-ordered reuse
  trvb_dev_free_raw(ctx, kmag);
  return st;
}

extern "C" int trvb_sjl_ifft(trvb_ctx* ctx, trvb_mesh src, int ell, int m, double r,
                             double amp, trvb_mesh dst) {
  TRVB_REQUIRE(ctx && src.data && dst.data, "trvb_sjl_ifft: null argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_sjl_ifft: root context only");
  TRVB_REQUIRE(src.layout != TRVB_REAL && dst.layout == TRVB_COMPLEX && dst.data != src.data,
               "trvb_sjl_ifft: src must be a Fourier mesh, dst a distinct COMPLEX mesh");
  auto it = ctx->sjl.find(ell);
  TRVB_REQUIRE(it != ctx->sjl.end(), "trvb_sjl_ifft: no spline table for ell = %d "
               "(call trvb_sjl_table first)", ell);
  SjlView sj;
  sj.y = it->second.d_y; sj.c = it->second.d_c;
  sj.nsample = it->second.nsample; sj.step = it->second.step; sj.ell = ell;
  const RowLaunch rl = row_launch(ctx->num_sms, ctx->g.n[0], ctx->g.n[1], ctx->g.n[2]);
  k_sjl_spectrum<<<rl.grid, rl.block, 0, ctx->stream>>>(
    kview_of(ctx, src), ctx->g, tables_of(ctx), sj, ell, m, r, amp, (double2*)dst.data);
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(ctx, dst, dst);
}
