// trvb_fourier.cu -- transforms and Fourier-space field construction.
//
// cuFFT is the only library call (forward/inverse 3-D FFT, Z2Z / D2Z / Z2D);
// everything around it -- scaling, mean subtraction, interlacing, window
// compensation, shell filtering with y_lm weights, spherical-Bessel weighting
// -- is a hand-written vectorised kernel adjacent to the transform.
// Replaces S/field.cpp:1496-1720 (transforms), 1764-1785 (compensation),
// 1792-1906 (band-limited y_lm-weighted IFFT), 1908-2010 (j_l-weighted IFFT).
#include "trvb_common.cuh"

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
__global__ void k_interlace(double2* __restrict__ f, const double2* __restrict__ fs,
                            GridDesc g, int layout) {
  const int n2s = kdim2(g, layout);
  const long long total = (long long)g.n[0] * g.n[1] * n2s;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % n2s);
    const int j = (int)((t / n2s) % g.n[1]);
    const int i = (int)(t / ((long long)n2s * g.n[1]));
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
  }
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

__global__ void k_compensate(double2* __restrict__ f, GridDesc g, int layout, Tables tb) {
  const int n2s = kdim2(g, layout);
  const long long total = (long long)g.n[0] * g.n[1] * n2s;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % n2s);
    const int j = (int)((t / n2s) % g.n[1]);
    const int i = (int)(t / ((long long)n2s * g.n[1]));
    const double w = window_at(tb, g.order, i, j, k);
    double2 a = f[t];
    a.x /= w; a.y /= w;
    f[t] = a;
  }
}

// Shell-filtered, y_lm-weighted, window-compensated spectrum on the sub grid
// (S/field.cpp:1815-1847).  One thread per sub-grid Fourier cell.
//   gs = sub grid, gp = parent grid (tables and `src` live on the parent).
__global__ void __launch_bounds__(256)
k_shell_spectrum(KView src, GridDesc gp, GridDesc gs, Tables tb, int ell, int m,
                 double klo, double khi, int use_shell, double amp,
                 double2* __restrict__ dst) {
  const bool same = gs.n[0] == gp.n[0] && gs.n[1] == gp.n[1] && gs.n[2] == gp.n[2];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < gs.nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const int ks = (int)(t % gs.n[2]);
    const int js = (int)((t / gs.n[2]) % gs.n[1]);
    const int is = (int)(t / ((long long)gs.n[2] * gs.n[1]));
    int mi, mj, mk, ip, jp, kp;
    bool rep = true;
    if (same) {
      ip = is; jp = js; kp = ks;
      mi = signed_index(ip, gp.n[0]); mj = signed_index(jp, gp.n[1]);
      mk = signed_index(kp, gp.n[2]);
    } else {
      mi = signed_index(is, gs.n[0]); mj = signed_index(js, gs.n[1]);
      mk = signed_index(ks, gs.n[2]);
      // Only modes strictly inside the sub-grid Nyquist are representable
      // on both grids without ambiguity.
      rep = (2 * abs(mi) < gs.n[0]) && (2 * abs(mj) < gs.n[1]) && (2 * abs(mk) < gs.n[2]);
      ip = mi >= 0 ? mi : mi + gp.n[0];
      jp = mj >= 0 ? mj : mj + gp.n[1];
      kp = mk >= 0 ? mk : mk + gp.n[2];
    }
    double2 out = make_double2(0., 0.);
    if (rep) {
      // kv = i * dk (S/field.cpp:555-562), |k| without contraction.
      const double kx = __dmul_rn((double)mi, gp.dk[0]);
      const double ky = __dmul_rn((double)mj, gp.dk[1]);
      const double kz = __dmul_rn((double)mk, gp.dk[2]);
      const double kmag = vec3_norm_exact(kx, ky, kz);
      if (!use_shell || (klo <= kmag && kmag < khi)) {
        cplx fk = kload(src, ip, jp, kp);
        const double w = window_at(tb, gp.order, ip, jp, kp);
        fk.re /= w; fk.im /= w;
        cplx y = ylm_reduced(ell, m, kx, ky, kz);
        cplx v = cmul(y, fk);
        out.x = v.re * amp; out.y = v.im * amp;
      }
    }
    dst[t] = out;
  }
}

// j_l(|k| r) y_lm(khat) src(k)/W(k) * amp on the full grid (S/field.cpp:1936-1961).
__global__ void __launch_bounds__(256)
k_sjl_spectrum(KView src, GridDesc g, Tables tb, SjlView sj, int ell, int m,
               double r, double amp, double2* __restrict__ dst) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nmesh;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % g.n[2]);
    const int j = (int)((t / g.n[2]) % g.n[1]);
    const int i = (int)(t / ((long long)g.n[2] * g.n[1]));
    const double kx = __dmul_rn((double)signed_index(i, g.n[0]), g.dk[0]);
    const double ky = __dmul_rn((double)signed_index(j, g.n[1]), g.dk[1]);
    const double kz = __dmul_rn((double)signed_index(k, g.n[2]), g.dk[2]);
    const double kmag = vec3_norm_exact(kx, ky, kz);
    cplx fk = kload(src, i, j, k);
    const double w = window_at(tb, g.order, i, j, k);
    fk.re /= w; fk.im /= w;
    cplx y = ylm_reduced(ell, m, kx, ky, kz);
    cplx v = cmul(y, fk);
    const double jl = sjl_eval(sj, __dmul_rn(kmag, r));
    dst[t] = make_double2(jl * v.re * amp, jl * v.im * amp);
  }
}

int get_plan(trvb_ctx* ctx, cufftType type, cufftHandle* out) {
  cufftHandle* slot; bool* has;
  if (type == CUFFT_Z2Z) { slot = &ctx->plan_z2z; has = &ctx->has_z2z; }
  else if (type == CUFFT_D2Z) { slot = &ctx->plan_d2z; has = &ctx->has_d2z; }
  else { slot = &ctx->plan_z2d; has = &ctx->has_z2d; }
  if (!*has) {
    TRVB_CUFFT(cufftPlan3d(slot, ctx->g.n[0], ctx->g.n[1], ctx->g.n[2], type));
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

KView kview_of(const trvb_ctx* ctx, trvb_mesh m) {
  KView v;
  v.p = (const double2*)m.data; v.layout = m.layout;
  v.n0 = ctx->g.n[0]; v.n1 = ctx->g.n[1]; v.n2 = ctx->g.n[2]; v.nh = ctx->g.nh;
  return v;
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
  const long long n = (long long)ctx->g.n[0] * ctx->g.n[1] * kdim2(ctx->g, kmesh.layout);
  k_interlace<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(
    (double2*)kmesh.data, (const double2*)kmesh_s.data, ctx->g, kmesh.layout);
  TRVB_LAUNCH_CHECK();
  return 0;
}

extern "C" int trvb_compensate(trvb_ctx* ctx, trvb_mesh kmesh) {
  TRVB_REQUIRE(ctx && kmesh.data && kmesh.layout != TRVB_REAL,
               "trvb_compensate: needs a Fourier-space mesh");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_compensate: root context only");
  const long long n = (long long)ctx->g.n[0] * ctx->g.n[1] * kdim2(ctx->g, kmesh.layout);
  k_compensate<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(
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
  TRVB_REQUIRE(src.layout != TRVB_REAL && dst.layout == TRVB_COMPLEX,
               "trvb_shell_ifft: src must be a Fourier mesh, dst COMPLEX");
  TRVB_REQUIRE(abs(m) <= ell && ell >= 0, "trvb_shell_ifft: bad (l, m) = (%d, %d)", ell, m);
  TRVB_REQUIRE(dst.data != src.data || sub == ctx, "trvb_shell_ifft: aliasing across grids");
  TRVB_REQUIRE(dst.data != src.data, "trvb_shell_ifft: src and dst must differ");
  const int use_shell = !(klo < 0. && khi < 0.);
  k_shell_spectrum<<<grid_for(sub, sub->g.nmesh, 256), 256, 0, ctx->stream>>>(
    kview_of(ctx, src), ctx->g, sub->g, tables_of(ctx), ell, m, klo, khi, use_shell,
    amp, (double2*)dst.data);
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(sub, dst, dst);
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
  k_sjl_spectrum<<<grid_for(ctx, ctx->g.nmesh, 256), 256, 0, ctx->stream>>>(
    kview_of(ctx, src), ctx->g, tables_of(ctx), sj, ell, m, r, amp, (double2*)dst.data);
  TRVB_LAUNCH_CHECK();
  return trvb_fft_inverse(ctx, dst, dst);
}
