// trvb_dist.cu -- the mesh phase of the periodic-box estimators spread over the GPUs of
// one NCCL communicator (include/trvb.h: trvb_dmesh_*).
//
// The reference's multi-GPU mode hands the full-grid transforms to cuFFT-Xt from one
// process (S/field.cpp:212-235); with the grid replicated per GPU instead (trvb_comm.cu)
// the particle assignment and the two full-grid FFTs are the part of an estimator call
// that does not shrink with the number of GPUs (C5 on 8 GPUs: 40 of 64 ms).  Here the R
// ranks own
//   * x-slabs of the configuration-space mesh: rank r assigns the particles whose stencil
//     meets its n0/R planes (a window of those planes plus a margin, filled through the
//     ordinary assignment kernels of a window-sized context) and 2-D transforms them;
//   * k_y-slabs of the Fourier-space mesh: after ONE all-to-all (each rank sends (R-1)/R
//     of its slab: 8.6 GB in total at 1024^3, 1.1 GB per GPU over NVSwitch) the transform
//     along x runs on n1/R rows of k_y.
// What the pair phase reads -- the modes the sub-grid represents, a few per cent of the
// mesh -- is gathered onto every rank (grouped broadcasts of each rank's rows) and read
// through KView's low-|k| storage; the shot-noise xi(r) takes the inverse route (1-D
// along x, all-to-all, 2-D on the planes) and is reduced to its radial histogram per
// slab, summed over the ranks (trvb_shot_bispec_reduce_slab).
#include "trvb_common.cuh"

#include <algorithm>
#include <string>
#include <vector>

struct trvb_dmesh {
  trvb_ctx* ctx = nullptr;
  trvb_comm* comm = nullptr;
  int R = 1, r = 0;
  int nx = 0, nj = 0;          // planes / k_y rows per rank
  int margin = 4;              // window planes either side of the slab
  trvb_ctx* win = nullptr;     // context of the assignment window (nx + 2 margin planes)
  cufftHandle plan_fwd = 0, plan_inv = 0, plan_x = 0;
  size_t ws_fwd = 0, ws_inv = 0, ws_x = 0;
  bool has_fwd = false, has_inv = false, has_x = false;
  double2* U = nullptr;        // delta n(k): [n0][nj][nh], this rank's rows of k_y
  double2* lowk_T = nullptr;   // transient of the gather
};

namespace {

constexpr int DIST_MARGIN = 4;
std::atomic<long long> g_dmesh_density_calls{0};

// TRV_DIST_TRACE=1: stage times of the distributed calls (CUDA events on the stream), printed
// to stderr by every rank.
struct StageTrace {
  bool on; cudaStream_t stream; const char* what; int rank;
  std::vector<std::pair<const char*, cudaEvent_t> > marks;
  StageTrace(const char* what_, cudaStream_t s, int rank_) : stream(s), what(what_), rank(rank_) {
    const char* env = getenv("TRV_DIST_TRACE");
    on = env && env[0] == '1';
    mark("start");
  }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, stream);
    marks.emplace_back(name, e);
  }
  ~StageTrace() {
    if (!on) return;
    cudaStreamSynchronize(stream);
    std::string line = std::string("[dist] rank ") + std::to_string(rank) + " " + what + ":";
    for (size_t i = 1; i < marks.size(); i++) {
      float ms = 0.f; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      char buf[96]; snprintf(buf, sizeof buf, " %s %.3f", marks[i].first, ms);
      line += buf;
    }
    fprintf(stderr, "%s\n", line.c_str());
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
};

struct Tables { const double* ralias[3]; };   // 1 / (per-axis factor of C1(k)), parent grid
Tables tables_of(const trvb_ctx* ctx) {
  Tables t;
  for (int a = 0; a < 3; a++) t.ralias[a] = ctx->d_ralias[a];
  return t;
}

// Particles whose home cell along x lies in [c_lo, c_lo + width) (periodic), appended to the
// output arrays with x shifted into the window's frame.  A block takes 1024 particles per
// step and reserves their output slots with ONE atomic (a counter bumped per warp is the
// bottleneck at 1e8 particles: millions of serialised atomics on one address).
__global__ void __launch_bounds__(256)
k_window_select(const double* __restrict__ x, const double* __restrict__ y,
                const double* __restrict__ z, long long n, double inv_dx, int n0, int c_lo,
                int width, double shift, double L0, double* __restrict__ ox,
                double* __restrict__ oy, double* __restrict__ oz,
                unsigned long long* __restrict__ counter) {
  constexpr int PER = 4;
  __shared__ unsigned prefix[PER * 8];
  __shared__ unsigned long long base_slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nchunks = (n + 256 * PER - 1) / (256 * PER);
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    double px[PER]; unsigned ballot[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const long long t = chunk * (256 * PER) + u * 256 + threadIdx.x;
      bool take = false;
      px[u] = 0.;
      if (t < n) {
        px[u] = x[t];
        int c = (int)floor(px[u] * inv_dx);
        c = min(max(c, 0), n0 - 1);
        int d = (c - c_lo) % n0;
        if (d < 0) d += n0;
        take = d < width;
      }
      ballot[u] = __ballot_sync(0xffffffffu, take);
      if (lane == 0) prefix[u * 8 + warp] = __popc(ballot[u]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned total = 0;
      for (int e = 0; e < PER * 8; e++) { const unsigned c = prefix[e]; prefix[e] = total; total += c; }
      base_slot = total ? atomicAdd(counter, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; u++) {
      if ((ballot[u] >> lane) & 1u) {
        const long long t = chunk * (256 * PER) + u * 256 + threadIdx.x;
        const unsigned long long slot = base_slot + prefix[u * 8 + warp]
          + __popc(ballot[u] & ((1u << lane) - 1u));
        double xs = px[u] - shift;
        if (xs < 0.) xs += L0;
        if (xs >= L0) xs -= L0;
        ox[slot] = xs; oy[slot] = y[t]; oz[slot] = z[t];
      }
    }
    __syncthreads();
  }
}

// T[xl][j][k] -> S[q][xl][jl][k] with j = q nj + jl: the block that goes to rank q is contiguous.
__global__ void __launch_bounds__(256)
k_pack_rows(const double2* __restrict__ T, int nx, int n1, int nj, int nh,
            double2* __restrict__ S) {
  for_each_cell(nx, n1, nh, [&](int xl, int j, int k, long long t) {
    const int q = j / nj, jl = j - q * nj;
    S[(((long long)q * nx + xl) * nj + jl) * nh + k] = T[t];
  });
}

// the inverse: S[q][xl][jl][k] -> T[xl][q nj + jl][k]
__global__ void __launch_bounds__(256)
k_unpack_rows(const double2* __restrict__ S, int nx, int n1, int nj, int nh,
              double2* __restrict__ T) {
  for_each_cell(nx, n1, nh, [&](int xl, int j, int k, long long t) {
    const int q = j / nj, jl = j - q * nj;
    T[t] = S[(((long long)q * nx + xl) * nj + jl) * nh + k];
  });
}

__global__ void k_add_first(double2* p, double add_re) { p[0].x += add_re; }

// Rows of the sub-grid spectrum whose k_y this rank holds, in k_y-major order:
// Tsub[js][is][ks] = U[i(is)][j(js) - j0][ks] (0 for the sub-grid's Nyquist planes).
__global__ void __launch_bounds__(256)
k_lowk_rows(const double2* __restrict__ U, int n0, int n1, int nh, int j0, int nj, int s0,
            int s1, int s2, int owner_rank, int rank_nj, double2* __restrict__ Tsub) {
  const int sh = s2 / 2 + 1;
  for_each_cell(s1, s0, sh, [&](int js, int is, int ks, long long t) {
    const int mj = signed_index(js, s1);
    const bool rep_j = 2 * abs(mj) < s1;
    const int j = mj >= 0 ? mj : mj + n1;
    const int owner = rep_j ? j / rank_nj : 0;
    if (owner != owner_rank) return;
    const int mi = signed_index(is, s0);
    double2 v = make_double2(0., 0.);
    if (rep_j && 2 * abs(mi) < s0 && 2 * ks < s2) {
      const int i = mi >= 0 ? mi : mi + n0;
      v = U[((long long)i * nj + (j - j0)) * nh + ks];
    }
    Tsub[t] = v;
  });
}

// dst[is][js][ks] = Tsub[js][is][ks]
__global__ void __launch_bounds__(256)
k_lowk_transpose(const double2* __restrict__ Tsub, int s0, int s1, int sh,
                 double2* __restrict__ dst) {
  for_each_cell(s0, s1, sh, [&](int is, int js, int ks, long long t) {
    dst[t] = Tsub[((long long)js * s0 + is) * sh + ks];
  });
}

// (fa conj(fb) / C1 - S) / V on this rank's rows of k_y, for fa = U + add_a delta_k0 and
// fb = U + add_b delta_k0 (S/field.cpp:3273-3298; non-interlaced meshes).
__global__ void __launch_bounds__(256)
k_shot_spectrum_rows(const double2* __restrict__ U, GridDesc g, Tables tb, int j0, int nj,
                     double add_a, double add_b, double S_re, double S_im,
                     double2* __restrict__ dst) {
  const double inv_vol = 1. / g.vol;
  for_each_cell(g.n[0], nj, g.nh, [&](int i, int jl, int k, long long t) {
    const int j = j0 + jl;
    double2 a = U[t], b = a;
    if ((i | j | k) == 0) { a.x += add_a; b.x += add_b; }
    const double rc1 = tb.ralias[0][i] * tb.ralias[1][j] * tb.ralias[2][k];
    const double re = (a.x * b.x + a.y * b.y) * rc1 - S_re;
    const double im = (a.y * b.x - a.x * b.y) * rc1 - S_im;
    dst[t] = make_double2(re * inv_vol, im * inv_vol);
  });
}

int exec_with_area(trvb_ctx* ctx, cufftHandle plan, size_t ws, cufftResult (*run)(cufftHandle, void*, void*, int),
                   void* in, void* out, int dir) {
  void* area = nullptr;
  if (ws) TRVB_CUDA(trvb_dev_alloc_raw(ctx, &area, ws));
  if (ws) TRVB_CUFFT(cufftSetWorkArea(plan, area));
  cufftResult rc = run(plan, in, out, dir);
  if (ws) trvb_dev_free_raw(ctx, area);   // stream-ordered reuse
  TRVB_CUFFT(rc);
  g_trvb_fft_execs++;
  return 0;
}

cufftResult run_d2z(cufftHandle p, void* in, void* out, int) {
  return cufftExecD2Z(p, (cufftDoubleReal*)in, (cufftDoubleComplex*)out);
}
cufftResult run_z2d(cufftHandle p, void* in, void* out, int) {
  return cufftExecZ2D(p, (cufftDoubleComplex*)in, (cufftDoubleReal*)out);
}
cufftResult run_z2z(cufftHandle p, void* in, void* out, int dir) {
  return cufftExecZ2Z(p, (cufftDoubleComplex*)in, (cufftDoubleComplex*)out, dir);
}

int ensure_plans(trvb_dmesh* dm) {
  trvb_ctx* ctx = dm->ctx;
  const GridDesc& g = ctx->g;
  int n2d[2] = {g.n[1], g.n[2]};
  int re_embed[2] = {g.n[1], g.n[2]}, cx_embed[2] = {g.n[1], g.nh};
  if (!dm->has_fwd) {
    TRVB_CUFFT(cufftCreate(&dm->plan_fwd));
    TRVB_CUFFT(cufftSetAutoAllocation(dm->plan_fwd, 0));
    TRVB_CUFFT(cufftMakePlanMany(dm->plan_fwd, 2, n2d, re_embed, 1, g.n[1] * g.n[2], cx_embed, 1,
                                 g.n[1] * g.nh, CUFFT_D2Z, dm->nx, &dm->ws_fwd));
    TRVB_CUFFT(cufftSetStream(dm->plan_fwd, ctx->stream));
    dm->has_fwd = true;
  }
  if (!dm->has_inv) {
    TRVB_CUFFT(cufftCreate(&dm->plan_inv));
    TRVB_CUFFT(cufftSetAutoAllocation(dm->plan_inv, 0));
    TRVB_CUFFT(cufftMakePlanMany(dm->plan_inv, 2, n2d, cx_embed, 1, g.n[1] * g.nh, re_embed, 1,
                                 g.n[1] * g.n[2], CUFFT_Z2D, dm->nx, &dm->ws_inv));
    TRVB_CUFFT(cufftSetStream(dm->plan_inv, ctx->stream));
    dm->has_inv = true;
  }
  if (!dm->has_x) {
    // lines along x of [n0][nj][nh]: element stride nj nh, consecutive lines one apart
    int n1d[1] = {g.n[0]}, embed[1] = {g.n[0]};
    const int lines = dm->nj * g.nh;
    TRVB_CUFFT(cufftCreate(&dm->plan_x));
    TRVB_CUFFT(cufftSetAutoAllocation(dm->plan_x, 0));
    TRVB_CUFFT(cufftMakePlanMany(dm->plan_x, 1, n1d, embed, lines, 1, embed, lines, 1, CUFFT_Z2Z,
                                 lines, &dm->ws_x));
    TRVB_CUFFT(cufftSetStream(dm->plan_x, ctx->stream));
    dm->has_x = true;
  }
  return 0;
}

}  // namespace

extern "C" long long trvb_dmesh_call_count(void) { return g_dmesh_density_calls; }

extern "C" int trvb_dmesh_supported(const trvb_ctx* ctx, int nranks) {
  if (!ctx || ctx->parent || nranks < 2) return 0;
  const GridDesc& g = ctx->g;
  if (g.n[0] % nranks || g.n[1] % nranks) return 0;
  if (g.n[0] / nranks < 2 * DIST_MARGIN) return 0;   // windows of neighbours must not wrap onto themselves
  return 1;
}

extern "C" int trvb_dmesh_create(trvb_ctx* ctx, trvb_comm* comm, trvb_dmesh** out) {
  TRVB_REQUIRE(ctx && comm && out, "trvb_dmesh_create: null argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_dmesh_create: root context only");
  const int R = trvb_comm_size(comm);
  TRVB_REQUIRE(trvb_dmesh_supported(ctx, R), "trvb_dmesh_create: %d x %d planes do not split "
               "over %d ranks", ctx->g.n[0], ctx->g.n[1], R);
  trvb_dmesh* dm = new trvb_dmesh();
  dm->ctx = ctx; dm->comm = comm; dm->R = R; dm->r = trvb_comm_rank(comm);
  dm->nx = ctx->g.n[0] / R; dm->nj = ctx->g.n[1] / R; dm->margin = DIST_MARGIN;
  const int nw[3] = {dm->nx + 2 * dm->margin, ctx->g.n[1], ctx->g.n[2]};
  const double Lw[3] = {ctx->g.dr[0] * nw[0], ctx->g.L[1], ctx->g.L[2]};
  int st = trvb_ctx_create(&dm->win, ctx->device, nw, Lw, ctx->g.order);
  if (st) { delete dm; return st; }
  // one stream for the whole phase
  trvb_arena_retire_stream(dm->win->device, dm->win->stream);
  cudaStreamDestroy(dm->win->stream);
  dm->win->stream = ctx->stream;
  dm->win->borrowed_stream = true;
  *out = dm;
  return 0;
}

extern "C" void trvb_dmesh_destroy(trvb_dmesh* dm) {
  if (!dm) return;
  cudaSetDevice(dm->ctx->device);
  cudaStreamSynchronize(dm->ctx->stream);
  if (dm->has_fwd) cufftDestroy(dm->plan_fwd);
  if (dm->has_inv) cufftDestroy(dm->plan_inv);
  if (dm->has_x) cufftDestroy(dm->plan_x);
  if (dm->U) trvb_dev_free_raw(dm->ctx, dm->U);
  if (dm->win) trvb_ctx_destroy(dm->win);
  delete dm;
}

extern "C" int trvb_dmesh_get(trvb_ctx* ctx, trvb_comm* comm, trvb_dmesh** out) {
  TRVB_REQUIRE(ctx && comm && out, "trvb_dmesh_get: null argument");
  trvb_dmesh* dm = ctx->dmesh;
  if (dm && (dm->comm != comm || dm->R != trvb_comm_size(comm) || dm->r != trvb_comm_rank(comm))) {
    trvb_dmesh_destroy(dm);
    ctx->dmesh = dm = nullptr;
  }
  if (!dm) {
    int st = trvb_dmesh_create(ctx, comm, &dm);
    if (st) return st;
    ctx->dmesh = dm;
  }
  *out = dm;
  return 0;
}

extern "C" int trvb_dmesh_planes(const trvb_dmesh* dm, int* x0, int* nx) {
  TRVB_REQUIRE(dm && x0 && nx, "trvb_dmesh_planes: null argument");
  *x0 = dm->r * dm->nx; *nx = dm->nx;
  return 0;
}

// delta n(k) of `n` unit-weight particles (device arrays holding the WHOLE catalogue on
// every rank), left distributed by rows of k_y inside `dm`; `k0_add` is added to the k = 0
// mode (mean subtraction, S/field.cpp:1229-1244).  Collective.
extern "C" int trvb_dmesh_density(trvb_dmesh* dm, long long n, const double* x, const double* y,
                                  const double* z, double k0_add) {
  TRVB_REQUIRE(dm && x && y && z && n > 0, "trvb_dmesh_density: bad argument");
  g_dmesh_density_calls++;
  trvb_ctx* ctx = dm->ctx; trvb_ctx* win = dm->win;
  const GridDesc& g = ctx->g;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  int st = ensure_plans(dm);
  if (st) return st;
  const int x0 = dm->r * dm->nx, M = dm->margin;
  StageTrace trace("density", ctx->stream, dm->r);
  // -- particles of the window --
  double* sel[3] = {nullptr, nullptr, nullptr};
  unsigned long long* d_count = nullptr;
  for (int a = 0; a < 3; a++) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&sel[a], sizeof(double) * (size_t)n));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_count, sizeof(unsigned long long)));
  TRVB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
  {
    const int blocks = (int)std::min<long long>(div_up(n, 1024), (long long)ctx->num_sms * 8);
    k_window_select<<<blocks, 256, 0, ctx->stream>>>(
      x, y, z, n, 1. / g.dr[0], g.n[0], x0 - 2, dm->nx + 4, (double)(x0 - M) * g.dr[0], g.L[0],
      sel[0], sel[1], sel[2], d_count);
    TRVB_LAUNCH_CHECK();
  }
  unsigned long long h_count = 0;
  TRVB_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  trvb_dev_free_raw(ctx, d_count);
  trace.mark("select");
  // -- assignment on the window, 2-D transforms of the slab's planes --
  trvb_mesh wmesh; wmesh.layout = TRVB_REAL; wmesh.k0_add = 0.; wmesh.data = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, &wmesh.data, trvb_mesh_bytes(win, TRVB_REAL)));
  if (h_count > 0) {
    trvb_cat* cat = nullptr;
    st = trvb_cat_create(win, &cat, (long long)h_count, sel[0], sel[1], sel[2], nullptr, nullptr, 2);
    if (st == 0) {
      st = trvb_assign(win, cat, TRVB_W_UNIT, 0, 0, 1., /*density_units=*/0, /*accumulate=*/0,
                       /*shifted=*/0, /*mode=*/0, wmesh);
      trvb_cat_destroy(cat);
    }
    if (st) return st;
  } else {
    TRVB_CUDA(cudaMemsetAsync(wmesh.data, 0, trvb_mesh_bytes(win, TRVB_REAL), ctx->stream));
  }
  for (int a = 0; a < 3; a++) trvb_dev_free_raw(ctx, sel[a]);
  trace.mark("sort+assign");
  const size_t slab_c = (size_t)dm->nx * g.n[1] * g.nh;   // complex elements of a slab
  double2* T = nullptr; double2* S = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&T, sizeof(double2) * slab_c));
  double* planes = (double*)wmesh.data + (size_t)M * g.n[1] * g.n[2];
  st = exec_with_area(ctx, dm->plan_fwd, dm->ws_fwd, run_d2z, planes, T, 0);
  if (st) return st;
  trvb_dev_free_raw(ctx, wmesh.data);
  trace.mark("fft2d");
  // -- rows of k_y to their owners, transform along x --
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&S, sizeof(double2) * slab_c));
  {
    const RowLaunch rl = row_launch(ctx->num_sms, dm->nx, g.n[1], g.nh);
    k_pack_rows<<<rl.grid, rl.block, 0, ctx->stream>>>(T, dm->nx, g.n[1], dm->nj, g.nh, S);
    TRVB_LAUNCH_CHECK();
  }
  trvb_dev_free_raw(ctx, T);
  trace.mark("pack");
  if (!dm->U) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&dm->U, sizeof(double2) * slab_c));
  st = trvb_comm_alltoall(ctx, dm->comm, (const double*)S, (double*)dm->U,
                          2LL * dm->nx * dm->nj * g.nh);
  if (st) return st;
  trvb_dev_free_raw(ctx, S);
  trace.mark("alltoall");
  st = exec_with_area(ctx, dm->plan_x, dm->ws_x, run_z2z, dm->U, dm->U, CUFFT_FORWARD);
  if (st) return st;
  trace.mark("fft_x");
  if (dm->r == 0 && k0_add != 0.) {
    k_add_first<<<1, 1, 0, ctx->stream>>>(dm->U, k0_add);
    TRVB_LAUNCH_CHECK();
  }
  return 0;
}

// The modes of the distributed mesh that the grid of `sub` represents, gathered onto every
// rank as a HALF mesh of `sub`'s extents (raw modes: readers divide by the window
// themselves, through KView's low-|k| storage).  Collective.
extern "C" int trvb_dmesh_gather_lowk(trvb_dmesh* dm, trvb_ctx* sub, trvb_mesh dst) {
  TRVB_REQUIRE(dm && sub && dst.data && dm->U, "trvb_dmesh_gather_lowk: bad argument");
  TRVB_REQUIRE(sub->parent == dm->ctx && dst.layout == TRVB_HALF,
               "trvb_dmesh_gather_lowk: `sub` must be a sub-grid of the mesh and dst HALF");
  trvb_ctx* ctx = dm->ctx;
  const GridDesc& g = ctx->g;
  const int s0 = sub->g.n[0], s1 = sub->g.n[1], s2 = sub->g.n[2], sh = sub->g.nh;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  StageTrace trace("gather_lowk", ctx->stream, dm->r);
  double2* Tsub = nullptr;
  const size_t total = (size_t)s0 * s1 * sh;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&Tsub, sizeof(double2) * total));
  {
    const RowLaunch rl = row_launch(ctx->num_sms, s1, s0, sh);
    k_lowk_rows<<<rl.grid, rl.block, 0, ctx->stream>>>(dm->U, g.n[0], g.n[1], g.nh, dm->r * dm->nj,
                                                       dm->nj, s0, s1, s2, dm->r, dm->nj, Tsub);
    TRVB_LAUNCH_CHECK();
  }
  trace.mark("rows");
  // runs of consecutive rows with one owner
  std::vector<int> root; std::vector<long long> off, cnt;
  const long long row = 2LL * s0 * sh;   // doubles per k_y row
  auto owner_of = [&](int js) {
    const int mj = signed_index(js, s1);
    if (2 * std::abs(mj) >= s1) return 0;
    const int j = mj >= 0 ? mj : mj + g.n[1];
    return j / dm->nj;
  };
  for (int js = 0; js < s1;) {
    const int o = owner_of(js);
    int je = js + 1;
    while (je < s1 && owner_of(je) == o) je++;
    root.push_back(o); off.push_back(row * js); cnt.push_back(row * (je - js));
    js = je;
  }
  int st = trvb_comm_bcast_segments(ctx, dm->comm, (double*)Tsub, (int)root.size(), root.data(),
                                    off.data(), cnt.data());
  if (st) return st;
  trace.mark("bcast");
  {
    const RowLaunch rl = row_launch(ctx->num_sms, s0, s1, sh);
    k_lowk_transpose<<<rl.grid, rl.block, 0, ctx->stream>>>(Tsub, s0, s1, sh, (double2*)dst.data);
    TRVB_LAUNCH_CHECK();
  }
  trvb_dev_free_raw(ctx, Tsub);
  trace.mark("transpose");
  ctx->lowk_ptr = dst.data;
  for (int a = 0; a < 3; a++) ctx->lowk_dims[a] = sub->g.n[a];
  return 0;
}

extern "C" void trvb_dmesh_forget_lowk(trvb_dmesh* dm) {
  if (dm && dm->ctx) dm->ctx->lowk_ptr = nullptr;
}

// xi(r) = IFFT[(fa conj(fb) / C1 - S) / V] on this rank's planes (REAL, [nx][n1][n2] at
// device address `xi_planes`), for fa = delta n + add_a delta_k0 and fb = delta n + add_b
// delta_k0 of the distributed mesh (trvb_shot_xi, S/field.cpp:3273-3345).  Collective.
extern "C" int trvb_dmesh_shot_xi(trvb_dmesh* dm, double add_a, double add_b, const double S[2],
                                  double* xi_planes) {
  TRVB_REQUIRE(dm && S && xi_planes && dm->U, "trvb_dmesh_shot_xi: bad argument");
  TRVB_REQUIRE(S[1] == 0., "trvb_dmesh_shot_xi: real shot-noise amplitude required");
  trvb_ctx* ctx = dm->ctx;
  const GridDesc& g = ctx->g;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  int st = ensure_plans(dm);
  if (st) return st;
  const size_t slab_c = (size_t)dm->nx * g.n[1] * g.nh;
  StageTrace trace("shot_xi", ctx->stream, dm->r);
  double2* W = nullptr; double2* Rb = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&W, sizeof(double2) * slab_c));
  {
    const RowLaunch rl = row_launch(ctx->num_sms, g.n[0], dm->nj, g.nh);
    k_shot_spectrum_rows<<<rl.grid, rl.block, 0, ctx->stream>>>(
      dm->U, g, tables_of(ctx), dm->r * dm->nj, dm->nj, add_a, add_b, S[0], S[1], W);
    TRVB_LAUNCH_CHECK();
  }
  trace.mark("spectrum");
  st = exec_with_area(ctx, dm->plan_x, dm->ws_x, run_z2z, W, W, CUFFT_INVERSE);
  if (st) return st;
  trace.mark("fft_x");
  // planes [q nx, (q + 1) nx) of W are contiguous: no packing on this side
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&Rb, sizeof(double2) * slab_c));
  st = trvb_comm_alltoall(ctx, dm->comm, (const double*)W, (double*)Rb, 2LL * dm->nx * dm->nj * g.nh);
  if (st) return st;
  trace.mark("alltoall");
  {
    const RowLaunch rl = row_launch(ctx->num_sms, dm->nx, g.n[1], g.nh);
    k_unpack_rows<<<rl.grid, rl.block, 0, ctx->stream>>>(Rb, dm->nx, g.n[1], dm->nj, g.nh, W);
    TRVB_LAUNCH_CHECK();
  }
  trvb_dev_free_raw(ctx, Rb);
  trace.mark("unpack");
  st = exec_with_area(ctx, dm->plan_inv, dm->ws_inv, run_z2d, W, xi_planes, 0);
  trvb_dev_free_raw(ctx, W);
  trace.mark("fft2d");
  return st;
}
