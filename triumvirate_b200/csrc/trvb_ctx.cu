// trvb_ctx.cu -- context, memory, tables and cuFFT plan cache of libtrvb.so.
#include "trvb_common.cuh"

#include <cstring>
#include <mutex>
#include <unordered_map>

static thread_local std::string g_err;
// relaxed atomics: several host threads launch in single-process multi-GPU mode
std::atomic<long long> g_trvb_launches{0};
std::atomic<long long> g_trvb_fft_execs{0};
static std::atomic<long long> g_arena_mallocs{0};

void trvb_set_error(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

extern "C" const char* trvb_last_error(void) { return g_err.c_str(); }
extern "C" const char* trvb_version(void) { return "triumvirate_b200 0.1 (sm_100a)"; }
extern "C" int trvb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" int trvb_current_device(void) {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
  return d;
}
extern "C" int trvb_set_current_device(int device) {
  TRVB_CUDA(cudaSetDevice(device));
  return 0;
}
extern "C" long long trvb_launch_count(void) { return g_trvb_launches; }
extern "C" void trvb_launch_count_reset(void) { g_trvb_launches = 0; g_trvb_fft_execs = 0; }
extern "C" long long trvb_fft_exec_count(void) { return g_trvb_fft_execs; }
extern "C" long long trvb_arena_malloc_count(void) { return g_arena_mallocs; }

// ---------------------------------------------------------------------
// Caching arena.
// ---------------------------------------------------------------------
namespace {

struct ArenaBlock { void* p; size_t bytes; cudaStream_t stream; };
struct Arena {
  std::mutex mu;
  std::multimap<size_t, ArenaBlock> free_blocks;          // by size
  std::unordered_map<void*, size_t> live;                 // handed out
  size_t cached = 0;
  // Driver-side free memory as last measured, kept current by subtracting the
  // arena's own cudaMallocs: cudaMemGetInfo costs ~1 ms and was seen to stall
  // for up to 100 ms on B200 hosts, so it is not called per estimator step.
  bool free_known = false;
  size_t driver_free = 0, driver_total = 0;
};
Arena g_arena[64];
const cudaStream_t kRetired = (cudaStream_t)(-1);   // owner stream drained and destroyed

size_t arena_round(size_t bytes) {
  const size_t q = bytes < (1u << 20) ? 512 : (size_t)(2u << 20);
  return ((bytes ? bytes : 8) + q - 1) / q * q;
}

}  // namespace

cudaError_t trvb_arena_alloc(int device, cudaStream_t stream, void** p, size_t bytes) {
  Arena& a = g_arena[device & 63];
  const size_t want = arena_round(bytes);
  {
    std::lock_guard<std::mutex> lock(a.mu);
    // Exact size classes only: the estimator asks for a handful of distinct
    // sizes, and a near fit would take the block the next request needs.
    auto it = a.free_blocks.find(want);
    if (it != a.free_blocks.end()) {
      ArenaBlock b = it->second;
      a.free_blocks.erase(it);
      a.cached -= b.bytes;
      if (b.stream != stream && b.stream != kRetired) {
        cudaError_t e = cudaStreamSynchronize(b.stream);
        if (e != cudaSuccess) return e;
      }
      a.live[b.p] = b.bytes;
      *p = b.p;
      return cudaSuccess;
    }
  }
  g_arena_mallocs++;
  cudaError_t e = cudaMalloc(p, want);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    trvb_arena_trim(device);      // give cached blocks back and retry once
    e = cudaMalloc(p, want);
  }
  std::lock_guard<std::mutex> lock(a.mu);
  if (e != cudaSuccess) { a.free_known = false; return e; }
  a.live[*p] = want;
  if (a.free_known) a.driver_free = a.driver_free > want ? a.driver_free - want : 0;
  return cudaSuccess;
}

cudaError_t trvb_arena_free(int device, cudaStream_t stream, void* p) {
  if (!p) return cudaSuccess;
  Arena& a = g_arena[device & 63];
  std::lock_guard<std::mutex> lock(a.mu);
  auto it = a.live.find(p);
  if (it == a.live.end()) return cudaErrorInvalidValue;
  ArenaBlock b{p, it->second, stream};
  a.live.erase(it);
  a.free_blocks.emplace(b.bytes, b);
  a.cached += b.bytes;
  return cudaSuccess;
}

void trvb_arena_retire_stream(int device, cudaStream_t stream) {
  Arena& a = g_arena[device & 63];
  std::lock_guard<std::mutex> lock(a.mu);
  for (auto& kv : a.free_blocks) if (kv.second.stream == stream) kv.second.stream = kRetired;
}

size_t trvb_arena_cached_bytes(int device) {
  Arena& a = g_arena[device & 63];
  std::lock_guard<std::mutex> lock(a.mu);
  return a.cached;
}

void trvb_arena_trim(int device) {
  Arena& a = g_arena[device & 63];
  std::multimap<size_t, ArenaBlock> blocks;
  {
    std::lock_guard<std::mutex> lock(a.mu);
    blocks.swap(a.free_blocks);
    a.cached = 0;
    a.free_known = false;
  }
  if (blocks.empty()) return;
  cudaDeviceSynchronize();
  for (auto& kv : blocks) cudaFree(kv.second.p);
}

extern "C" void trvb_mem_info_invalidate(int device);
extern "C" void trvb_arena_release(int device) {
  if (trvb_arena_cached_bytes(device) == 0) return;   // never touches a GPU it did not use
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
  if (cudaSetDevice(device) == cudaSuccess) trvb_arena_trim(device); else cudaGetLastError();
  if (prev >= 0) cudaSetDevice(prev);
}

int trvb_scratch(trvb_ctx* ctx, size_t bytes, double** out) {
  if (ctx->scratch_bytes < bytes) {
    if (ctx->d_scratch) {
      TRVB_CUDA(trvb_dev_free_raw(ctx, ctx->d_scratch));
      ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    }
    size_t want = bytes < (size_t)(1 << 20) ? (size_t)(1 << 20) : bytes;
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&ctx->d_scratch, want));
    ctx->scratch_bytes = want;
  }
  *out = ctx->d_scratch;
  return 0;
}

// Per-axis correction tables, evaluated on the host with the reference's
// expressions so the device only multiplies three table entries.
//   sinc:  S/field.cpp:1136-1145   u = M_PI * m / double(n); sin(u)/u (1 at m=0)
//   alias: S/field.cpp:3445-3502   s2 = sin(u)^2 (0 at m = 0);
//          ngp 1; cic 1 - 2/3 s2; tsc 1 - s2 + 2/15 s2^2;
//          pcs 1 - 4/3 s2 + 2/5 s2^2 - 4/315 s2^3
static int build_tables(trvb_ctx* ctx) {
  for (int ax = 0; ax < 3; ax++) {
    const int n = ctx->g.n[ax];
    std::vector<double> sinc(n), alias(n), ralias(n);
    for (int i = 0; i < n; i++) {
      int m = signed_index(i, n);
      double u = M_PI * m / double(n);
      sinc[i] = (m != 0) ? std::sin(u) / u : 1.;
      double s2 = (m != 0) ? std::sin(u) * std::sin(u) : 0.;
      double a = 1.;
      switch (ctx->g.order) {
        case 1: a = 1.; break;
        case 2: a = (1. - 2./3. * s2); break;
        case 3: a = (1. - s2 + 2./15. * s2 * s2); break;
        case 4: a = (1. - 4./3. * s2 + 2./5. * s2 * s2 - 4./315. * s2 * s2 * s2); break;
      }
      alias[i] = a;
      ralias[i] = 1. / a;
    }
    TRVB_CUDA(cudaMalloc(&ctx->d_sinc[ax], sizeof(double) * n));
    TRVB_CUDA(cudaMalloc(&ctx->d_alias[ax], sizeof(double) * n));
    TRVB_CUDA(cudaMemcpyAsync(ctx->d_sinc[ax], sinc.data(), sizeof(double) * n,
                              cudaMemcpyHostToDevice, ctx->stream));
    TRVB_CUDA(cudaMemcpyAsync(ctx->d_alias[ax], alias.data(), sizeof(double) * n,
                              cudaMemcpyHostToDevice, ctx->stream));
    TRVB_CUDA(cudaMalloc(&ctx->d_ralias[ax], sizeof(double) * n));
    TRVB_CUDA(cudaMemcpyAsync(ctx->d_ralias[ax], ralias.data(), sizeof(double) * n,
                              cudaMemcpyHostToDevice, ctx->stream));
  }
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static void fill_grid(GridDesc& g, const int n[3], const double L[3], int order) {
  for (int i = 0; i < 3; i++) {
    g.n[i] = n[i]; g.L[i] = L[i];
    g.dr[i] = L[i] / n[i];
    g.dk[i] = 2. * M_PI / L[i];
  }
  g.nh = n[2] / 2 + 1;
  g.nmesh = (long long)n[0] * n[1] * n[2];
  g.vol = L[0] * L[1] * L[2];
  g.vol_cell = g.vol / double(g.nmesh);
  g.order = order;
}

extern "C" int trvb_ctx_create(trvb_ctx** out, int device, const int ngrid[3],
                               const double boxsize[3], int assignment_order) {
  TRVB_REQUIRE(out && ngrid && boxsize, "trvb_ctx_create: null argument");
  TRVB_REQUIRE(assignment_order >= 1 && assignment_order <= 4,
               "trvb_ctx_create: assignment order %d not in 1..4", assignment_order);
  for (int i = 0; i < 3; i++) {
    TRVB_REQUIRE(ngrid[i] > 0 && boxsize[i] > 0.,
                 "trvb_ctx_create: non-positive grid/box along axis %d", i);
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    trvb_set_error("trvb_ctx_create: no CUDA device available (%s); this "
                   "library has no CPU fallback", cudaGetErrorString(e));
    return 3;
  }
  TRVB_REQUIRE(device >= 0 && device < ndev, "trvb_ctx_create: device %d of %d",
               device, ndev);
  // The caller's current device is put back on every return path.
  struct DeviceRestore {
    int prev = -1;
    DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  TRVB_CUDA(cudaSetDevice(device));
  trvb_ctx* ctx = new trvb_ctx();
  ctx->device = device;
  fill_grid(ctx->g, ngrid, boxsize, assignment_order);
  TRVB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  TRVB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  int st = build_tables(ctx);
  if (st) { delete ctx; return st; }
  trvb_mem_info_invalidate(device);   // the next trvb_mem_info measures afresh
  *out = ctx;
  return 0;
}

extern "C" int trvb_subgrid_create(trvb_ctx* parent, trvb_ctx** out,
                                   const int nsub[3]) {
  TRVB_REQUIRE(parent && out && nsub, "trvb_subgrid_create: null argument");
  TRVB_REQUIRE(parent->parent == nullptr,
               "trvb_subgrid_create: parent must be a root context");
  for (int i = 0; i < 3; i++) {
    TRVB_REQUIRE(nsub[i] > 0 && nsub[i] <= parent->g.n[i],
                 "trvb_subgrid_create: nsub[%d]=%d outside (0, %d]", i, nsub[i],
                 parent->g.n[i]);
  }
  const std::vector<int> key(nsub, nsub + 3);
  auto hit = parent->subgrids.find(key);
  if (hit != parent->subgrids.end()) { *out = hit->second; return 0; }
  TRVB_CUDA(cudaSetDevice(parent->device));
  trvb_ctx* ctx = new trvb_ctx();
  ctx->device = parent->device;
  ctx->parent = parent;
  fill_grid(ctx->g, nsub, parent->g.L, parent->g.order);
  // TRV_OVERLAP=1 gives the sub-grid its own (highest-priority) stream so that the pair
  // branch can run beside the shot-noise branch of the parent grid (trvb_ctx_fork /
  // trvb_ctx_join order the two).  Measured on B200 it gains nothing on C2 or C5 (7.61 vs
  // 7.57 ms; both branches are bandwidth-bound and the timeline has no gaps to fill), so
  // by default everything stays on the parent's stream.
  const char* want_overlap = getenv("TRV_OVERLAP");
  const bool no_overlap_flag = !(want_overlap != nullptr && want_overlap[0] == '1');
  if (no_overlap_flag) {
    ctx->stream = parent->stream;
  } else {
    // Highest priority: the block scheduler hands free SM slots to the sub-grid's small
    // kernels first, so they interleave with the parent grid's long bandwidth-bound
    // kernels instead of queueing behind their thousands of blocks.
    int prio_lo = 0, prio_hi = 0;
    TRVB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TRVB_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));
    TRVB_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    ctx->own_stream = true;
  }
  ctx->num_sms = parent->num_sms;
  parent->subgrids[key] = ctx;
  *out = ctx;
  return 0;
}

static void destroy_ctx_now(trvb_ctx* ctx);

extern "C" void trvb_ctx_destroy(trvb_ctx* ctx) {
  if (!ctx) return;
  if (ctx->parent) return;   // sub-grid contexts live and die with their parent
  for (auto& kv : ctx->subgrids) destroy_ctx_now(kv.second);
  ctx->subgrids.clear();
  destroy_ctx_now(ctx);
}

static void destroy_ctx_now(trvb_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->dmesh) { trvb_dmesh_destroy(ctx->dmesh); ctx->dmesh = nullptr; }
  if (ctx->has_z2z) cufftDestroy(ctx->plan_z2z);
  if (ctx->has_d2z) cufftDestroy(ctx->plan_d2z);
  if (ctx->has_z2d) cufftDestroy(ctx->plan_z2d);
  for (auto& kv : ctx->batch_plans) cufftDestroy(kv.second);
  for (auto& kv : ctx->line_plans) cufftDestroy(kv.second);
  for (int ax = 0; ax < 3; ax++) {
    if (ctx->d_sinc[ax]) cudaFree(ctx->d_sinc[ax]);
    if (ctx->d_alias[ax]) cudaFree(ctx->d_alias[ax]);
    if (ctx->d_ralias[ax]) cudaFree(ctx->d_ralias[ax]);
  }
  for (auto& kv : ctx->sjl) {
    if (kv.second.d_y) cudaFree(kv.second.d_y);
    if (kv.second.d_c) cudaFree(kv.second.d_c);
  }
  if (ctx->d_scratch) trvb_arena_free(ctx->device, ctx->stream, ctx->d_scratch);
  if ((!ctx->parent || ctx->own_stream) && ctx->stream && !ctx->borrowed_stream) {
    trvb_arena_retire_stream(ctx->device, ctx->stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->own_stream) cudaEventDestroy(ctx->fork_event);
  }
  delete ctx;
}

extern "C" int trvb_ctx_fork(trvb_ctx* parent, trvb_ctx* sub) {
  TRVB_REQUIRE(parent && sub, "trvb_ctx_fork: null context");
  if (sub->stream == parent->stream) return 0;
  TRVB_CUDA(cudaEventRecord(sub->fork_event, parent->stream));
  TRVB_CUDA(cudaStreamWaitEvent(sub->stream, sub->fork_event, 0));
  return 0;
}

extern "C" int trvb_ctx_join(trvb_ctx* parent, trvb_ctx* sub) {
  TRVB_REQUIRE(parent && sub, "trvb_ctx_join: null context");
  if (sub->stream == parent->stream) return 0;
  TRVB_CUDA(cudaEventRecord(sub->fork_event, sub->stream));
  TRVB_CUDA(cudaStreamWaitEvent(parent->stream, sub->fork_event, 0));
  return 0;
}

extern "C" int trvb_ctx_set_deterministic(trvb_ctx* ctx, int on) {
  TRVB_REQUIRE(ctx != nullptr, "trvb_ctx_set_deterministic: null context");
  ctx->deterministic = on ? 1 : 0;
  return 0;
}

extern "C" int trvb_ctx_sync(trvb_ctx* ctx) {
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" void* trvb_ctx_stream(trvb_ctx* ctx) { return (void*)ctx->stream; }
extern "C" int trvb_ctx_device(const trvb_ctx* ctx) { return ctx ? ctx->device : -1; }
extern "C" long long trvb_ctx_nmesh(const trvb_ctx* ctx) { return ctx->g.nmesh; }

extern "C" size_t trvb_mesh_bytes(const trvb_ctx* ctx, int layout) {
  const GridDesc& g = ctx->g;
  switch (layout) {
    case TRVB_REAL: return sizeof(double) * (size_t)g.nmesh;
    case TRVB_COMPLEX: return 2 * sizeof(double) * (size_t)g.nmesh;
    case TRVB_HALF: return 2 * sizeof(double) * (size_t)g.n[0] * g.n[1] * g.nh;
  }
  return 0;
}

extern "C" int trvb_mem_info(trvb_ctx* ctx, size_t* free_bytes, size_t* total_bytes) {
  TRVB_CUDA(cudaSetDevice(ctx->device));
  Arena& a = g_arena[ctx->device & 63];
  std::lock_guard<std::mutex> lock(a.mu);
  if (!a.free_known) {
    TRVB_CUDA(cudaMemGetInfo(&a.driver_free, &a.driver_total));
    a.free_known = true;
  }
  // Blocks cached by the arena are available to this library.
  *free_bytes = a.driver_free + a.cached;
  *total_bytes = a.driver_total;
  return 0;
}

extern "C" void trvb_mem_info_invalidate(int device) {
  Arena& a = g_arena[device & 63];
  std::lock_guard<std::mutex> lock(a.mu);
  a.free_known = false;
}

extern "C" int trvb_malloc(trvb_ctx* ctx, void** dptr, size_t bytes) {
  TRVB_CUDA(cudaSetDevice(ctx->device));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, dptr, bytes));
  return 0;
}

extern "C" int trvb_free(trvb_ctx* ctx, void* dptr) {
  if (!dptr) return 0;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  // The arena orders reuse by the stream of the freeing context.  With TRV_OVERLAP=1 a
  // sub-grid context runs on a stream of its own while its meshes are owned (and freed) by
  // the root context: work still queued on that stream must finish before the block can
  // be handed to the root's stream.  (Opt-in path; by default there is one stream.)
  trvb_ctx* root = ctx->parent ? ctx->parent : ctx;
  for (auto& kv : root->subgrids) {
    if (kv.second->own_stream && kv.second->stream != ctx->stream) {
      TRVB_CUDA(cudaStreamSynchronize(kv.second->stream));
    }
  }
  if (ctx->parent && ctx->own_stream) TRVB_CUDA(cudaStreamSynchronize(root->stream));
  TRVB_CUDA(trvb_dev_free_raw(ctx, dptr));   // stream-ordered: no synchronisation otherwise
  return 0;
}

extern "C" int trvb_memset0(trvb_ctx* ctx, void* dptr, size_t bytes) {
  TRVB_CUDA(cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
  return 0;
}

extern "C" int trvb_h2d(trvb_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  TRVB_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int trvb_d2h(trvb_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  TRVB_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int trvb_d2d(trvb_ctx* ctx, void* dst, const void* src, size_t bytes) {
  TRVB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

extern "C" int trvb_sjl_table(trvb_ctx* ctx, int ell, const double* y,
                              const double* c, int nsample, double step) {
  TRVB_REQUIRE(ctx && y && c && nsample >= 2 && step > 0.,
               "trvb_sjl_table: bad argument");
  trvb_ctx* root = ctx->parent ? ctx->parent : ctx;
  SjlTable& t = root->sjl[ell];
  if (t.d_y && t.nsample == nsample && t.step == step
      && std::memcmp(t.h_y.data(), y, sizeof(double) * nsample) == 0
      && std::memcmp(t.h_c.data(), c, sizeof(double) * nsample) == 0) return 0;
  if (t.d_y) { TRVB_CUDA(cudaStreamSynchronize(root->stream)); cudaFree(t.d_y); cudaFree(t.d_c); }
  TRVB_CUDA(cudaMalloc(&t.d_y, sizeof(double) * nsample));
  TRVB_CUDA(cudaMalloc(&t.d_c, sizeof(double) * nsample));
  TRVB_CUDA(cudaMemcpyAsync(t.d_y, y, sizeof(double) * nsample,
                            cudaMemcpyHostToDevice, root->stream));
  TRVB_CUDA(cudaMemcpyAsync(t.d_c, c, sizeof(double) * nsample,
                            cudaMemcpyHostToDevice, root->stream));
  TRVB_CUDA(cudaStreamSynchronize(root->stream));
  t.nsample = nsample; t.step = step;
  t.h_y.assign(y, y + nsample); t.h_c.assign(c, c + nsample);
  return 0;
}
