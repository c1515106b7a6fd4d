// trvb_comm.cu -- the one exchange step of the multi-GPU path: an all-reduce(sum) of the
// partial data vectors over NCCL (NVLink 5 / NVSwitch), behind the C ABI
// (include/trvb.h: trvb_comm_*, trvb_allreduce).
//
// The reference's multi-GPU mode is single-process cuFFT-Xt (S/field.cpp:212-235, devices
// from S/monitor.cpp:258-324); here the mesh state is replicated, the (k1, k2) bin pairs
// are dealt to the GPUs and every entry is produced by exactly one of them, so that the
// sum over ranks only adds zeros.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries a
// copy -- e.g. the one bundled with PyTorch -- reuses it instead of loading a second one,
// and single-GPU users need no NCCL at all.
#include "trvb_common.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>

namespace {

// The slice of nccl.h this file needs (NCCL 2.x ABI).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };   // ncclDataType_t
enum { ncclSum = 0 };       // ncclRedOp_t

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.handle) return 0;
  // Order matters: a process must end up with ONE libnccl.so.2 (the dynamic loader
  // resolves every later request for that soname -- e.g. PyTorch's -- to the copy that is
  // already mapped).  1. a copy already loaded by the process; 2. the copy named by
  // TRV_NCCL_LIB (the Python shim points it at the one PyTorch bundles, which is the newest
  // in this image); 3. the system library.
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    const char* env = getenv("TRV_NCCL_LIB");
    if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (h) break;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) {
    trvb_set_error("NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
    return 4;
  }
  NcclApi api;
  api.handle = h;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.Send = (decltype(api.Send))dlsym(h, "ncclSend");
  api.Recv = (decltype(api.Recv))dlsym(h, "ncclRecv");
  api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
  api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce
      || !api.GroupStart || !api.GroupEnd || !api.GetErrorString || !api.Send || !api.Recv
      || !api.Broadcast) {
    trvb_set_error("libnccl.so.2 lacks an expected symbol");
    dlclose(h);
    return 4;
  }
  g_nccl = api;
  return 0;
}

#define TRVB_NCCL(call)                                                      \
  do {                                                                       \
    ncclResult_t r__ = (call);                                               \
    if (r__ != ncclSuccess) {                                                \
      trvb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,           \
                     g_nccl.GetErrorString(r__));                            \
      return 2000 + (int)r__;                                                \
    }                                                                        \
  } while (0)

}  // namespace

struct trvb_comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  int device = 0;
  double* d_buf = nullptr;       // staging for host-buffer reductions
  size_t buf_doubles = 0;
};

extern "C" int trvb_nccl_version(void) {
  if (load_nccl()) return 0;
  int v = 0;
  if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
  return v;
}

extern "C" int trvb_comm_unique_id(char id[128]) {
  TRVB_REQUIRE(id != nullptr, "trvb_comm_unique_id: null argument");
  int st = load_nccl();
  if (st) return st;
  ncclUniqueId uid;
  TRVB_NCCL(g_nccl.GetUniqueId(&uid));
  std::memcpy(id, uid.internal, 128);
  return 0;
}

extern "C" int trvb_comm_create(trvb_comm** out, int device, int nranks, int rank,
                                const char id[128]) {
  TRVB_REQUIRE(out && id && nranks >= 1 && rank >= 0 && rank < nranks,
               "trvb_comm_create: bad argument (rank %d of %d)", rank, nranks);
  int st = load_nccl();
  if (st) return st;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
  TRVB_CUDA(cudaSetDevice(device));
  ncclUniqueId uid;
  std::memcpy(uid.internal, id, 128);
  trvb_comm* c = new trvb_comm();
  c->nranks = nranks; c->rank = rank; c->device = device;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, uid, rank);
  if (prev >= 0) cudaSetDevice(prev);
  if (r != ncclSuccess) {
    trvb_set_error("ncclCommInitRank(rank %d of %d, device %d) -> %s", rank, nranks, device,
                   g_nccl.GetErrorString(r));
    delete c;
    return 2000 + (int)r;
  }
  *out = c;
  return 0;
}

extern "C" void trvb_comm_destroy(trvb_comm* c) {
  if (!c) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
  cudaSetDevice(c->device);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
}

extern "C" int trvb_comm_size(const trvb_comm* c) { return c ? c->nranks : 1; }
extern "C" int trvb_comm_rank(const trvb_comm* c) { return c ? c->rank : 0; }

// In-place sum over the ranks of `n` doubles in DEVICE memory, enqueued on the context's
// stream (no host synchronisation).
extern "C" int trvb_allreduce_device(trvb_ctx* ctx, trvb_comm* c, double* dbuf, long long n) {
  TRVB_REQUIRE(ctx && c && dbuf && n >= 0, "trvb_allreduce_device: bad argument");
  TRVB_REQUIRE(ctx->device == c->device, "trvb_allreduce_device: context on device %d, "
               "communicator on device %d", ctx->device, c->device);
  if (n == 0) return 0;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  TRVB_NCCL(g_nccl.AllReduce(dbuf, dbuf, (size_t)n, ncclFloat64, ncclSum, c->comm, ctx->stream));
  return 0;
}

// In-place sum over the ranks of `n` doubles in HOST memory (the result vectors of an
// estimator call, <= 26 KB): staged through a device buffer; returns after the sum is back.
extern "C" int trvb_allreduce(trvb_ctx* ctx, trvb_comm* c, double* buf, long long n) {
  TRVB_REQUIRE(ctx && c && buf && n >= 0, "trvb_allreduce: bad argument");
  if (n == 0) return 0;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  if (c->buf_doubles < (size_t)n) {
    if (c->d_buf) { TRVB_CUDA(cudaStreamSynchronize(ctx->stream)); TRVB_CUDA(cudaFree(c->d_buf)); }
    const size_t want = ((size_t)n + 1023) / 1024 * 1024;
    TRVB_CUDA(cudaMalloc(&c->d_buf, sizeof(double) * want));
    c->buf_doubles = want;
  }
  TRVB_CUDA(cudaMemcpyAsync(c->d_buf, buf, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice,
                            ctx->stream));
  int st = trvb_allreduce_device(ctx, c, c->d_buf, n);
  if (st) return st;
  TRVB_CUDA(cudaMemcpyAsync(buf, c->d_buf, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost,
                            ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Personalised exchange on the context's stream (no host synchronisation): block q of
// `send` (`n` doubles each) goes to rank q, block q of `recv` comes from rank q.  The one
// data-path exchange of the distributed mesh phase (x-slab <-> k_y-slab transposition of
// the 3-D FFT); over NVSwitch every pair of GPUs has its own full-bandwidth path.
extern "C" int trvb_comm_alltoall(trvb_ctx* ctx, trvb_comm* c, const double* send, double* recv,
                                  long long n) {
  TRVB_REQUIRE(ctx && c && send && recv && n >= 0, "trvb_comm_alltoall: bad argument");
  TRVB_REQUIRE(ctx->device == c->device, "trvb_comm_alltoall: context on device %d, "
               "communicator on device %d", ctx->device, c->device);
  if (n == 0) return 0;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  TRVB_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < c->nranks; q++) {
    ncclResult_t r1 = g_nccl.Send(send + (size_t)q * n, (size_t)n, ncclFloat64, q, c->comm, ctx->stream);
    ncclResult_t r2 = g_nccl.Recv(recv + (size_t)q * n, (size_t)n, ncclFloat64, q, c->comm, ctx->stream);
    if (r1 != ncclSuccess || r2 != ncclSuccess) {
      g_nccl.GroupEnd();
      trvb_set_error("trvb_comm_alltoall: ncclSend/ncclRecv with rank %d -> %s", q,
                     g_nccl.GetErrorString(r1 != ncclSuccess ? r1 : r2));
      return 2000 + (int)(r1 != ncclSuccess ? r1 : r2);
    }
  }
  TRVB_NCCL(g_nccl.GroupEnd());
  return 0;
}

// Segment s of the device buffer `buf` (offset[s], count[s] doubles) is held by rank root[s]
// and wanted by all, on the context's stream.  Small gathers (latency-bound: C2's 24 MB over
// 8 GPUs 0.26 -> 0.18 ms) go as ONE group of point-to-point transfers -- a single fused NCCL
// launch, where a broadcast per segment costs a launch each; large ones (egress-bound: an
// owner would send its rows R - 1 times, C5's 1.26 GB 2.1 -> 3.4 ms) as grouped broadcasts.
extern "C" int trvb_comm_bcast_segments(trvb_ctx* ctx, trvb_comm* c, double* buf, int nseg,
                                        const int* root, const long long* offset,
                                        const long long* count) {
  TRVB_REQUIRE(ctx && c && buf && nseg >= 0 && (nseg == 0 || (root && offset && count)),
               "trvb_comm_bcast_segments: bad argument");
  TRVB_REQUIRE(ctx->device == c->device, "trvb_comm_bcast_segments: context on device %d, "
               "communicator on device %d", ctx->device, c->device);
  if (nseg == 0 || c->nranks == 1) return 0;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  long long total = 0;
  for (int s = 0; s < nseg; s++) total += std::max(0LL, count[s]);
  const bool point_to_point = total <= (8LL << 20);   // <= 64 MB
  TRVB_NCCL(g_nccl.GroupStart());
  ncclResult_t bad = ncclSuccess;
  for (int s = 0; s < nseg && bad == ncclSuccess; s++) {
    if (count[s] <= 0) continue;
    if (!point_to_point) {
      bad = g_nccl.Broadcast(buf + offset[s], buf + offset[s], (size_t)count[s], ncclFloat64,
                             root[s], c->comm, ctx->stream);
    } else if (root[s] == c->rank) {
      for (int q = 0; q < c->nranks && bad == ncclSuccess; q++) {
        if (q == c->rank) continue;
        bad = g_nccl.Send(buf + offset[s], (size_t)count[s], ncclFloat64, q, c->comm, ctx->stream);
      }
    } else {
      bad = g_nccl.Recv(buf + offset[s], (size_t)count[s], ncclFloat64, root[s], c->comm, ctx->stream);
    }
  }
  if (bad != ncclSuccess) {
    g_nccl.GroupEnd();
    trvb_set_error("trvb_comm_bcast_segments: NCCL -> %s", g_nccl.GetErrorString(bad));
    return 2000 + (int)bad;
  }
  TRVB_NCCL(g_nccl.GroupEnd());
  return 0;
}
