// trvb_xpass.cu -- periodic-box mesh phase on one GPU with the x passes of both full-grid
// transforms fused around the shot-noise spectrum (include/trvb.h: trvb_box_fields_fused).
//
//   delta n(x) REAL --2-D D2Z of the n0 planes (cuFFT)--> T[n0][n1][nh]
//     --k_xpass_fused: FFT along x, low-|k| modes out, (fa conj fb / C1 - S) / V, inverse
//       FFT along x, in place (csrc/trvb_xpass.cuh)-->
//   T --2-D Z2D of the planes (cuFFT)--> xi(r) REAL
//
// It replaces, for the box bispectrum in throughput mode, cuFFT's 3-D D2Z
// (S/field.cpp:1496-1655), k_shot_spectrum (S/field.cpp:3273-3298) and cuFFT's 3-D Z2D
// (S/field.cpp:3318-3345): the half spectrum is read once and written once between the
// two 2-D transforms (2 x 1.08 GB at 512^3) instead of three times each way, and the full
// delta n(k) -- which no later step reads -- is never stored.
#include "trvb_common.cuh"
#include "trvb_xpass.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <vector>

namespace {

template <int N, int CK, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_xpass_fused(double2* __restrict__ T, xpass::Pointwise pw, const double2* __restrict__ tw_g) {
  using namespace xpass;
  extern __shared__ __align__(16) unsigned char xp_smem[];
  double2* tile = reinterpret_cast<double2*>(xp_smem);
  double2* tw = tile + N * CK;
  const int tid = threadIdx.x;
  for (int t = tid; t < N; t += NT) tw[t] = tw_g[t];
  const Column col = column_of(pw, (long long)blockIdx.x * CK + tid % CK);
  __syncthreads();
  constexpr int NS = Radix<N>::NS;
  stage_first<N, CK, NT>(tid, T, pw.ncols, col, tile, tw);
  __syncthreads();
  if constexpr (NS >= 3) { stage_fwd<N, CK, NT, 1>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 4) { stage_fwd<N, CK, NT, 2>(tid, tile, tw); __syncthreads(); }
  stage_junction<N, CK, NT>(tid, tile, pw, col);
  __syncthreads();
  if constexpr (NS >= 4) { stage_inv<N, CK, NT, 2>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 3) { stage_inv<N, CK, NT, 1>(tid, tile, tw); __syncthreads(); }
  stage_last<N, CK, NT>(tid, tile, tw, col, T, pw.ncols);
}

// The same pass with the tile brought in by cp.async: no thread waits on a global load with
// its registers tied up, and a CTA has its whole tile in flight from its first instruction
// (the register-staged first stage had 8 of a thread's 16 loads in flight at a time and
// spent 40 % of its warp-cycles on the long scoreboard: profiles/r02_ncu_k_xpass_fused.txt).
// TWG: twiddles read through L1 from the global table instead of a shared-memory copy (the
// 16 KB table of N = 1024 is what keeps a third CTA off the SM).
template <int N, int CK, int NT, int MINB, bool TWG = false>
__global__ void __launch_bounds__(NT, MINB)
k_xpass_fused_async(double2* __restrict__ T, xpass::Pointwise pw, const double2* __restrict__ tw_g) {
  using namespace xpass;
  extern __shared__ __align__(16) unsigned char xp_smem[];
  double2* tile = reinterpret_cast<double2*>(xp_smem);
  const double2* tw = TWG ? tw_g : tile + N * CK;
  const int tid = threadIdx.x;
  const long long c0 = (long long)blockIdx.x * CK;
  stage_load_async<N, CK, NT>(tid, T, pw.ncols, c0, pw.ncols, tile);
  if (!TWG) {
    double2* tws = tile + N * CK;
    for (int t = tid; t < N; t += NT) tws[t] = tw_g[t];
  }
  const Column col = column_of(pw, c0 + tid % CK);
  stage_load_wait();
  __syncthreads();
  constexpr int NS = Radix<N>::NS;
  stage_fwd<N, CK, NT, 0>(tid, tile, tw);
  __syncthreads();
  if constexpr (NS >= 3) { stage_fwd<N, CK, NT, 1>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 4) { stage_fwd<N, CK, NT, 2>(tid, tile, tw); __syncthreads(); }
  stage_junction<N, CK, NT>(tid, tile, pw, col);
  __syncthreads();
  if constexpr (NS >= 4) { stage_inv<N, CK, NT, 2>(tid, tile, tw); __syncthreads(); }
  if constexpr (NS >= 3) { stage_inv<N, CK, NT, 1>(tid, tile, tw); __syncthreads(); }
  stage_last<N, CK, NT>(tid, tile, tw, col, T, pw.ncols);
}

// exp(-2 pi i t / N), t < N, evaluated in long double on the host; one table per (device, N)
// for the life of the process (16 KB at most).
std::mutex g_tw_mutex;
std::map<std::pair<int, int>, double2*> g_tw_tables;

}  // namespace

int trvb_twiddle_table(trvb_ctx* ctx, int N, const double2** out) {
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  const auto key = std::make_pair(ctx->device, N);
  auto it = g_tw_tables.find(key);
  if (it == g_tw_tables.end()) {
    std::vector<double2> h(N);
    for (int t = 0; t < N; t++) {
      const long double a = -2.0L * 3.141592653589793238462643383279502884L * t / N;
      h[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    // exact values on the axes
    h[0] = make_double2(1., 0.);
    if (N % 4 == 0) { h[N / 4] = make_double2(0., -1.); h[N / 2] = make_double2(-1., 0.); h[3 * N / 4] = make_double2(0., 1.); }
    double2* d = nullptr;
    TRVB_CUDA(cudaMalloc((void**)&d, sizeof(double2) * N));
    TRVB_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
    it = g_tw_tables.emplace(key, d).first;
  }
  *out = it->second;
  return 0;
}

namespace {

template <int N, int CK, int MINB, int NT = 256, bool ASYNC = false, bool TWG = false>
int launch_xpass(trvb_ctx* ctx, double2* T, const xpass::Pointwise& pw, const double2* tw) {
  auto kernel = ASYNC ? k_xpass_fused_async<N, CK, NT, MINB, TWG> : k_xpass_fused<N, CK, NT, MINB>;
  const size_t smem = sizeof(double2) * ((size_t)N * CK + (TWG ? 0 : N));
  static std::mutex attr_mutex;
  static std::map<int, bool> attr_done;   // per device
  {
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (!attr_done[ctx->device]) {
      TRVB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done[ctx->device] = true;
    }
  }
  const long long tiles = (pw.ncols + CK - 1) / CK;
  TRVB_REQUIRE(tiles < 2147483647LL, "trvb_box_fields_fused: too many column tiles");
  kernel<<<(unsigned)tiles, NT, smem, ctx->stream>>>(T, pw, tw);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Batched 2-D plans over the x-planes of ctx's grid, without their own work areas (the
// area comes from the arena for the duration of one execution).  Kept in ctx->line_plans
// (destroyed with the context) under keys that no 1-D plan uses.
int get_plane_plan(trvb_ctx* ctx, cufftType type, cufftHandle* out, size_t* work_bytes) {
  const GridDesc& g = ctx->g;
  const std::vector<long long> key = {(long long)type, -2, g.n[0], g.n[1], g.n[2]};
  auto it = ctx->line_plans.find(key);
  if (it == ctx->line_plans.end()) {
    cufftHandle plan;
    TRVB_CUFFT(cufftCreate(&plan));
    TRVB_CUFFT(cufftSetAutoAllocation(plan, 0));
    int n2d[2] = {g.n[1], g.n[2]};
    int re_embed[2] = {g.n[1], g.n[2]}, cx_embed[2] = {g.n[1], g.nh};
    size_t ws = 0;
    cufftResult rc;
    if (type == CUFFT_D2Z) {
      rc = cufftMakePlanMany(plan, 2, n2d, re_embed, 1, g.n[1] * g.n[2], cx_embed, 1,
                             g.n[1] * g.nh, CUFFT_D2Z, g.n[0], &ws);
    } else {
      rc = cufftMakePlanMany(plan, 2, n2d, cx_embed, 1, g.n[1] * g.nh, re_embed, 1,
                             g.n[1] * g.n[2], CUFFT_Z2D, g.n[0], &ws);
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

std::atomic<long long> g_fused_calls{0};
// TRV_XPASS_TRACE=1: CUDA-event times of the three steps of the last call (ms), read back
// with trvb_box_fields_fused_last_ms (the call then ends with a stream synchronisation).
std::mutex g_trace_mutex;
double g_last_ms[3] = {0., 0., 0.};

bool length_supported(int n) {
  return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048;
}

}  // namespace

extern "C" int trvb_box_fields_fused_supported(const trvb_ctx* ctx) {
  if (!ctx || ctx->parent) return 0;
  const GridDesc& g = ctx->g;
  if (!length_supported(g.n[0])) return 0;
  if ((long long)g.n[1] * g.n[2] >= 2147483647LL) return 0;
  return 1;
}

extern "C" int trvb_box_fields_fused(trvb_ctx* ctx, trvb_ctx* sub, trvb_mesh x, double add_a,
                                     double add_b, const double S[2], trvb_mesh lowk,
                                     trvb_mesh xi) {
  TRVB_REQUIRE(ctx && sub && x.data && S && lowk.data && xi.data,
               "trvb_box_fields_fused: null argument");
  TRVB_REQUIRE(trvb_box_fields_fused_supported(ctx),
               "trvb_box_fields_fused: n0 = %d is not a supported length (32 .. 2048, power of two)",
               ctx->g.n[0]);
  TRVB_REQUIRE(sub->parent == ctx, "trvb_box_fields_fused: `sub` must be a sub-grid of the mesh");
  TRVB_REQUIRE(x.layout == TRVB_REAL && xi.layout == TRVB_REAL && lowk.layout == TRVB_HALF,
               "trvb_box_fields_fused: x and xi REAL, lowk HALF");
  TRVB_REQUIRE(S[1] == 0., "trvb_box_fields_fused: real shot-noise amplitude required");
  TRVB_REQUIRE(x.data != xi.data, "trvb_box_fields_fused: xi aliases x");
  const GridDesc& g = ctx->g;
  TRVB_CUDA(cudaSetDevice(ctx->device));
  g_fused_calls++;
  const double2* tw = nullptr;
  int st = trvb_twiddle_table(ctx, g.n[0], &tw);
  if (st) return st;
  cufftHandle fwd, inv;
  size_t ws_fwd = 0, ws_inv = 0;
  st = get_plane_plan(ctx, CUFFT_D2Z, &fwd, &ws_fwd); if (st) return st;
  st = get_plane_plan(ctx, CUFFT_Z2D, &inv, &ws_inv); if (st) return st;

  const char* env_trace = getenv("TRV_XPASS_TRACE");
  const bool trace = env_trace && env_trace[0] == '1';
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  auto mark = [&](int i) {
    if (trace) { cudaEventCreate(&ev[i]); cudaEventRecord(ev[i], ctx->stream); }
  };

  double2* T = nullptr;
  void* area = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&T, trvb_mesh_bytes(ctx, TRVB_HALF)));
  struct Guard {
    trvb_ctx* ctx; double2*& T; void*& area;
    ~Guard() { if (area) trvb_dev_free_raw(ctx, area); if (T) trvb_dev_free_raw(ctx, T); }
  } guard{ctx, T, area};

  // -- 2-D transforms of the planes --
  if (ws_fwd) { TRVB_CUDA(trvb_dev_alloc_raw(ctx, &area, ws_fwd)); TRVB_CUFFT(cufftSetWorkArea(fwd, area)); }
  mark(0);
  TRVB_CUFFT(cufftExecD2Z(fwd, (cufftDoubleReal*)x.data, (cufftDoubleComplex*)T));
  g_trvb_fft_execs++;
  if (area) { trvb_dev_free_raw(ctx, area); area = nullptr; }

  // -- x pass: forward, low-|k| modes, spectrum, inverse --
  TRVB_CUDA(cudaMemsetAsync(lowk.data, 0, trvb_mesh_bytes(sub, TRVB_HALF), ctx->stream));
  mark(1);
  xpass::Pointwise pw;
  pw.n0 = g.n[0]; pw.n1 = g.n[1]; pw.nh = g.nh; pw.ncols = (long long)g.n[1] * g.nh;
  pw.ralias0 = ctx->d_ralias[0]; pw.ralias1 = ctx->d_ralias[1]; pw.ralias2 = ctx->d_ralias[2];
  pw.add_a = add_a; pw.add_b = add_b; pw.S_re = S[0]; pw.S_im = S[1]; pw.inv_vol = 1. / g.vol;
  pw.s0 = sub->g.n[0]; pw.s1 = sub->g.n[1]; pw.s2 = sub->g.n[2]; pw.sh = sub->g.nh;
  pw.lowk = (double2*)lowk.data;
  // Launch shapes measured on B200 (scripts/xpass_bench.py, profiles/r02_xpass_variants.txt):
  // 512: cp.async tile, 128 threads (four butterflies per thread and stage: the compiler
  // overlaps their shared-memory loads), three CTAs per SM: 0.55 ms against 0.71 ms for the
  // register-staged 256-thread kernel.  1024: the tile is 64 KB with FOUR columns (64-byte
  // segments per plane); cp.async, 128 threads and the twiddles read through L1 (no 16 KB
  // copy in shared memory: a third CTA fits): 7.1 ms against 7.7 ms register-staged.
  // TRV_XPASS_VARIANT=1 selects the other kernel of the pair (A/B runs).
  const char* env_v = getenv("TRV_XPASS_VARIANT");
  const bool other = env_v && env_v[0] == '1';
  switch (g.n[0]) {
    case 32:   st = launch_xpass<32, 8, 3>(ctx, T, pw, tw); break;
    case 64:   st = launch_xpass<64, 8, 3>(ctx, T, pw, tw); break;
    case 128:  st = launch_xpass<128, 8, 3>(ctx, T, pw, tw); break;
    case 256:  st = other ? launch_xpass<256, 8, 3>(ctx, T, pw, tw)
                          : launch_xpass<256, 8, 3, 128, true>(ctx, T, pw, tw); break;
    case 512:  st = other ? launch_xpass<512, 8, 3>(ctx, T, pw, tw)
                          : launch_xpass<512, 8, 3, 128, true>(ctx, T, pw, tw); break;
    case 1024: st = other ? launch_xpass<1024, 4, 2>(ctx, T, pw, tw)
                          : launch_xpass<1024, 4, 3, 128, true, true>(ctx, T, pw, tw); break;
    default:   st = launch_xpass<2048, 4, 1>(ctx, T, pw, tw); break;
  }
  if (st) return st;

  if (ws_inv) { TRVB_CUDA(trvb_dev_alloc_raw(ctx, &area, ws_inv)); TRVB_CUFFT(cufftSetWorkArea(inv, area)); }
  mark(2);
  TRVB_CUFFT(cufftExecZ2D(inv, (cufftDoubleComplex*)T, (cufftDoubleReal*)xi.data));
  g_trvb_fft_execs++;
  mark(3);
  if (trace) {
    TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lock(g_trace_mutex);
    for (int i = 0; i < 3; i++) {
      float ms = 0.f; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); g_last_ms[i] = ms;
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  }

  // readers index the low-|k| mesh by the big grid (KView::p0) while it is registered
  ctx->lowk_ptr = lowk.data;
  for (int a = 0; a < 3; a++) ctx->lowk_dims[a] = sub->g.n[a];
  return 0;
}

extern "C" long long trvb_box_fields_fused_call_count(void) { return g_fused_calls; }

extern "C" void trvb_box_fields_fused_last_ms(double out[3]) {
  std::lock_guard<std::mutex> lock(g_trace_mutex);
  for (int i = 0; i < 3; i++) out[i] = g_last_ms[i];
}

extern "C" void trvb_ctx_forget_lowk(trvb_ctx* ctx) {
  if (ctx) ctx->lowk_ptr = nullptr;
}
