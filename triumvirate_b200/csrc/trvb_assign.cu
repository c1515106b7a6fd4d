// trvb_assign.cu -- device-resident catalogue, cell sort and particle-to-mesh
// assignment (NGP/CIC/TSC/PCS, optional half-cell shift) for libtrvb.so.
//
// Replaces the OpenMP scatter loops of S/field.cpp:618-1112 and the weight
// kernels of S/field.cpp:1259-1269,1379-1391.  Two modes:
//   throughput     particles are counting-sorted by (32x16x8-cell tile, 4x4 column
//                  of the tile) and their columns gathered into that order as
//                  32-byte records {n x / L, w}.  TSC/PCS (k_assign_coop): a warp
//                  spreads ONE particle per instruction, lanes = (row, z-cell) of
//                  its stencil, so the red.global.add.f64 of a stencil row coalesce
//                  into one or two 32-byte L2 sectors (the L2 retires ~one RED
//                  sector per slice per clock: sectors, not elements, are the
//                  cost).  NGP/CIC: one thread per particle.  Opt-in
//                  (TRV_ASSIGN_TILE=1, unshifted TSC/PCS): k_assign_tile, one CTA
//                  per tile accumulates the tile's footprint (35x19x11 cells) in
//                  SHARED memory -- half-warps spread one particle per step with
//                  plain LDS/DFMA/STS, columns of equal (x, y) parity have
//                  disjoint stencils so the 8 half-warps of a CTA work without
//                  atomics -- and flushes it with one RED per footprint cell
//                  (3.4x fewer RED sectors).  Measured: no faster than the
//                  cooperative scatter (SM-latency bound at 8 warps per SM), see
//                  profiles/r01d_assign_tile_vs_coop.txt.
//   deterministic  particles are counting-sorted by home cell with ascending
//                  particle id inside a cell; one thread per OUTPUT cell
//                  gathers its contributions and adds them in ascending
//                  particle order with non-contracted IEEE operations, i.e.
//                  the single-threaded reference order => bit-identical mesh.
#include "trvb_common.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <thread>

namespace {

// ---------------------------------------------------------------------
// Window functions: the exact operation sequence of S/field.cpp.
// ---------------------------------------------------------------------

// loc_grid = ngrid * pos / boxsize (+0.5 and wrap for the shadow mesh),
// S/field.cpp:643-644, 689-694.
__device__ __forceinline__ double shift_loc(double loc, int n, int shifted) {
  if (shifted) {
    loc = __dadd_rn(loc, 0.5);
    if (loc > (double)n) loc = __dsub_rn(loc, (double)n);
  }
  return loc;
}

__device__ __forceinline__ double grid_loc(double pos, int n, double L, int shifted) {
  return shift_loc(__ddiv_rn(__dmul_rn((double)n, pos), L), n, shifted);
}

template <int ORDER>
__device__ __forceinline__ void window_1d(double loc, int n, int* ijk, double* win);

// NGP, S/field.cpp:646-655.
template <>
__device__ __forceinline__ void window_1d<1>(double loc, int n, int* ijk, double* win) {
  int idx = __double2int_rz(loc);
  if (__dsub_rn(loc, (double)idx) >= 0.5) idx = (idx == n - 1) ? 0 : idx + 1;
  ijk[0] = idx;
  win[0] = 1.;
}

// CIC, S/field.cpp:756-766.
template <>
__device__ __forceinline__ void window_1d<2>(double loc, int n, int* ijk, double* win) {
  int idx = __double2int_rz(loc);
  ijk[0] = idx;
  ijk[1] = (idx == n - 1) ? 0 : idx + 1;
  double s = __dsub_rn(loc, (double)idx);
  win[0] = __dsub_rn(1., s);
  win[1] = s;
}

// TSC, S/field.cpp:867-895.  (The shadow-mesh variant at S/field.cpp:944-950
// indexes out of range near the upper edge -- SURVEY.md F5b; the periodic wrap
// of the primary mesh is used for both here.)
template <>
__device__ __forceinline__ void window_1d<3>(double loc, int n, int* ijk, double* win) {
  int idx = __double2int_rz(loc);
  double s = __dsub_rn(loc, (double)idx);
  if (s < 0.5) {
    ijk[0] = (idx == 0) ? n - 1 : idx - 1;
    ijk[1] = idx;
    ijk[2] = (idx == n - 1) ? 0 : idx + 1;
    double a = __dsub_rn(0.5, s), b = __dadd_rn(0.5, s);
    win[0] = __dmul_rn(__dmul_rn(0.5, a), a);
    win[1] = __dsub_rn(0.75, __dmul_rn(s, s));
    win[2] = __dmul_rn(__dmul_rn(0.5, b), b);
  } else {
    ijk[0] = idx;
    ijk[1] = (idx == n - 1) ? 0 : idx + 1;
    ijk[2] = (ijk[1] == n - 1) ? 0 : ijk[1] + 1;
    s = __dsub_rn(1., s);
    double a = __dsub_rn(0.5, s), b = __dadd_rn(0.5, s);
    win[0] = __dmul_rn(__dmul_rn(0.5, b), b);
    win[1] = __dsub_rn(0.75, __dmul_rn(s, s));
    win[2] = __dmul_rn(__dmul_rn(0.5, a), a);
  }
}

// PCS, S/field.cpp:1015-1033.
template <>
__device__ __forceinline__ void window_1d<4>(double loc, int n, int* ijk, double* win) {
  int idx = __double2int_rz(loc);
  ijk[0] = (idx == 0) ? n - 1 : idx - 1;
  ijk[1] = idx;
  ijk[2] = (idx == n - 1) ? 0 : idx + 1;
  ijk[3] = (ijk[2] == n - 1) ? 0 : ijk[2] + 1;
  const double c6 = 1. / 6;
  double s = __dsub_rn(loc, (double)idx);
  double u = __dsub_rn(1., s);
  win[0] = __dmul_rn(__dmul_rn(__dmul_rn(c6, u), u), u);
  win[1] = __dmul_rn(c6, __dadd_rn(
    __dsub_rn(4., __dmul_rn(__dmul_rn(6., s), s)),
    __dmul_rn(__dmul_rn(__dmul_rn(3., s), s), s)));
  win[2] = __dmul_rn(c6, __dadd_rn(
    __dsub_rn(4., __dmul_rn(__dmul_rn(6., u), u)),
    __dmul_rn(__dmul_rn(__dmul_rn(3., u), u), u)));
  win[3] = __dmul_rn(__dmul_rn(__dmul_rn(c6, s), s), s);
}

// Home cell index along one axis (the integer part of loc), clamped.
__device__ __forceinline__ int home_index(double pos, int n, double L, int shifted) {
  double loc = grid_loc(pos, n, L, shifted);
  int idx = __double2int_rz(loc);
  return min(max(idx, 0), n - 1);
}

// ---------------------------------------------------------------------
// Particle weights (kinds in trvb.h).
// ---------------------------------------------------------------------

struct CatView {
  const double* x; const double* y; const double* z;
  const double* w;     // may be null
  const double* lx; const double* ly; const double* lz;   // may be null
  const double* cw;    // custom complex weights, may be null
  long long n;
};

__device__ __forceinline__ cplx particle_weight(const CatView& c, long long i,
                                                int kind, int L, int M) {
  cplx out; out.im = 0.;
  const double w = c.w ? c.w[i] : 1.;
  if (kind == TRVB_W_UNIT) { out.re = 1.; return out; }
  if (kind == TRVB_W_W) { out.re = w; return out; }
  if (kind == TRVB_W_CUSTOM) { out.re = c.cw[2 * i]; out.im = c.cw[2 * i + 1]; return out; }
  cplx y; y.re = 1.; y.im = 0.;
  if (!(L == 0 && M == 0)) y = ylm_reduced(L, M, c.lx[i], c.ly[i], c.lz[i]);
  if (kind == TRVB_W_YLM_W) {
    out.re = y.re * w; out.im = y.im * w;
  } else if (kind == TRVB_W_CYLM_W2) {
    const double w2 = w * w;
    out.re = y.re * w2; out.im = -y.im * w2;
  } else {  // TRVB_W_YLM_W3
    const double w3 = w * w * w;
    out.re = y.re * w3; out.im = y.im * w3;
  }
  return out;
}

// The catalogue in sort order: one 32-byte record per particle (a single sector
// for the spreading kernels) holding the GRID coordinates loc = n x / L of the
// unshifted mesh -- evaluated once by the sort, with the reference's operation
// order, instead of three fp64 divisions per particle per assignment -- and w.
struct SortedView {
  const double4* p4;                                      // {loc_x, loc_y, loc_z, w}
  const double* lx; const double* ly; const double* lz;   // may be null
  const double* cw;                                       // may be null
  long long n;
};

__device__ __forceinline__ cplx particle_weight(const SortedView& c, long long i,
                                                const double4& p, int kind, const YlmCoef& yc) {
  const int L = yc.ell, M = yc.m;
  cplx out; out.im = 0.;
  const double w = p.w;
  if (kind == TRVB_W_UNIT) { out.re = 1.; return out; }
  if (kind == TRVB_W_W) { out.re = w; return out; }
  if (kind == TRVB_W_CUSTOM) { out.re = c.cw[2 * i]; out.im = c.cw[2 * i + 1]; return out; }
  cplx y; y.re = 1.; y.im = 0.;
  if (!(L == 0 && M == 0)) y = ylm_eval(yc, c.lx[i], c.ly[i], c.lz[i]);
  if (kind == TRVB_W_YLM_W) {
    out.re = y.re * w; out.im = y.im * w;
  } else if (kind == TRVB_W_CYLM_W2) {
    const double w2 = w * w;
    out.re = y.re * w2; out.im = -y.im * w2;
  } else {  // TRVB_W_YLM_W3
    const double w3 = w * w * w;
    out.re = y.re * w3; out.im = y.im * w3;
  }
  return out;
}

// ---------------------------------------------------------------------
// Counting sort by tile (throughput) or by home cell (deterministic).
// ---------------------------------------------------------------------

// Throughput sort key: home-cell tile of TILE_X x TILE_Y x TILE_Z cells, then the
// column (COL_W x COL_W cells in x, y; full tile depth) inside it.
constexpr int TILE_X = 32, TILE_Y = 16, TILE_Z = 8, COL_W = 4;
constexpr int COLS_X = TILE_X / COL_W, COLS_Y = TILE_Y / COL_W;   // columns per tile edge
constexpr int KEYS_PER_TILE = COLS_X * COLS_Y;

struct SortDesc {
  int n[3]; double L[3]; int shifted;
  int by_cell;          // 0: (tile, column) key, 1: home-cell key
  int nk[3];            // key-grid extents: tiles (by_cell == 0) or cells
};

__device__ __forceinline__ int sort_key(const SortDesc& d, double x, double y, double z) {
  int i = home_index(x, d.n[0], d.L[0], d.shifted);
  int j = home_index(y, d.n[1], d.L[1], d.shifted);
  int k = home_index(z, d.n[2], d.L[2], d.shifted);
  if (!d.by_cell) {
    const int tile = ((i / TILE_X) * d.nk[1] + j / TILE_Y) * d.nk[2] + k / TILE_Z;
    return tile * KEYS_PER_TILE + ((i % TILE_X) / COL_W) * COLS_Y + (j % TILE_Y) / COL_W;
  }
  return (i * d.nk[1] + j) * d.nk[2] + k;
}

__global__ void k_sort_count(CatView c, SortDesc d, int* __restrict__ counts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c.n;
       i += (long long)gridDim.x * blockDim.x) {
    atomicAdd(&counts[sort_key(d, c.x[i], c.y[i], c.z[i])], 1);
  }
}

// Exclusive scan of `counts` (length nkeys, plus one trailing total slot),
// three kernels: chunk sums, scan of chunk sums, chunk scan + offset.
constexpr int SCAN_CHUNK = 2048;
constexpr int SCAN_THREADS = 256;

__global__ void k_scan_chunk_sums(const int* __restrict__ counts, long long nkeys,
                                  int* __restrict__ chunk_sums) {
  __shared__ int sm[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_CHUNK;
  int v = 0;
  for (int t = threadIdx.x; t < SCAN_CHUNK; t += SCAN_THREADS) {
    long long idx = base + t;
    if (idx < nkeys) v += counts[idx];
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int wdx = 0; wdx < SCAN_THREADS / 32; wdx++) s += sm[wdx];
    chunk_sums[blockIdx.x] = s;
  }
}

__global__ void k_scan_chunk_offsets(int* __restrict__ chunk_sums, int nchunks) {
  // Single block: serial over tiles of 1024 with an in-block scan.
  __shared__ int sm[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += 1024) {
    int idx = base + threadIdx.x;
    int v = (idx < nchunks) ? chunk_sums[idx] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = sm[threadIdx.x];
    if (idx < nchunks) chunk_sums[idx] = carry + incl - v;   // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
}

__global__ void k_scan_apply(int* __restrict__ counts, long long nkeys,
                             const int* __restrict__ chunk_offsets) {
  // In place: counts -> exclusive offsets.  One block per chunk, thread t
  // owns SCAN_CHUNK/SCAN_THREADS consecutive entries.
  constexpr int PER = SCAN_CHUNK / SCAN_THREADS;
  __shared__ int sm[SCAN_THREADS];
  const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * PER;
  int vals[PER];
  int local = 0;
#pragma unroll
  for (int q = 0; q < PER; q++) {
    long long idx = base + q;
    vals[q] = (idx < nkeys) ? counts[idx] : 0;
    local += vals[q];
  }
  sm[threadIdx.x] = local;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {
    int t = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
    __syncthreads();
    sm[threadIdx.x] += t;
    __syncthreads();
  }
  int run = chunk_offsets[blockIdx.x] + sm[threadIdx.x] - local;
#pragma unroll
  for (int q = 0; q < PER; q++) {
    long long idx = base + q;
    if (idx < nkeys) counts[idx] = run;
    run += vals[q];
  }
}

// `s4` non-null: also store the packed record at its sorted slot (a full
// 32-byte sector write; no later gather of the positions is needed).
__global__ void k_sort_scatter(CatView c, SortDesc d, int* __restrict__ cursor,
                               int* __restrict__ order, double4* __restrict__ s4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c.n;
       i += (long long)gridDim.x * blockDim.x) {
    const double x = c.x[i], y = c.y[i], z = c.z[i];
    int key = sort_key(d, x, y, z);
    int pos = atomicAdd(&cursor[key], 1);
    if (order) order[pos] = (int)i;
    if (s4) {
      s4[pos] = make_double4(grid_loc(x, d.n[0], d.L[0], 0), grid_loc(y, d.n[1], d.L[1], 0),
                             grid_loc(z, d.n[2], d.L[2], 0), c.w ? c.w[i] : 1.);
    }
  }
}

// After the atomic scatter the ids inside a cell are in arbitrary order;
// sort each cell's (short) segment ascending so that the gather kernel sees
// the reference's particle order.  cursor[key] now holds the segment END.
__global__ void k_sort_segments(const int* __restrict__ seg_end, long long nkeys,
                                int* __restrict__ order) {
  for (long long key = blockIdx.x * (long long)blockDim.x + threadIdx.x; key < nkeys;
       key += (long long)gridDim.x * blockDim.x) {
    int b = (key == 0) ? 0 : seg_end[key - 1];
    int e = seg_end[key];
    for (int a = b + 1; a < e; a++) {
      int v = order[a];
      int q = a - 1;
      while (q >= b && order[q] > v) { order[q + 1] = order[q]; q--; }
      order[q + 1] = v;
    }
  }
}

// Gather catalogue columns into sorted order: the packed records (when the
// scatter did not write them), lines of sight and custom weights.
__global__ void k_gather_sorted(CatView c, SortDesc d, const int* __restrict__ order,
                                double4* __restrict__ s4, double* __restrict__ slos,
                                double* __restrict__ scw) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < c.n;
       s += (long long)gridDim.x * blockDim.x) {
    const long long i = order[s];
    if (s4) {
      s4[s] = make_double4(grid_loc(c.x[i], d.n[0], d.L[0], 0), grid_loc(c.y[i], d.n[1], d.L[1], 0),
                           grid_loc(c.z[i], d.n[2], d.L[2], 0), c.w ? c.w[i] : 1.);
    }
    if (slos) { slos[s] = c.lx[i]; slos[c.n + s] = c.ly[i]; slos[2 * c.n + s] = c.lz[i]; }
    if (scw) { scw[2 * s] = c.cw[2 * i]; scw[2 * s + 1] = c.cw[2 * i + 1]; }
  }
}

// ---------------------------------------------------------------------
// Throughput assignment.
// ---------------------------------------------------------------------

// One thread per particle, p^3 REDs (NGP/CIC).  `c` is the SORTED view.
template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(256)
k_assign_scatter(SortedView c, GridDesc g, int shifted,
                 int kind, YlmCoef yc, double scale, double* __restrict__ mesh) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c.n;
       i += (long long)gridDim.x * blockDim.x) {
    int ijk[3][ORDER];
    double win[3][ORDER];
    const double4 p = c.p4[i];
    window_1d<ORDER>(shift_loc(p.x, g.n[0], shifted), g.n[0], ijk[0], win[0]);
    window_1d<ORDER>(shift_loc(p.y, g.n[1], shifted), g.n[1], ijk[1], win[1]);
    window_1d<ORDER>(shift_loc(p.z, g.n[2], shifted), g.n[2], ijk[2], win[2]);
    cplx wt = particle_weight(c, i, p, kind, yc);
    const double bre = __dmul_rn(scale, wt.re);
    const double bim = COMPLEX ? __dmul_rn(scale, wt.im) : 0.;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
      const double wa_re = __dmul_rn(bre, win[0][a]);
      const double wa_im = COMPLEX ? __dmul_rn(bim, win[0][a]) : 0.;
#pragma unroll
      for (int b = 0; b < ORDER; b++) {
        const double wb_re = __dmul_rn(wa_re, win[1][b]);
        const double wb_im = COMPLEX ? __dmul_rn(wa_im, win[1][b]) : 0.;
        const long long row = ((long long)ijk[0][a] * g.n[1] + ijk[1][b]) * g.n[2];
#pragma unroll
        for (int cidx = 0; cidx < ORDER; cidx++) {
          const long long gid = row + ijk[2][cidx];
          if (gid >= 0 && gid < g.nmesh) {   // S/field.cpp:1042
            if (COMPLEX) {
              atomicAdd(&mesh[2 * gid], __dmul_rn(wb_re, win[2][cidx]));
              atomicAdd(&mesh[2 * gid + 1], __dmul_rn(wb_im, win[2][cidx]));
            } else {
              atomicAdd(&mesh[gid], __dmul_rn(wb_re, win[2][cidx]));
            }
          }
        }
      }
    }
  }
}

// Warp-cooperative spreading (TSC/PCS).  Each lane first evaluates the 1-D
// windows of one particle into shared memory; the warp then walks its 32
// particles, every instruction covering RPP stencil rows x ORDER z-cells
// (x re/im) of ONE particle, so that the REDs of a row fall into one or two
// 32-byte sectors.  Same products, in the same order, as k_assign_scatter.
template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(256)
k_assign_coop(SortedView c, GridDesc g, int shifted,
              int kind, YlmCoef yc, double scale, double* __restrict__ mesh) {
  constexpr int NW = 8;                               // warps per block
  constexpr int LPR = ORDER * (COMPLEX ? 2 : 1);      // lanes per stencil row
  constexpr int RPP = 32 / LPR;                       // rows per instruction
  constexpr int NROW = ORDER * ORDER;
  constexpr int NPASS = (NROW + RPP - 1) / RPP;
  __shared__ double s_win[NW][32][3 * ORDER];
  __shared__ int s_idx[NW][32][3 * ORDER];
  __shared__ double s_wt[NW][32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane / LPR, q = lane - r * LPR;
  const int cz = COMPLEX ? (q >> 1) : q;
  const int comp = COMPLEX ? (q & 1) : 0;
  const long long nchunk = (c.n + 31) / 32;
  for (long long chunk = (long long)blockIdx.x * NW + warp; chunk < nchunk;
       chunk += (long long)gridDim.x * NW) {
    const long long i = chunk * 32 + lane;
    const bool ok = i < c.n;
    __syncwarp();
    if (ok) {
      int ijk[ORDER]; double win[ORDER];
      const double4 p = c.p4[i];
      window_1d<ORDER>(shift_loc(p.x, g.n[0], shifted), g.n[0], ijk, win);
#pragma unroll
      for (int t = 0; t < ORDER; t++) { s_win[warp][lane][t] = win[t]; s_idx[warp][lane][t] = ijk[t]; }
      window_1d<ORDER>(shift_loc(p.y, g.n[1], shifted), g.n[1], ijk, win);
#pragma unroll
      for (int t = 0; t < ORDER; t++) { s_win[warp][lane][ORDER + t] = win[t]; s_idx[warp][lane][ORDER + t] = ijk[t]; }
      window_1d<ORDER>(shift_loc(p.z, g.n[2], shifted), g.n[2], ijk, win);
#pragma unroll
      for (int t = 0; t < ORDER; t++) { s_win[warp][lane][2 * ORDER + t] = win[t]; s_idx[warp][lane][2 * ORDER + t] = ijk[t]; }
      const cplx wt = particle_weight(c, i, p, kind, yc);
      s_wt[warp][lane][0] = __dmul_rn(scale, wt.re);
      s_wt[warp][lane][1] = __dmul_rn(scale, wt.im);
    }
    __syncwarp();
    const int count = (int)min((long long)32, c.n - chunk * 32);
    if (r < RPP) {
      for (int s = 0; s < count; s++) {
        const double* w = s_win[warp][s];
        const int* id = s_idx[warp][s];
        const double base = s_wt[warp][s][comp];
        const double wz = w[2 * ORDER + cz];
        const int kz = id[2 * ORDER + cz];
#pragma unroll
        for (int pass = 0; pass < NPASS; pass++) {
          const int row = r + RPP * pass;
          if (row < NROW) {
            const int a = row / ORDER, b = row - a * ORDER;
            const long long gid = ((long long)id[a] * g.n[1] + id[ORDER + b]) * g.n[2] + kz;
            const double v = __dmul_rn(__dmul_rn(__dmul_rn(base, w[a]), w[ORDER + b]), wz);
            if (gid >= 0 && gid < g.nmesh) {   // S/field.cpp:1042
              atomicAdd(&mesh[COMPLEX ? 2 * gid + comp : gid], v);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------
// Tile-owned shared-memory accumulation (TSC/PCS, unshifted mesh).
// ---------------------------------------------------------------------

// Footprint of a tile's stencils: cells [t0 - 1, t0 + TILE + 1] per axis.
constexpr int FP_X = TILE_X + 3, FP_Y = TILE_Y + 3, FP_Z = TILE_Z + 3;
constexpr int TILE_SLOTS = 32;     // particles staged per warp per round (16 per column)
constexpr int TILE_WARPS = 4;      // x 2 columns each = the 8 columns of one colour

// Shared-tile pitches in doubles.  A half-warp spreads one particle: lane ->
// stencil row (a, b), walking the row's z-cells.  REAL: row pitch 24 words puts
// b = 0..3 on banks {0, 24, 16, 8}; plane pitch == 2 (mod 32) words shifts a by one
// double: the 16 rows of a half-warp hit 16 distinct bank pairs.  COMPLEX (16-byte
// cells): row pitch 56 words, plane pitch == 4 (mod 32): each quarter-warp covers
// all 32 banks once.
template <bool COMPLEX> struct TileGeom;
template <> struct TileGeom<false> {
  static constexpr int CW = 1, ROW = 12, PLANE = FP_Y * 12 + 13;      // 241 doubles = 482 words
};
template <> struct TileGeom<true> {
  static constexpr int CW = 2, ROW = 28, PLANE = FP_Y * 28 + 14;      // 546 doubles = 1092 words
};

// Per-particle staging, read as 16-byte words during the spread: per stencil
// entry along x {window weight, (z offset << 32) | x offset} and along y {window
// weight, y offset}; the z weights, already multiplied by the particle weight.
// The z-run of a stencil row is contiguous in the tile.
template <bool COMPLEX> struct StageT;
template <> struct StageT<false> { double2 x[4], y[4]; double2 zp[2]; };     // 160 B
template <> struct StageT<true> { double2 x[4], y[4]; double2 zp[4]; };      // 192 B

template <bool COMPLEX>
constexpr size_t tile_smem_bytes() {
  return sizeof(double) * ((FP_X * TileGeom<COMPLEX>::PLANE + 1) & ~1)
    + sizeof(StageT<COMPLEX>) * TILE_SLOTS * TILE_WARPS;
}

// Reference-order scatter of ONE particle straight to global memory (the body
// of k_assign_scatter): slow path for particles outside [0, L).
template <int ORDER, bool COMPLEX>
__device__ void scatter_one(const double4& p, double bre, double bim, const GridDesc& g,
                            double* __restrict__ mesh) {
  int ijk[3][ORDER];
  double win[3][ORDER];
  window_1d<ORDER>(p.x, g.n[0], ijk[0], win[0]);
  window_1d<ORDER>(p.y, g.n[1], ijk[1], win[1]);
  window_1d<ORDER>(p.z, g.n[2], ijk[2], win[2]);
  for (int a = 0; a < ORDER; a++) {
    for (int b = 0; b < ORDER; b++) {
      const long long row = ((long long)ijk[0][a] * g.n[1] + ijk[1][b]) * g.n[2];
      const double wab = __dmul_rn(win[0][a], win[1][b]);
      for (int cidx = 0; cidx < ORDER; cidx++) {
        const long long gid = row + ijk[2][cidx];
        if (gid >= 0 && gid < g.nmesh) {   // S/field.cpp:1042
          const double wabc = __dmul_rn(wab, win[2][cidx]);
          if (COMPLEX) {
            atomicAdd(&mesh[2 * gid], __dmul_rn(bre, wabc));
            atomicAdd(&mesh[2 * gid + 1], __dmul_rn(bim, wabc));
          } else {
            atomicAdd(&mesh[gid], __dmul_rn(bre, wabc));
          }
        }
      }
    }
  }
}

// Local (tile footprint) index of mesh cell `idx` along an axis of n cells whose
// footprint starts at cell `origin` (= t0 - 1, may be -1); -1 if outside the
// footprint or outside the mesh (positions beyond the box edge: the reference
// does not wrap those, see scatter_one).  Requires n >= extent.
__device__ __forceinline__ int local_index(int idx, int origin, int n, int extent) {
  if ((unsigned)idx >= (unsigned)n) return -1;
  int la = idx - origin;
  if (la >= n) la -= n;
  if (la < 0) la += n;
  return ((unsigned)la < (unsigned)extent) ? la : -1;
}

template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(TILE_WARPS * 32)
k_assign_tile(SortedView c, const int* __restrict__ key_start, int nt1, int nt2,
              GridDesc g, int kind, YlmCoef yc, double scale, double* __restrict__ mesh) {
  typedef TileGeom<COMPLEX> G;
  typedef StageT<COMPLEX> Stage;
  constexpr int NT = (FP_X * G::PLANE + 1) & ~1;        // doubles in the shared tile (even)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Stage* st = reinterpret_cast<Stage*>(tile + NT) + warp * TILE_SLOTS;

  // The tile's KEYS_PER_TILE (= 32) column offsets live in one register per lane
  // (+ the end offset in every lane): one load latency per warp, no further
  // dependent offset loads.
  static_assert(KEYS_PER_TILE == 32, "one column offset per lane");
  const int key0 = blockIdx.x * KEYS_PER_TILE;
  const int key_lane = key_start[key0 + lane];
  const int key_end = key_start[key0 + KEYS_PER_TILE];
  if (__shfl_sync(0xffffffffu, key_lane, 0) == key_end) return;   // empty: mesh already zero
  const int tk = blockIdx.x % nt2, tj = (blockIdx.x / nt2) % nt1, ti = blockIdx.x / (nt2 * nt1);
  const int o0 = ti * TILE_X - 1, o1 = tj * TILE_Y - 1, o2 = tk * TILE_Z - 1;

  // Spreading role of a lane: half-warp h works on column h of the warp's pair,
  // lane & 15 is the stencil row (a, b).  Idle lanes and idle steps add zeros to
  // a private scratch run in the padding of plane `lane` (no divergence).
  const int h = lane >> 4, row = lane & 15;
  const int ra = (row < ORDER * ORDER) ? row / ORDER : 0;
  const int rb = (row < ORDER * ORDER) ? row - ra * ORDER : 0;
  const bool row_active = row < ORDER * ORDER;
  const int scratch_off = lane * G::PLANE + FP_Y * G::ROW;

  // Segments [b, e) of the two columns this warp spreads in a colour: the 8 columns
  // with (cx, cy) == (colour & 1, colour >> 1) (mod 2) have disjoint 7-cell-wide
  // footprints; column q = 2 * warp + h of them goes to half-warp h.
  auto segments = [&](int colour, int* seg_b, int* seg_e) {
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int q = 2 * warp + hh;
      const int cx = (colour & 1) + 2 * (q & 3), cy = (colour >> 1) + 2 * (q >> 2);
      const int key = cx * COLS_Y + cy;
      seg_b[hh] = __shfl_sync(0xffffffffu, key_lane, key);
      const int nxt = __shfl_sync(0xffffffffu, key_lane, (key + 1) & 31);
      seg_e[hh] = (key + 1 < KEYS_PER_TILE) ? nxt : key_end;
    }
  };
  // Records are requested ahead of use so that their HBM latency hides behind the
  // zero-fill and the spreading: the first round of ALL four colours up front (at
  // survey densities a column rarely holds more than the 16 particles of a round),
  // later rounds of a colour one round ahead.
  auto fetch = [&](int colour, int r0, double4& dst) {
    int sb[2], se[2];
    segments(colour, sb, se);
    const int idx = (h ? sb[1] : sb[0]) + r0 + row;
    if (idx < (h ? se[1] : se[0])) dst = c.p4[idx];
  };
  double4 p_first[4];
#pragma unroll
  for (int colour = 0; colour < 4; colour++) {
    p_first[colour] = make_double4(0., 0., 0., 0.);
    fetch(colour, 0, p_first[colour]);
  }

  for (int t = threadIdx.x; t < NT / 2; t += blockDim.x) {
    reinterpret_cast<double2*>(tile)[t] = make_double2(0., 0.);
  }
  __syncthreads();

  for (int colour = 0; colour < 4; colour++) {
    int seg_b[2], seg_e[2];
    segments(colour, seg_b, seg_e);
    const int my_b = h ? seg_b[1] : seg_b[0], my_e = h ? seg_e[1] : seg_e[0];
    const int longest = max(seg_e[0] - seg_b[0], seg_e[1] - seg_b[1]);
    double4 p_next = p_first[0];
#pragma unroll
    for (int t = 1; t < 4; t++) if (colour == t) p_next = p_first[t];   // registers, no local array
    for (int r0 = 0; r0 < longest; r0 += 16) {
      const double4 p = p_next;
      if (r0 + 16 < longest) fetch(colour, r0 + 16, p_next);
      // ---- stage: one lane per particle, slot = lane ---------------------------
      __syncwarp();
      if (my_b + r0 + row < my_e) {
        const long long i = my_b + r0 + row;
        const cplx wt = particle_weight(c, i, p, kind, yc);
        const double bre = __dmul_rn(scale, wt.re);
        const double bim = COMPLEX ? __dmul_rn(scale, wt.im) : 0.;
        Stage& s = st[lane];
        int ix[ORDER], iy[ORDER], iz[ORDER];
        double wx[ORDER], wy[ORDER], wz[ORDER];
        window_1d<ORDER>(p.x, g.n[0], ix, wx);
        window_1d<ORDER>(p.y, g.n[1], iy, wy);
        window_1d<ORDER>(p.z, g.n[2], iz, wz);
        // Inside the box the stencil indices are consecutive (mod n), so the first
        // one fixes the run in the tile footprint.
        bool ok = true;
#pragma unroll
        for (int t = 0; t < ORDER; t++) {
          ok = ok && (unsigned)ix[t] < (unsigned)g.n[0] && (unsigned)iy[t] < (unsigned)g.n[1]
            && (unsigned)iz[t] < (unsigned)g.n[2];
        }
        const int lx0 = local_index(ix[0], o0, g.n[0], FP_X - ORDER + 1);
        const int ly0 = local_index(iy[0], o1, g.n[1], FP_Y - ORDER + 1);
        const int lz0 = local_index(iz[0], o2, g.n[2], FP_Z - ORDER + 1);
        ok = ok && lx0 >= 0 && ly0 >= 0 && lz0 >= 0;
        if (!ok) {
          // Position outside [0, L): follow the reference's index arithmetic and
          // its bounds guard in global memory; contribute nothing to the tile.
          scatter_one<ORDER, COMPLEX>(p, bre, bim, g, mesh);
        }
        const int zoff = ok ? lz0 * G::CW : 0;
#pragma unroll
        for (int t = 0; t < ORDER; t++) {
          s.x[t] = make_double2(wx[t], __hiloint2double(zoff, ok ? (lx0 + t) * G::PLANE : 0));
          // high word: 1 marks a particle that contributes nothing to the tile
          s.y[t] = make_double2(wy[t], __hiloint2double(ok ? 0 : 1, ok ? (ly0 + t) * G::ROW : 0));
        }
        double zr[4], zi[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          zr[t] = (ok && t < ORDER) ? __dmul_rn(bre, wz[t < ORDER ? t : 0]) : 0.;
          zi[t] = (COMPLEX && ok && t < ORDER) ? __dmul_rn(bim, wz[t < ORDER ? t : 0]) : 0.;
        }
        if constexpr (COMPLEX) {
#pragma unroll
          for (int t = 0; t < 4; t++) s.zp[t] = make_double2(zr[t], zi[t]);
        } else {
          s.zp[0] = make_double2(zr[0], zr[1]);
          s.zp[1] = make_double2(zr[2], zr[3]);
        }
      }
      __syncwarp();
      // ---- spread: one particle per half-warp per step, staging prefetched ------
      const int cnt = min(16, my_e - my_b - r0);
      const int steps = min(16, longest - r0);
      const Stage* sp = st + h * 16;
      double2 X = sp->x[ra], Y = sp->y[rb];
      double2 Z[COMPLEX ? 4 : 2];
#pragma unroll
      for (int t = 0; t < (COMPLEX ? 4 : 2); t++) Z[t] = sp->zp[t];
      for (int sdx = 0; sdx < steps; sdx++) {
        // Idle rows / steps and particles outside the box touch no memory at all (a dummy
        // zero-add to a real cell could lose another warp's update to it).
        const bool live = row_active && sdx < cnt && __double2hiint(Y.y) == 0;
        const double wxy = live ? X.x * Y.x : 0.;
        double* cell = tile
          + (live ? __double2loint(X.y) + __double2loint(Y.y) + __double2hiint(X.y) : scratch_off);
        double2 Zc[COMPLEX ? 4 : 2];
#pragma unroll
        for (int t = 0; t < (COMPLEX ? 4 : 2); t++) Zc[t] = Z[t];
        if constexpr (COMPLEX) {
          double2 v[ORDER];
#pragma unroll
          for (int t = 0; t < ORDER; t++) {
            v[t] = live ? reinterpret_cast<const double2*>(cell)[t] : make_double2(0., 0.);
          }
          // next particle's staging (slots past the column's end hold stale, masked data)
          sp = st + h * 16 + min(sdx + 1, 15);
          X = sp->x[ra]; Y = sp->y[rb];
#pragma unroll
          for (int t = 0; t < 4; t++) Z[t] = sp->zp[t];
#pragma unroll
          for (int t = 0; t < ORDER; t++) {
            v[t].x = fma(wxy, Zc[t].x, v[t].x);
            v[t].y = fma(wxy, Zc[t].y, v[t].y);
            if (live) reinterpret_cast<double2*>(cell)[t] = v[t];
          }
        } else {
          double v[ORDER];
#pragma unroll
          for (int t = 0; t < ORDER; t++) v[t] = live ? cell[t] : 0.;
          sp = st + h * 16 + min(sdx + 1, 15);
          X = sp->x[ra]; Y = sp->y[rb];
          Z[0] = sp->zp[0]; Z[1] = sp->zp[1];
          const double zc[4] = {Zc[0].x, Zc[0].y, Zc[1].x, Zc[1].y};
#pragma unroll
          for (int t = 0; t < ORDER; t++) if (live) cell[t] = fma(wxy, zc[t], v[t]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }

  // Flush the footprint with REDs.  A thread owns one (ly, lz) of the 19 x 11 plane
  // (two passes) and walks the planes along x: constant strides, consecutive lanes
  // along z, no index arithmetic in the loop.
  const long long xstride = (long long)g.n[1] * g.n[2] * G::CW;
  for (int pq = threadIdx.x; pq < FP_Y * FP_Z; pq += blockDim.x) {
    const int ly = pq / FP_Z, lz = pq - ly * FP_Z;
    int gy = o1 + ly, gz = o2 + lz;                     // n >= footprint: one wrap suffices
    gy += (gy < 0) ? g.n[1] : 0; gy -= (gy >= g.n[1]) ? g.n[1] : 0;
    gz += (gz < 0) ? g.n[2] : 0; gz -= (gz >= g.n[2]) ? g.n[2] : 0;
    int gx = (o0 < 0) ? o0 + g.n[0] : o0;
    double* dst = mesh + (((long long)gx * g.n[1] + gy) * g.n[2] + gz) * G::CW;
    const double* src = tile + ly * G::ROW + lz * G::CW;
#pragma unroll 5
    for (int lx = 0; lx < FP_X; lx++) {
      const double vre = src[0];
      if (COMPLEX) {
        const double vim = src[1];
        if (vre != 0. || vim != 0.) { atomicAdd(dst, vre); atomicAdd(dst + 1, vim); }
      } else {
        if (vre != 0.) atomicAdd(dst, vre);
      }
      src += G::PLANE;
      dst += xstride;
      if (++gx == g.n[0]) { gx = 0; dst -= xstride * g.n[0]; }
    }
  }
}

// ---------------------------------------------------------------------
// Column-owned shared-memory accumulation (TSC/PCS, unshifted mesh, DENSE catalogues).
// ---------------------------------------------------------------------
// A warp owns one sort key = one 4x4x8-cell column: it accumulates the column's particles
// into its footprint (7 x 7 x 11 cells) in shared memory with plain LDS/DADD/STS -- one
// particle per instruction group, lanes = (stencil row, z-cell) exactly as in
// k_assign_coop -- and flushes the footprint with one RED per non-zero cell.  No block
// barriers, no colouring, 5-10 KB of shared memory per warp.  REDs per particle drop from
// 28 sectors to 539 / (particles in the column).  Opt-in (TRV_ASSIGN_COL=1): even on
// well-filled columns (the 5e7 randoms of BASELINE config 3, ~220 per occupied column) the
// serial LDS -> DADD -> STS chain per particle makes it slower than the direct scatter.
constexpr int COLF_X = COL_W + 3, COLF_Y = COL_W + 3, COLF_Z = TILE_Z + 3;   // 7 x 7 x 11
constexpr int COL_ROW = 12;                       // z pitch (doubles or cells)
constexpr int COL_PLANE = COLF_Y * COL_ROW + 1;   // 85: rows of a stencil spread over the banks
constexpr int COL_CELLS = COLF_X * COL_PLANE;     // 595 cells per warp tile
constexpr int COL_WARPS = 4;

template <int ORDER, bool COMPLEX>
constexpr size_t col_smem_bytes() {
  return (size_t)COL_WARPS * (sizeof(double) * COL_CELLS * (COMPLEX ? 2 : 1)
                              + sizeof(double) * 32 * 3 * ORDER + sizeof(int) * 32 * 3 * ORDER
                              + sizeof(double) * 32 * 2);
}

template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(COL_WARPS * 32)
k_assign_col(SortedView c, const int* __restrict__ key_start, long long nkeys, int nt1, int nt2,
             GridDesc g, int kind, YlmCoef yc, double scale, double* __restrict__ mesh) {
  constexpr int CW = COMPLEX ? 2 : 1;
  constexpr int LPR = ORDER * CW;                     // lanes per stencil row
  constexpr int RPP = 32 / LPR;                       // rows per instruction
  constexpr int NROW = ORDER * ORDER;
  constexpr int NPASS = (NROW + RPP - 1) / RPP;
  extern __shared__ __align__(16) unsigned char col_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tile = reinterpret_cast<double*>(col_smem) + (size_t)warp * COL_CELLS * CW;
  double* s_win = reinterpret_cast<double*>(col_smem) + (size_t)COL_WARPS * COL_CELLS * CW
    + (size_t)warp * 32 * 3 * ORDER;
  double* s_wt = reinterpret_cast<double*>(col_smem) + (size_t)COL_WARPS * COL_CELLS * CW
    + (size_t)COL_WARPS * 32 * 3 * ORDER + (size_t)warp * 32 * 2;
  int* s_idx = reinterpret_cast<int*>(reinterpret_cast<double*>(col_smem)
    + (size_t)COL_WARPS * COL_CELLS * CW + (size_t)COL_WARPS * 32 * 3 * ORDER
    + (size_t)COL_WARPS * 32 * 2) + (size_t)warp * 32 * 3 * ORDER;
  const int r = lane / LPR, q = lane - r * LPR;
  const int cz = COMPLEX ? (q >> 1) : q;
  const int comp = COMPLEX ? (q & 1) : 0;

  for (long long key = (long long)blockIdx.x * COL_WARPS + warp; key < nkeys;
       key += (long long)gridDim.x * COL_WARPS) {
    const int seg_b = key_start[key], seg_e = key_start[key + 1];
    if (seg_b == seg_e) continue;                      // warp-uniform
    // key = tile * KEYS_PER_TILE + cx * COLS_Y + cy (sort_key)
    const int col = (int)(key % KEYS_PER_TILE);
    const long long tl = key / KEYS_PER_TILE;
    const int tk = (int)(tl % nt2), tj = (int)((tl / nt2) % nt1), ti = (int)(tl / ((long long)nt2 * nt1));
    const int o0 = ti * TILE_X + (col / COLS_Y) * COL_W - 1;
    const int o1 = tj * TILE_Y + (col % COLS_Y) * COL_W - 1;
    const int o2 = tk * TILE_Z - 1;
    __syncwarp();
    for (int t = lane; t < COL_CELLS * CW; t += 32) tile[t] = 0.;
    for (int base = seg_b; base < seg_e; base += 32) {
      const int i = base + lane;
      __syncwarp();
      if (i < seg_e) {
        int ijk[ORDER]; double win[ORDER];
        const double4 p = c.p4[i];
        bool ok = true;
        window_1d<ORDER>(p.x, g.n[0], ijk, win);
#pragma unroll
        for (int t = 0; t < ORDER; t++) {
          const int l = local_index(ijk[t], o0, g.n[0], COLF_X);
          ok = ok && l >= 0; s_win[lane * 3 * ORDER + t] = win[t]; s_idx[lane * 3 * ORDER + t] = l;
        }
        window_1d<ORDER>(p.y, g.n[1], ijk, win);
#pragma unroll
        for (int t = 0; t < ORDER; t++) {
          const int l = local_index(ijk[t], o1, g.n[1], COLF_Y);
          ok = ok && l >= 0; s_win[lane * 3 * ORDER + ORDER + t] = win[t]; s_idx[lane * 3 * ORDER + ORDER + t] = l;
        }
        window_1d<ORDER>(p.z, g.n[2], ijk, win);
#pragma unroll
        for (int t = 0; t < ORDER; t++) {
          const int l = local_index(ijk[t], o2, g.n[2], COLF_Z);
          ok = ok && l >= 0; s_win[lane * 3 * ORDER + 2 * ORDER + t] = win[t]; s_idx[lane * 3 * ORDER + 2 * ORDER + t] = l;
        }
        const cplx wt = particle_weight(c, i, p, kind, yc);
        double bre = __dmul_rn(scale, wt.re), bim = __dmul_rn(scale, wt.im);
        if (!ok) {
          // Position outside [0, L): the reference's index arithmetic and bounds guard,
          // straight to global memory; nothing for the tile.
          scatter_one<ORDER, COMPLEX>(p, bre, COMPLEX ? bim : 0., g, mesh);
          bre = 0.; bim = 0.;
#pragma unroll
          for (int t = 0; t < 3 * ORDER; t++) s_idx[lane * 3 * ORDER + t] = -1;   // skipped below
        }
        s_wt[lane * 2] = bre; s_wt[lane * 2 + 1] = bim;
      }
      __syncwarp();
      const int count = min(32, seg_e - base);
      for (int sdx = 0; sdx < count; sdx++) {
        if (r < RPP) {
          const double* w = s_win + sdx * 3 * ORDER;
          const int* id = s_idx + sdx * 3 * ORDER;
          const double bw = s_wt[sdx * 2 + comp];
          const double wz = w[2 * ORDER + cz];
          const int lz = id[2 * ORDER + cz];
#pragma unroll
          for (int pass = 0; pass < NPASS; pass++) {
            const int row = r + RPP * pass;
            if (row < NROW && lz >= 0) {
              const int a = row / ORDER, b = row - a * ORDER;
              const int cell = id[a] * COL_PLANE + id[ORDER + b] * COL_ROW + lz;
              const double v = __dmul_rn(__dmul_rn(__dmul_rn(bw, w[a]), w[ORDER + b]), wz);
              tile[cell * CW + comp] += v;
            }
          }
        }
        __syncwarp();
      }
    }
    // Flush the footprint: lanes over (ly, lz), planes along x.
    for (int pq = lane; pq < COLF_Y * COLF_Z; pq += 32) {
      const int ly = pq / COLF_Z, lz = pq - ly * COLF_Z;
      int gy = o1 + ly, gz = o2 + lz;
      gy += (gy < 0) ? g.n[1] : 0; gy -= (gy >= g.n[1]) ? g.n[1] : 0;
      gz += (gz < 0) ? g.n[2] : 0; gz -= (gz >= g.n[2]) ? g.n[2] : 0;
#pragma unroll
      for (int lx = 0; lx < COLF_X; lx++) {
        int gx = o0 + lx;
        gx += (gx < 0) ? g.n[0] : 0; gx -= (gx >= g.n[0]) ? g.n[0] : 0;
        const int cell = lx * COL_PLANE + ly * COL_ROW + lz;
        const long long gid = ((long long)gx * g.n[1] + gy) * g.n[2] + gz;
        const double vre = tile[cell * CW];
        if (COMPLEX) {
          const double vim = tile[cell * CW + 1];
          if (vre != 0. || vim != 0.) { atomicAdd(&mesh[2 * gid], vre); atomicAdd(&mesh[2 * gid + 1], vim); }
        } else if (vre != 0.) {
          atomicAdd(&mesh[gid], vre);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------
// Deterministic assignment: one thread per output cell, ordered gather.
// ---------------------------------------------------------------------

constexpr int DET_CAP = 96;   // candidates buffered per cell before fallback

template <int ORDER>
__device__ __forceinline__ bool contribution(
  const SortedView& c, int pid /* slot in the sorted view */, const GridDesc& g, int shifted,
  int ci, int cj, int ck, double& wprod_x, double& wprod_y, double& wprod_z
) {
  int ijk[ORDER]; double win[ORDER];
  bool hit;
  const double4 p = c.p4[pid];
  window_1d<ORDER>(shift_loc(p.x, g.n[0], shifted), g.n[0], ijk, win);
  hit = false;
#pragma unroll
  for (int a = 0; a < ORDER; a++) if (ijk[a] == ci) { wprod_x = win[a]; hit = true; }
  if (!hit) return false;
  window_1d<ORDER>(shift_loc(p.y, g.n[1], shifted), g.n[1], ijk, win);
  hit = false;
#pragma unroll
  for (int a = 0; a < ORDER; a++) if (ijk[a] == cj) { wprod_y = win[a]; hit = true; }
  if (!hit) return false;
  window_1d<ORDER>(shift_loc(p.z, g.n[2], shifted), g.n[2], ijk, win);
  hit = false;
#pragma unroll
  for (int a = 0; a < ORDER; a++) if (ijk[a] == ck) { wprod_z = win[a]; hit = true; }
  return hit;
}

template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(128)
k_assign_gather(SortedView c, const int* __restrict__ order,
                const int* __restrict__ cell_start, GridDesc g, int shifted,
                int kind, YlmCoef yc, double scale, double pre /* 1/vol_cell or 1 */,
                int accumulate, double* __restrict__ mesh) {
  // Home cells whose particles can reach output index q along an axis:
  // q - LO .. q + HI (periodic).  PCS/TSC: home in {q-2..q+1}; CIC/NGP: {q-1, q}.
  constexpr int LO = (ORDER >= 3) ? 2 : 1;
  constexpr int HI = (ORDER >= 3) ? 1 : 0;
  constexpr int SPAN = LO + HI + 1;
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (gid >= g.nmesh) return;
  const int ck = (int)(gid % g.n[2]);
  const int cj = (int)((gid / g.n[2]) % g.n[1]);
  const int ci = (int)(gid / ((long long)g.n[2] * g.n[1]));

  int cand_pid[DET_CAP];
  double cand_re[DET_CAP];
  double cand_im[COMPLEX ? DET_CAP : 1];
  int ncand = 0;
  bool overflow = false;

  auto value_of = [&](int slot, double wx, double wy, double wz, double& vre, double& vim) {
    cplx wt = particle_weight(c, slot, c.p4[slot], kind, yc);
    // ((((inv_vol_cell * w) * Wx) * Wy) * Wz), S/field.cpp:1044-1048.
    double bre = __dmul_rn(pre, wt.re);
    if (scale != 1.) bre = __dmul_rn(bre, scale);
    vre = __dmul_rn(__dmul_rn(__dmul_rn(bre, wx), wy), wz);
    if (COMPLEX) {
      double bim = __dmul_rn(pre, wt.im);
      if (scale != 1.) bim = __dmul_rn(bim, scale);
      vim = __dmul_rn(__dmul_rn(__dmul_rn(bim, wx), wy), wz);
    } else {
      vim = 0.;
    }
  };

  for (int da = 0; da < SPAN && !overflow; da++) {
    int hi_ = ci - LO + da; hi_ = (hi_ % g.n[0] + g.n[0]) % g.n[0];
    for (int db = 0; db < SPAN && !overflow; db++) {
      int hj = cj - LO + db; hj = (hj % g.n[1] + g.n[1]) % g.n[1];
      for (int dc = 0; dc < SPAN; dc++) {
        int hk = ck - LO + dc; hk = (hk % g.n[2] + g.n[2]) % g.n[2];
        const long long hcell = ((long long)hi_ * g.n[1] + hj) * g.n[2] + hk;
        const int b = cell_start[hcell], e = cell_start[hcell + 1];
        for (int s = b; s < e; s++) {
          const int pid = order[s];
          double wx, wy, wz;
          if (!contribution<ORDER>(c, s, g, shifted, ci, cj, ck, wx, wy, wz)) continue;
          if (ncand >= DET_CAP) { overflow = true; break; }
          double vre, vim;
          value_of(s, wx, wy, wz, vre, vim);
          // Insertion keeps candidates ascending in particle id.
          int q = ncand - 1;
          while (q >= 0 && cand_pid[q] > pid) {
            cand_pid[q + 1] = cand_pid[q];
            cand_re[q + 1] = cand_re[q];
            if (COMPLEX) cand_im[q + 1] = cand_im[q];
            q--;
          }
          cand_pid[q + 1] = pid; cand_re[q + 1] = vre;
          if (COMPLEX) cand_im[q + 1] = vim;
          ncand++;
        }
        if (overflow) break;
      }
    }
  }

  double acc_re = 0., acc_im = 0.;
  if (accumulate) {
    acc_re = COMPLEX ? mesh[2 * gid] : mesh[gid];
    if (COMPLEX) acc_im = mesh[2 * gid + 1];
  }
  if (!overflow) {
    for (int q = 0; q < ncand; q++) {
      acc_re = __dadd_rn(acc_re, cand_re[q]);
      if (COMPLEX) acc_im = __dadd_rn(acc_im, cand_im[q]);
    }
  } else {
    // Dense cell: repeated selection of the next particle id (no storage).
    int last = -1;
    while (true) {
      int best = 0x7fffffff, best_slot = 0; double bx = 0., by = 0., bz = 0.;
      for (int da = 0; da < SPAN; da++) {
        int hi_ = ci - LO + da; hi_ = (hi_ % g.n[0] + g.n[0]) % g.n[0];
        for (int db = 0; db < SPAN; db++) {
          int hj = cj - LO + db; hj = (hj % g.n[1] + g.n[1]) % g.n[1];
          for (int dc = 0; dc < SPAN; dc++) {
            int hk = ck - LO + dc; hk = (hk % g.n[2] + g.n[2]) % g.n[2];
            const long long hcell = ((long long)hi_ * g.n[1] + hj) * g.n[2] + hk;
            const int b = cell_start[hcell], e = cell_start[hcell + 1];
            for (int s = b; s < e; s++) {
              const int pid = order[s];
              if (pid <= last || pid >= best) continue;
              double wx, wy, wz;
              if (!contribution<ORDER>(c, s, g, shifted, ci, cj, ck, wx, wy, wz)) continue;
              best = pid; best_slot = s; bx = wx; by = wy; bz = wz;
            }
          }
        }
      }
      if (best == 0x7fffffff) break;
      double vre, vim;
      value_of(best_slot, bx, by, bz, vre, vim);
      acc_re = __dadd_rn(acc_re, vre);
      if (COMPLEX) acc_im = __dadd_rn(acc_im, vim);
      last = best;
    }
  }
  if (COMPLEX) { mesh[2 * gid] = acc_re; mesh[2 * gid + 1] = acc_im; }
  else mesh[gid] = acc_re;
}

// ---------------------------------------------------------------------
// Catalogue sums.
// ---------------------------------------------------------------------

__global__ void k_cat_sum(CatView c, int kind, int L, int M,
                          double* __restrict__ partial /* [gridDim][2] */) {
  __shared__ double sm[32];
  double re = 0., im = 0.;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c.n;
       i += (long long)gridDim.x * blockDim.x) {
    cplx wv = particle_weight(c, i, kind, L, M);
    re += wv.re; im += wv.im;
  }
  double sre = block_sum(re, sm);
  double sim = block_sum(im, sm);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sre; partial[2 * blockIdx.x + 1] = sim; }
}

__global__ void k_sum_partials(const double* __restrict__ partial, int nblocks, int width,
                               double* __restrict__ out) {
  // Fixed-order second stage: thread t sums column t over blocks.
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= width) return;
  double s = 0.;
  for (int b = 0; b < nblocks; b++) s += partial[(long long)b * width + t];
  out[t] = s;
}

#include "trvb_assign_own.cuh"

CatView view_of(const trvb_cat* cat) {
  CatView v;
  v.x = cat->x; v.y = cat->y; v.z = cat->z; v.w = cat->w;
  v.lx = cat->los; v.ly = cat->los ? cat->los + cat->n : nullptr;
  v.lz = cat->los ? cat->los + 2 * cat->n : nullptr;
  v.cw = cat->cw;
  v.n = cat->n;
  return v;
}

__global__ void k_los_to_soa(const double* __restrict__ los, long long n, double* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    out[i] = los[3 * i]; out[n + i] = los[3 * i + 1]; out[2 * n + i] = los[3 * i + 2];
  }
}

SortedView sorted_view_of(const trvb_cat* cat) {
  SortedView v;
  v.p4 = cat->s4;
  v.lx = cat->slos; v.ly = cat->slos ? cat->slos + cat->n : nullptr;
  v.lz = cat->slos ? cat->slos + 2 * cat->n : nullptr;
  v.cw = cat->scw;
  v.n = cat->n;
  return v;
}

int alloc_sorted(trvb_ctx* ctx, trvb_cat* cat) {
  const size_t nb = sizeof(double) * (size_t)cat->n;
  if (!cat->s4) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->s4, 4 * nb));
  if (cat->los && !cat->slos) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->slos, 3 * nb));
  if (cat->cw && !cat->scw) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->scw, 2 * nb));
  return 0;
}

// Gather what the scatter did not already place: with_records == false skips
// the packed {x, y, z, w} records.
int gather_sorted(trvb_ctx* ctx, trvb_cat* cat, bool with_records) {
  int st = alloc_sorted(ctx, cat);
  if (st) return st;
  cat->scw_valid = cat->cw != nullptr;
  if (!with_records && !cat->slos && !cat->scw) return 0;
  const int blocks = (int)std::min<long long>(div_up(cat->n, 256), (long long)ctx->num_sms * 16);
  SortDesc d;
  for (int a = 0; a < 3; a++) { d.n[a] = ctx->g.n[a]; d.L[a] = ctx->g.L[a]; d.nk[a] = 0; }
  d.shifted = 0; d.by_cell = 0;
  k_gather_sorted<<<blocks, 256, 0, ctx->stream>>>(view_of(cat), d, cat->order,
                                                  with_records ? cat->s4 : nullptr,
                                                  cat->slos, cat->scw);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// Counting sort of the particles of `cv` by sort key, enqueued on the context's
// stream: histogram, exclusive scan (offsets, nkeys + 1 entries), scatter of the ids
// (`order`, may be null) and of the packed records (`s4`, may be null).
int enqueue_counting_sort(trvb_ctx* ctx, const CatView& cv, const SortDesc& d, long long nkeys,
                          int* offsets, int* cursor, int* chunk_sums, int* order, double4* s4) {
  TRVB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int) * (size_t)(nkeys + 1), ctx->stream));
  const int threads = 256;
  const int blocks = (int)std::min<long long>(div_up(cv.n, threads), (long long)ctx->num_sms * 16);
  k_sort_count<<<blocks, threads, 0, ctx->stream>>>(cv, d, offsets);
  TRVB_LAUNCH_CHECK();
  const long long nscan = nkeys + 1;
  const int nchunks = div_up(nscan, SCAN_CHUNK);
  k_scan_chunk_sums<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(offsets, nscan, chunk_sums);
  TRVB_LAUNCH_CHECK();
  k_scan_chunk_offsets<<<1, 1024, 0, ctx->stream>>>(chunk_sums, nchunks);
  TRVB_LAUNCH_CHECK();
  k_scan_apply<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(offsets, nscan, chunk_sums);
  TRVB_LAUNCH_CHECK();
  TRVB_CUDA(cudaMemcpyAsync(cursor, offsets, sizeof(int) * (size_t)nkeys,
                            cudaMemcpyDeviceToDevice, ctx->stream));
  k_sort_scatter<<<blocks, threads, 0, ctx->stream>>>(cv, d, cursor, order, s4);
  TRVB_LAUNCH_CHECK();
  return 0;
}

// by_cell == 0: throughput order ((tile, column) of the UNSHIFTED home cell, plus
// the key offsets k_assign_tile needs; for the shifted shadow mesh it only
// provides locality).
// by_cell == 1: home cell of the (possibly shifted) mesh, ascending particle id
// inside a cell, plus the cell offsets the ordered gather needs.
int ensure_sorted(trvb_ctx* ctx, trvb_cat* cat, int shifted, int by_cell) {
  const GridDesc& g = ctx->g;
  if (!by_cell) shifted = 0;
  bool valid = cat->order != nullptr && cat->sort_kind == by_cell
    && cat->sort_shifted == shifted;
  for (int a = 0; a < 3; a++) {
    valid = valid && cat->sort_n[a] == g.n[a] && cat->sort_L[a] == g.L[a];
  }
  // The id permutation is only written when something is gathered through it (lines
  // of sight, custom weights, the cell-ordered mode): 4-byte writes to random sectors
  // cost the scatter as much as the 32-byte records.
  const bool need_order = by_cell || cat->los != nullptr || cat->cw != nullptr;
  if (need_order && !cat->order_valid) valid = false;
  if (valid) {
    if (cat->cw && !cat->scw_valid) return gather_sorted(ctx, cat, false);
    return 0;
  }
  SortDesc d;
  for (int a = 0; a < 3; a++) {
    d.n[a] = g.n[a]; d.L[a] = g.L[a];
    const int tile = (a == 0) ? TILE_X : (a == 1) ? TILE_Y : TILE_Z;
    d.nk[a] = by_cell ? g.n[a] : (g.n[a] + tile - 1) / tile;
  }
  d.shifted = shifted; d.by_cell = by_cell;
  const long long nkeys = (long long)d.nk[0] * d.nk[1] * d.nk[2] * (by_cell ? 1 : KEYS_PER_TILE);
  TRVB_REQUIRE(nkeys < 2147483647LL, "mesh too large for int sort keys");
  TRVB_REQUIRE(cat->n < 2147483647LL, "catalogue too large for int indices");
  if (!cat->order) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->order, sizeof(int) * (size_t)cat->n));
  { int st = alloc_sorted(ctx, cat); if (st) return st; }
  if (cat->cell_start) {
    TRVB_CUDA(trvb_dev_free_raw(ctx, cat->cell_start)); cat->cell_start = nullptr;
  }
  int* offsets = nullptr;   // nkeys + 1
  int* cursor = nullptr;    // nkeys
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&offsets, sizeof(int) * (size_t)(nkeys + 1)));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cursor, sizeof(int) * (size_t)nkeys));
  const int nchunks = div_up(nkeys + 1, SCAN_CHUNK);
  int* chunk_sums = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&chunk_sums, sizeof(int) * (size_t)nchunks));
  CatView cv = view_of(cat);
  const int threads = 256;
  // Throughput order: the scatter places the packed records itself.  Cell
  // order: ids are sorted inside each cell first, records gathered after.
  { int st = enqueue_counting_sort(ctx, cv, d, nkeys, offsets, cursor, chunk_sums,
                                   need_order ? cat->order : nullptr,
                                   by_cell ? nullptr : cat->s4);
    if (st) return st; }
  cat->order_valid = need_order;
  cat->chunked = false;
  if (by_cell) {
    const int sb = (int)std::min<long long>(div_up(nkeys, threads), (long long)ctx->num_sms * 32);
    k_sort_segments<<<sb, threads, 0, ctx->stream>>>(cursor, nkeys, cat->order);
    TRVB_LAUNCH_CHECK();
  }
  cat->cell_start = offsets;   // key offsets: cells (deterministic) or (tile, column)
  TRVB_CUDA(trvb_dev_free_raw(ctx, cursor));   // stream-ordered reuse: no sync
  TRVB_CUDA(trvb_dev_free_raw(ctx, chunk_sums));
  for (int a = 0; a < 3; a++) { cat->sort_n[a] = g.n[a]; cat->sort_L[a] = g.L[a]; }
  cat->sort_shifted = shifted; cat->sort_kind = by_cell;
  return gather_sorted(ctx, cat, by_cell != 0);
}

// Warp-cooperative form of the ordered gather (TSC/PCS).  A warp owns a tile of
// GT_X x GT_Y x GT_Z = 32 output cells (lane = cell).  The particles that can reach
// the tile live in the (GT + 3)^3-shaped union of home cells; lane l keeps the list
// cursors of union cells l, l + 32, ... in registers, and the warp merges those
// (already id-ordered) lists: every round takes the smallest pending particle id
// (__reduce_min_sync), all lanes evaluate that ONE particle against their own cell
// and add it -- so each cell accumulates in ascending particle id, the reference's
// single-threaded order, with no per-cell candidate buffer, no per-lane sort and
// no divergence (the thread-per-cell kernel spent its time in 64 mostly-empty,
// divergent list visits per cell: 740 ms for 1e7 particles on 512^3, PCS).
constexpr int GT_X = 2, GT_Y = 4, GT_Z = 4;
constexpr int DT_X = 4, DT_Y = 4, DT_Z = 8;     // k_assign_gather_tile (multiples of GT)

template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(128)
k_assign_gather_warp(SortedView c, const int* __restrict__ order,
                     const int* __restrict__ cell_start, GridDesc g, int shifted,
                     int kind, YlmCoef yc, double scale, double pre /* 1/vol_cell or 1 */,
                     int accumulate, int nt1, int nt2, long long ntiles,
                     const int* __restrict__ big_tiles, const int* __restrict__ big_count,
                     int bt1, int bt2, double* __restrict__ mesh) {
  constexpr int LO = 2, SPAN = 4;                       // homes q - 2 .. q + 1
  constexpr int UX = GT_X + SPAN - 1, UY = GT_Y + SPAN - 1, UZ = GT_Z + SPAN - 1;
  constexpr int NU = UX * UY * UZ;
  constexpr int PER = (NU + 31) / 32;
  const int lane = threadIdx.x & 31;
  // Either every tile of the mesh, or (big_tiles != null) the 2 x 1 x 2 tiles of each
  // DT-sized tile that k_assign_gather_tile left behind (too many candidates to sort).
  constexpr int SUB = (DT_X / GT_X) * (DT_Y / GT_Y) * (DT_Z / GT_Z);
  const long long nwork = big_tiles ? (long long)(*big_count) * SUB : ntiles;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long work = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
       work < nwork; work += nwarps) {
  int ti, tj, tk;
  if (big_tiles) {
    const int big = big_tiles[work / SUB], sub = (int)(work % SUB);
    const int bk = big % bt2, bj = (big / bt2) % bt1, bi = big / (bt2 * bt1);
    ti = bi * (DT_X / GT_X) + sub / ((DT_Y / GT_Y) * (DT_Z / GT_Z));
    tj = bj * (DT_Y / GT_Y) + (sub / (DT_Z / GT_Z)) % (DT_Y / GT_Y);
    tk = bk * (DT_Z / GT_Z) + sub % (DT_Z / GT_Z);
    if (ti >= (g.n[0] + GT_X - 1) / GT_X || tj >= nt1 || tk >= nt2) continue;
  } else {
    tk = (int)(work % nt2); tj = (int)((work / nt2) % nt1); ti = (int)(work / ((long long)nt2 * nt1));
  }
  const int ci = ti * GT_X + (lane >> 4), cj = tj * GT_Y + ((lane >> 2) & 3), ck = tk * GT_Z + (lane & 3);
  const bool valid = ci < g.n[0] && cj < g.n[1] && ck < g.n[2];

  // List cursors of this lane's union cells.
  int cur[PER], end[PER], head[PER];
#pragma unroll
  for (int t = 0; t < PER; t++) {
    const int u = lane + 32 * t;
    cur[t] = 0; end[t] = 0; head[t] = 0x7fffffff;
    if (u < NU) {
      const int uz = u % UZ, uy = (u / UZ) % UY, ux = u / (UZ * UY);
      int hx = ti * GT_X - LO + ux, hy = tj * GT_Y - LO + uy, hz = tk * GT_Z - LO + uz;
      hx = (hx % g.n[0] + g.n[0]) % g.n[0];
      hy = (hy % g.n[1] + g.n[1]) % g.n[1];
      hz = (hz % g.n[2] + g.n[2]) % g.n[2];
      const long long hcell = ((long long)hx * g.n[1] + hy) * g.n[2] + hz;
      cur[t] = cell_start[hcell]; end[t] = cell_start[hcell + 1];
    }
  }
#pragma unroll
  for (int t = 0; t < PER; t++) if (cur[t] < end[t]) head[t] = order[cur[t]];

  double acc_re = 0., acc_im = 0.;
  const long long gid = valid ? ((long long)ci * g.n[1] + cj) * g.n[2] + ck : 0;
  if (accumulate && valid) {
    acc_re = COMPLEX ? mesh[2 * gid] : mesh[gid];
    if (COMPLEX) acc_im = mesh[2 * gid + 1];
  }

  while (true) {
    int mine = head[0];
#pragma unroll
    for (int t = 1; t < PER; t++) mine = min(mine, head[t]);
    const int next = __reduce_min_sync(0xffffffffu, mine);
    if (next == 0x7fffffff) break;
    // Particle ids are unique: exactly one lane holds `next`, in one of its lists.
    int slot = -1;
#pragma unroll
    for (int t = 0; t < PER; t++) {
      if (head[t] == next) {
        slot = cur[t];
        cur[t]++;
        head[t] = (cur[t] < end[t]) ? order[cur[t]] : 0x7fffffff;
      }
    }
    const unsigned owner = __ballot_sync(0xffffffffu, slot >= 0);
    slot = __shfl_sync(0xffffffffu, slot, __ffs(owner) - 1);
    double wx, wy, wz;
    if (valid && contribution<ORDER>(c, slot, g, shifted, ci, cj, ck, wx, wy, wz)) {
      const cplx wt = particle_weight(c, slot, c.p4[slot], kind, yc);
      // ((((inv_vol_cell * w) * Wx) * Wy) * Wz), S/field.cpp:1044-1048.
      double bre = __dmul_rn(pre, wt.re);
      if (scale != 1.) bre = __dmul_rn(bre, scale);
      acc_re = __dadd_rn(acc_re, __dmul_rn(__dmul_rn(__dmul_rn(bre, wx), wy), wz));
      if (COMPLEX) {
        double bim = __dmul_rn(pre, wt.im);
        if (scale != 1.) bim = __dmul_rn(bim, scale);
        acc_im = __dadd_rn(acc_im, __dmul_rn(__dmul_rn(__dmul_rn(bim, wx), wy), wz));
      }
    }
  }
  if (valid) {
    if (COMPLEX) { mesh[2 * gid] = acc_re; mesh[2 * gid + 1] = acc_im; }
    else mesh[gid] = acc_re;
  }
  }
}

// Tile form of the ordered gather (TSC/PCS): a warp owns DT_X x DT_Y x DT_Z = 128 output
// cells, four z-consecutive cells per lane in registers.  The particles that can reach the
// tile -- those of the (DT + 3)^3-shaped union of home cells, ~40 on a 512^3 mesh with 1e7
// particles -- are copied to shared memory as (id, slot) pairs and SORTED by id there
// (bitonic, one warp), then walked once in that order: every lane evaluates each particle
// against its own cells and adds it, so every cell accumulates in ascending particle id,
// the reference's single-threaded order.  The merge of k_assign_gather_warp (one
// __reduce_min_sync round trip per particle and per 32 cells) becomes one sort per 128
// cells.  Tiles with more than DT_CAP candidates (clustered catalogues) are listed for
// k_assign_gather_warp.
constexpr int DT_CAP = 256;      // candidates a warp sorts in shared memory
constexpr int DT_CHUNK = 64;     // candidates whose windows are staged at a time

// Per-candidate data of the walk, evaluated once per tile by one lane each (the walk itself
// would redo it in every lane: three fp64 window evaluations per particle and lane made the
// fp64 pipe the bound): first stencil index and window values per axis, and the weight.
template <int ORDER>
struct DetStage {
  int i0[3];
  int pad;
  double win[3][ORDER];
  double w_re, w_im;
};

template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(128, 4)
k_assign_gather_tile(SortedView c, const int* __restrict__ order,
                     const int* __restrict__ cell_start, GridDesc g, int shifted,
                     int kind, YlmCoef yc, double scale, double pre /* 1/vol_cell or 1 */,
                     int accumulate, int nt1, int nt2, long long ntiles,
                     int* __restrict__ big_tiles, int* __restrict__ big_count,
                     double* __restrict__ mesh) {
  // homes that reach output index q: q - 2 .. q + 1 (TSC, PCS), q - 1 .. q (NGP, CIC)
  constexpr int LO = (ORDER >= 3) ? 2 : 1, SPAN = (ORDER >= 3) ? 4 : 2;
  constexpr int UX = DT_X + SPAN - 1, UY = DT_Y + SPAN - 1, UZ = DT_Z + SPAN - 1;
  constexpr int ZPL = 4;                                // cells per lane, consecutive in z
  static_assert(DT_X * DT_Y * (DT_Z / ZPL) == 32, "one lane per column segment");
  __shared__ int s_id[4][DT_CAP];
  __shared__ int s_slot[4][DT_CAP];
  __shared__ DetStage<ORDER> s_stage[4][DT_CHUNK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long tile = blockIdx.x * (long long)(blockDim.x >> 5) + warp;
  if (tile >= ntiles) return;                           // whole warp
  const int tk = (int)(tile % nt2), tj = (int)((tile / nt2) % nt1), ti = (int)(tile / ((long long)nt2 * nt1));
  int* ids = s_id[warp];
  int* slots = s_slot[warp];
  DetStage<ORDER>* stage = s_stage[warp];

  // 1. Candidates per z-row of the union (cells consecutive in z are consecutive in the
  // cell-sorted order: a row is one slot range, two where it wraps around the box), and
  // where this lane's go in the list.
  constexpr int NROW = UX * UY * 2;
  constexpr int PER_ROW = (NROW + 31) / 32;
  int beg[PER_ROW], cnt[PER_ROW];
  int mine = 0;
  {
    const int z0 = tk * DT_Z - LO;                      // first z of the union, unwrapped
#pragma unroll
    for (int t = 0; t < PER_ROW; t++) {
      const int u = lane + 32 * t;
      beg[t] = 0; cnt[t] = 0;
      if (u < NROW) {
        const int seg = u & 1, uy = (u >> 1) % UY, ux = (u >> 1) / UY;
        int hx = ti * DT_X - LO + ux, hy = tj * DT_Y - LO + uy;
        hx = (hx % g.n[0] + g.n[0]) % g.n[0];
        hy = (hy % g.n[1] + g.n[1]) % g.n[1];
        // segment 0: the part of [z0, z0 + UZ) inside [0, n2); segment 1: the wrapped rest
        int za, zb;
        if (seg == 0) { za = max(z0, 0); zb = min(z0 + UZ, g.n[2]); }
        else if (z0 < 0) { za = g.n[2] + z0; zb = g.n[2]; }
        else { za = 0; zb = z0 + UZ - g.n[2]; }
        if (zb > za) {
          const long long row = ((long long)hx * g.n[1] + hy) * g.n[2];
          beg[t] = cell_start[row + za]; cnt[t] = cell_start[row + zb] - beg[t];
        }
      }
      mine += cnt[t];
    }
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total > DT_CAP) {
    if (lane == 0) big_tiles[atomicAdd(big_count, 1)] = (int)tile;
    return;
  }
  // 2. (id, slot) pairs, padded to a power of two.
  int npad = 32;
  while (npad < total) npad <<= 1;
  {
    int at = incl - mine;
#pragma unroll
    for (int t = 0; t < PER_ROW; t++) {
      for (int q = 0; q < cnt[t]; q++) { ids[at] = order[beg[t] + q]; slots[at] = beg[t] + q; at++; }
    }
    for (int i = total + lane; i < npad; i += 32) { ids[i] = 0x7fffffff; slots[i] = -1; }
  }
  __syncwarp();
  // 3. Bitonic sort by particle id (ids are unique).
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < npad; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const int a = ids[i], b = ids[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            ids[i] = b; ids[l] = a;
            const int sa = slots[i]; slots[i] = slots[l]; slots[l] = sa;
          }
        }
      }
      __syncwarp();
    }
  }
  // 4. Walk in id order, DT_CHUNK candidates at a time; lane = (x, y, z-quad) of the tile.
  const int ci = ti * DT_X + (lane >> 3), cj = tj * DT_Y + ((lane >> 1) & 3);
  const int ck0 = tk * DT_Z + (lane & 1) * ZPL;
  const bool col_valid = ci < g.n[0] && cj < g.n[1];
  double acc_re[ZPL], acc_im[ZPL];
  long long gid[ZPL];
#pragma unroll
  for (int m = 0; m < ZPL; m++) {
    const bool valid = col_valid && ck0 + m < g.n[2];
    gid[m] = valid ? ((long long)ci * g.n[1] + cj) * g.n[2] + ck0 + m : -1;
    acc_re[m] = 0.; acc_im[m] = 0.;
    if (accumulate && valid) {
      acc_re[m] = COMPLEX ? mesh[2 * gid[m]] : mesh[gid[m]];
      if (COMPLEX) acc_im[m] = mesh[2 * gid[m] + 1];
    }
  }
  for (int c0 = 0; c0 < total; c0 += DT_CHUNK) {
    const int nc = min(DT_CHUNK, total - c0);
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {               // one lane per candidate
      const int slot = slots[c0 + i];
      const double4 p = c.p4[slot];
      DetStage<ORDER>& st = stage[i];
      int ijk[ORDER]; double win[ORDER];
      window_1d<ORDER>(shift_loc(p.x, g.n[0], shifted), g.n[0], ijk, win);
      st.i0[0] = ijk[0];
#pragma unroll
      for (int a = 0; a < ORDER; a++) st.win[0][a] = win[a];
      window_1d<ORDER>(shift_loc(p.y, g.n[1], shifted), g.n[1], ijk, win);
      st.i0[1] = ijk[0];
#pragma unroll
      for (int a = 0; a < ORDER; a++) st.win[1][a] = win[a];
      window_1d<ORDER>(shift_loc(p.z, g.n[2], shifted), g.n[2], ijk, win);
      st.i0[2] = ijk[0];
#pragma unroll
      for (int a = 0; a < ORDER; a++) st.win[2][a] = win[a];
      const cplx wt = particle_weight(c, slot, p, kind, yc);
      // ((((inv_vol_cell * w) * Wx) * Wy) * Wz), S/field.cpp:1044-1048.
      double bre = __dmul_rn(pre, wt.re);
      if (scale != 1.) bre = __dmul_rn(bre, scale);
      double bim = 0.;
      if (COMPLEX) {
        bim = __dmul_rn(pre, wt.im);
        if (scale != 1.) bim = __dmul_rn(bim, scale);
      }
      st.w_re = bre; st.w_im = bim;
    }
    __syncwarp();
    if (!col_valid) continue;
    for (int i = 0; i < nc; i++) {
      const DetStage<ORDER>& st = stage[i];
      // the stencil of an axis is ORDER consecutive indices (periodic) from i0
      int ax = ci - st.i0[0]; ax += (ax < 0) ? g.n[0] : 0;
      if (ax >= ORDER) continue;
      int ay = cj - st.i0[1]; ay += (ay < 0) ? g.n[1] : 0;
      if (ay >= ORDER) continue;
      int az = ck0 - st.i0[2]; az += (az < 0) ? g.n[2] : 0;   // of this lane's first cell
      // cells ck0 .. ck0 + 3 -> window slots az .. az + 3 (mod n2)
      if (az >= ORDER && az + (ZPL - 1) < g.n[2]) continue;
      const double bxy_re = __dmul_rn(__dmul_rn(st.w_re, st.win[0][ax]), st.win[1][ay]);
      const double bxy_im = COMPLEX ? __dmul_rn(__dmul_rn(st.w_im, st.win[0][ax]), st.win[1][ay]) : 0.;
#pragma unroll
      for (int m = 0; m < ZPL; m++) {
        int a = az + m; a -= (a >= g.n[2]) ? g.n[2] : 0;
        if (a < ORDER && gid[m] >= 0) {
          const double wz = st.win[2][a];
          acc_re[m] = __dadd_rn(acc_re[m], __dmul_rn(bxy_re, wz));
          if (COMPLEX) acc_im[m] = __dadd_rn(acc_im[m], __dmul_rn(bxy_im, wz));
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < ZPL; m++) {
    if (gid[m] >= 0) {
      if (COMPLEX) { mesh[2 * gid[m]] = acc_re[m]; mesh[2 * gid[m] + 1] = acc_im[m]; }
      else mesh[gid[m]] = acc_re[m];
    }
  }
}

// Particle-wise scatter of a sorted view: warp-cooperative for TSC/PCS, thread per
// particle for NGP/CIC.
// ---------------------------------------------------------------------
// Tile-owned assignment: owner-sorted copies and launch (trvb_assign_own.cuh).
// ---------------------------------------------------------------------

void free_own_sorted(trvb_ctx* ctx, trvb_cat* cat) {
  trvb_dev_free_raw(ctx, cat->own_rec); trvb_dev_free_raw(ctx, cat->own_pk);
  trvb_dev_free_raw(ctx, cat->own_src); trvb_dev_free_raw(ctx, cat->own_offsets);
  trvb_dev_free_raw(ctx, cat->own_irr);
  cat->own_rec = nullptr; cat->own_pk = nullptr; cat->own_src = nullptr;
  cat->own_offsets = nullptr; cat->own_irr = nullptr;
  cat->own_total = 0; cat->own_nirr = 0; cat->own_shifted = -1; cat->own_order = 0;
}

OwnDesc own_desc(const GridDesc& g, int shifted) {
  OwnDesc d;
  for (int a = 0; a < 3; a++) { d.n[a] = g.n[a]; d.L[a] = g.L[a]; }
  d.shifted = shifted;
  d.nch[0] = (g.n[0] + OWN_PX - 1) / OWN_PX;
  d.nch[1] = (g.n[1] + OWN_PY - 1) / OWN_PY;
  d.nch[2] = (g.n[2] + OWN_ZS - 1) / OWN_ZS;
  d.kpt = OWN_ZS + g.order - 1;
  d.pk_in_w = 0;
  return d;
}

template <int ORDER>
int ensure_own_sorted(trvb_ctx* ctx, trvb_cat* cat, int shifted) {
  const GridDesc& g = ctx->g;
  const bool need_src = cat->los != nullptr || cat->cw != nullptr;
  bool valid = cat->own_rec != nullptr && cat->own_shifted == shifted && cat->own_order == ORDER
    && (!need_src || cat->own_src != nullptr);
  for (int a = 0; a < 3; a++) valid = valid && cat->own_n[a] == g.n[a] && cat->own_L[a] == g.L[a];
  if (valid) return 0;
  free_own_sorted(ctx, cat);
  OwnDesc d = own_desc(g, shifted);
  d.pk_in_w = cat->w == nullptr;   // unit weights: one 32-byte sector per copy, no key array
  const long long ntasks = (long long)d.nch[0] * d.nch[1] * d.nch[2];
  const long long nkeys = ntasks * d.kpt;
  TRVB_REQUIRE(nkeys < 2147483647LL, "mesh too large for int sort keys");
  TRVB_REQUIRE(cat->n < 2147483647LL / 8, "catalogue too large for int indices");
  int* offsets = nullptr; int* cursor = nullptr; int* chunk_sums = nullptr; int* misc = nullptr;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&offsets, sizeof(int) * (size_t)(nkeys + 1)));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cursor, sizeof(int) * (size_t)nkeys));
  const long long nscan = nkeys + 1;
  const int nchunks = div_up(nscan, SCAN_CHUNK);
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&chunk_sums, sizeof(int) * (size_t)nchunks));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&misc, sizeof(int) * 4));
  TRVB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int) * (size_t)nscan, ctx->stream));
  TRVB_CUDA(cudaMemsetAsync(misc, 0, sizeof(int) * 4, ctx->stream));
  const CatView cv = view_of(cat);
  const int threads = 256;
  const int blocks = (int)std::min<long long>(div_up(cat->n, threads), (long long)ctx->num_sms * 16);
  k_own_sort<ORDER><<<blocks, threads, 0, ctx->stream>>>(cv, d, offsets, nullptr, misc, nullptr,
                                                         nullptr, nullptr, nullptr);
  TRVB_LAUNCH_CHECK();
  k_scan_chunk_sums<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(offsets, nscan, chunk_sums);
  TRVB_LAUNCH_CHECK();
  k_scan_chunk_offsets<<<1, 1024, 0, ctx->stream>>>(chunk_sums, nchunks);
  TRVB_LAUNCH_CHECK();
  k_scan_apply<<<nchunks, SCAN_THREADS, 0, ctx->stream>>>(offsets, nscan, chunk_sums);
  TRVB_LAUNCH_CHECK();
  // The number of copies (1 to 8 per particle) sizes the lists: one small read-back.
  int h_total = 0, h_nirr = 0;
  TRVB_CUDA(cudaMemcpyAsync(&h_total, offsets + nkeys, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaMemcpyAsync(&h_nirr, misc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  TRVB_REQUIRE(h_total >= 0, "owner sort: copy count overflows int indices");
  const size_t ncopy = (size_t)std::max(h_total, 1);
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->own_rec, sizeof(double4) * ncopy));
  if (!d.pk_in_w) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->own_pk, sizeof(int) * ncopy));
  if (need_src) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->own_src, sizeof(int) * ncopy));
  if (h_nirr > 0) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->own_irr, sizeof(int) * (size_t)h_nirr));
  TRVB_CUDA(cudaMemcpyAsync(cursor, offsets, sizeof(int) * (size_t)nkeys,
                            cudaMemcpyDeviceToDevice, ctx->stream));
  k_own_sort<ORDER><<<blocks, threads, 0, ctx->stream>>>(cv, d, offsets, cursor, misc + 1,
                                                         cat->own_rec, cat->own_pk, cat->own_src,
                                                         cat->own_irr);
  TRVB_LAUNCH_CHECK();
  TRVB_CUDA(trvb_dev_free_raw(ctx, cursor));
  TRVB_CUDA(trvb_dev_free_raw(ctx, chunk_sums));
  TRVB_CUDA(trvb_dev_free_raw(ctx, misc));
  cat->own_offsets = offsets;
  cat->own_total = h_total; cat->own_nirr = h_nirr;
  for (int a = 0; a < 3; a++) { cat->own_n[a] = g.n[a]; cat->own_L[a] = g.L[a]; }
  cat->own_shifted = shifted; cat->own_order = ORDER;
  return 0;
}

template <int ORDER, bool COMPLEX, bool HEAVY>
int launch_own_general(trvb_ctx* ctx, const OwnView& v, const GridDesc& g, const OwnDesc& d, int kind,
                       const YlmCoef& yc, double s, int accumulate, double* mesh, unsigned nblocks,
                       trvb_cat* cat, int L, int M, int shifted) {
  // all of the SM's shared memory to the rings: seven one-warp CTAs per SM
  TRVB_CUDA(cudaFuncSetAttribute(k_assign_own<ORDER, COMPLEX, HEAVY>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  TRVB_CUDA(cudaFuncSetAttribute(k_assign_own<ORDER, COMPLEX, HEAVY>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)OwnGeom<ORDER>::SMEM));
  k_assign_own<ORDER, COMPLEX, HEAVY><<<nblocks, 32, OwnGeom<ORDER>::SMEM, ctx->stream>>>(
    v, g, d, kind, yc, s, accumulate, mesh);
  TRVB_LAUNCH_CHECK();
  if (cat->own_nirr > 0) {
    const int blocks = div_up(cat->own_nirr, 128);
    k_assign_irregular<ORDER, COMPLEX><<<blocks, 128, 0, ctx->stream>>>(
      view_of(cat), cat->own_irr, cat->own_nirr, g, shifted, kind, L, M, s, mesh);
    TRVB_LAUNCH_CHECK();
  }
  return 0;
}

template <int ORDER, bool COMPLEX>
int launch_own_kernels(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M, const YlmCoef& yc,
                       double s, int accumulate, int shifted, double* mesh) {
  const GridDesc& g = ctx->g;
  const OwnDesc d = own_desc(g, shifted);
  const long long ntasks = (long long)d.nch[0] * d.nch[1] * d.nch[2];
  OwnView v;
  v.rec = cat->own_rec; v.pk = cat->own_pk; v.src = cat->own_src; v.offsets = cat->own_offsets;
  v.lx = cat->los; v.ly = cat->los ? cat->los + cat->n : nullptr;
  v.lz = cat->los ? cat->los + 2 * cat->n : nullptr;
  v.cw = cat->cw;
  const long long nblocks = ntasks * (COMPLEX ? 2 : 1);
  TRVB_REQUIRE(nblocks < 2147483647LL, "too many assignment tasks");
  // weights that read the lines of sight or the custom column get their own instantiation
  const bool heavy = kind == TRVB_W_CUSTOM
    || ((kind == TRVB_W_YLM_W || kind == TRVB_W_CYLM_W2) && !(L == 0 && M == 0));
  const int kind_eff = (!heavy && kind != TRVB_W_UNIT) ? (int)TRVB_W_W : kind;
  if (kind == TRVB_W_CYLM_W2 && !heavy) {
    // y_00 = 1: conj(y_00) w^2 is the weight squared -- evaluated by the general path
    return launch_own_general<ORDER, COMPLEX, true>(ctx, v, g, d, kind, yc, s, accumulate, mesh,
                                                    (unsigned)nblocks, cat, L, M, shifted);
  }
  return heavy
    ? launch_own_general<ORDER, COMPLEX, true>(ctx, v, g, d, kind, yc, s, accumulate, mesh,
                                               (unsigned)nblocks, cat, L, M, shifted)
    : launch_own_general<ORDER, COMPLEX, false>(ctx, v, g, d, kind_eff, yc, s, accumulate, mesh,
                                                (unsigned)nblocks, cat, L, M, shifted);
}

template <int ORDER>
void launch_throughput_scatter(trvb_ctx* ctx, const SortedView& cv, int kind, const YlmCoef& yc,
                               double s, int shifted, bool cplx_mesh, double* mesh) {
  const GridDesc& g = ctx->g;
  const int threads = 256;
  if (ORDER >= 3) {
    const long long nchunk = (cv.n + 31) / 32;
    const int blocks = (int)std::min<long long>(div_up(nchunk, 8), (long long)ctx->num_sms * 64);
    if (cplx_mesh) {
      k_assign_coop<(ORDER >= 3 ? ORDER : 3), true><<<blocks, threads, 0, ctx->stream>>>(
        cv, g, shifted, kind, yc, s, mesh);
    } else {
      k_assign_coop<(ORDER >= 3 ? ORDER : 3), false><<<blocks, threads, 0, ctx->stream>>>(
        cv, g, shifted, kind, yc, s, mesh);
    }
  } else {
    const int blocks = div_up(cv.n, threads);
    if (cplx_mesh) {
      k_assign_scatter<ORDER, true><<<blocks, threads, 0, ctx->stream>>>(
        cv, g, shifted, kind, yc, s, mesh);
    } else {
      k_assign_scatter<ORDER, false><<<blocks, threads, 0, ctx->stream>>>(
        cv, g, shifted, kind, yc, s, mesh);
    }
  }
}

template <int ORDER>
int launch_assign(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M, double scale,
                  int density_units, int accumulate, int shifted, int mode,
                  trvb_mesh mesh) {
  const GridDesc& g = ctx->g;
  const bool cplx_mesh = mesh.layout == TRVB_COMPLEX;
  // y_LM recursion coefficients: evaluated once on the host (IEEE sqrt and division,
  // the same bits as on the device) instead of once per thread.
  const YlmCoef yc = ylm_coef(L, M);
  if (mode == 0) {
    {
      const char* env_tile_pre = getenv("TRV_ASSIGN_TILE");
      // the tile kernel needs global (tile, column) offsets: re-sort a chunked order
      if (env_tile_pre != nullptr && env_tile_pre[0] == '1' && cat->chunked) trvb_cat_invalidate_sort(cat);
    }
    const double s = density_units ? scale * (1. / g.vol_cell) : scale;
    // TRV_ASSIGN_OWN=1 selects the tile-owned, store-once assignment (trvb_assign_own.cuh):
    // no zero-fill, no RED, 1.08x the algorithmic DRAM traffic instead of 2.65x -- but on
    // B200 it only ties the warp-cooperative scatter below (PCS, C2: 1.48 vs 1.60 ms, and
    // 2.35 vs 2.03 ms once its 1.7x larger counting sort is included): it saturates the SM's
    // LSU wavefront pipe (87 %) where the scatter saturates the L2 RED units (76 %), see
    // profiles/r02_assign_own_vs_coop.txt.  Opt-in, parity-tested.
    {
      const char* env_own = getenv("TRV_ASSIGN_OWN");
      if (env_own && env_own[0] == '1' && !cat->chunked) {
        int st = ensure_own_sorted<ORDER>(ctx, cat, shifted);
        if (st) return st;
        return cplx_mesh
          ? launch_own_kernels<ORDER, true>(ctx, cat, kind, L, M, yc, s, accumulate, shifted,
                                            (double*)mesh.data)
          : launch_own_kernels<ORDER, false>(ctx, cat, kind, L, M, yc, s, accumulate, shifted,
                                             (double*)mesh.data);
      }
    }
    int st = ensure_sorted(ctx, cat, shifted, 0);
    if (st) return st;
    SortedView cv = sorted_view_of(cat);
    if (!accumulate) {
      TRVB_CUDA(cudaMemsetAsync(mesh.data, 0, trvb_mesh_bytes(ctx, mesh.layout), ctx->stream));
    }
    const int threads = 256;
    // TRV_ASSIGN_TILE=1 selects the tile-owned shared-memory kernel.  Measured on
    // B200 (profiles/r01d_assign_tile_vs_coop.txt) it only ties the cooperative
    // scatter on C2 (1.67 vs 1.61 ms) and loses on the dense C3 randoms: it trades
    // the L2 RED-sector bound for an SM latency bound at 8 warps per SM, so the
    // cooperative scatter stays the default.
    const char* env_tile = getenv("TRV_ASSIGN_TILE");
    const bool use_tile = env_tile != nullptr && env_tile[0] == '1';
    // The tile kernel needs the mesh to be at least one footprint wide (no
    // self-overlap of a tile's periodic footprint).
    const bool fits = g.n[0] >= FP_X && g.n[1] >= FP_Y && g.n[2] >= FP_Z;
    // Column-owned accumulation (TRV_ASSIGN_COL=1).  Measured on the 5e7 randoms of config 3
    // (~220 particles per occupied column, complex TSC meshes) it is SLOWER than the
    // cooperative scatter (fields_LM 58.7 vs 47.5 ms): the per-particle LDS -> DADD -> STS
    // chain is latency bound at 12-20 warps per SM, like the tile-owned variant.  Opt-in.
    const char* env_col = getenv("TRV_ASSIGN_COL");
    const long long nkeys_col = (long long)((g.n[0] + TILE_X - 1) / TILE_X)
      * ((g.n[1] + TILE_Y - 1) / TILE_Y) * ((g.n[2] + TILE_Z - 1) / TILE_Z) * KEYS_PER_TILE;
    const bool use_col = ORDER >= 3 && !shifted && !use_tile && !cat->chunked && cat->cell_start
      && g.n[0] >= COLF_X && g.n[1] >= COLF_Y && g.n[2] >= COLF_Z
      && env_col != nullptr && env_col[0] == '1';
    if (use_col) {
      constexpr int O = ORDER >= 3 ? ORDER : 3;
      const int nt1 = (g.n[1] + TILE_Y - 1) / TILE_Y, nt2 = (g.n[2] + TILE_Z - 1) / TILE_Z;
      const int blocks = (int)std::min<long long>(div_up(nkeys_col, COL_WARPS),
                                                  (long long)ctx->num_sms * 32);
      if (cplx_mesh) {
        TRVB_CUDA(cudaFuncSetAttribute(k_assign_col<O, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)col_smem_bytes<O, true>()));
        k_assign_col<O, true><<<blocks, COL_WARPS * 32, col_smem_bytes<O, true>(), ctx->stream>>>(
          cv, cat->cell_start, nkeys_col, nt1, nt2, g, kind, yc, s, (double*)mesh.data);
      } else {
        TRVB_CUDA(cudaFuncSetAttribute(k_assign_col<O, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)col_smem_bytes<O, false>()));
        k_assign_col<O, false><<<blocks, COL_WARPS * 32, col_smem_bytes<O, false>(), ctx->stream>>>(
          cv, cat->cell_start, nkeys_col, nt1, nt2, g, kind, yc, s, (double*)mesh.data);
      }
    } else if (ORDER >= 3 && !shifted && use_tile && fits) {
      constexpr int O = ORDER >= 3 ? ORDER : 3;
      const int nt[3] = {(g.n[0] + TILE_X - 1) / TILE_X, (g.n[1] + TILE_Y - 1) / TILE_Y,
                         (g.n[2] + TILE_Z - 1) / TILE_Z};
      const int ntiles = nt[0] * nt[1] * nt[2];
      if (cplx_mesh) {
        TRVB_CUDA(cudaFuncSetAttribute(k_assign_tile<O, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tile_smem_bytes<true>()));
        k_assign_tile<O, true><<<ntiles, TILE_WARPS * 32, tile_smem_bytes<true>(), ctx->stream>>>(
          cv, cat->cell_start, nt[1], nt[2], g, kind, yc, s, (double*)mesh.data);
      } else {
        TRVB_CUDA(cudaFuncSetAttribute(k_assign_tile<O, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tile_smem_bytes<false>()));
        k_assign_tile<O, false><<<ntiles, TILE_WARPS * 32, tile_smem_bytes<false>(), ctx->stream>>>(
          cv, cat->cell_start, nt[1], nt[2], g, kind, yc, s, (double*)mesh.data);
      }
    } else {
      launch_throughput_scatter<ORDER>(ctx, cv, kind, yc, s, shifted, cplx_mesh, (double*)mesh.data);
    }
    TRVB_LAUNCH_CHECK();
  } else {
    TRVB_REQUIRE(g.n[0] >= 4 && g.n[1] >= 4 && g.n[2] >= 4,
                 "deterministic assignment needs at least 4 cells per axis");
    int st = ensure_sorted(ctx, cat, shifted, 1);
    if (st) return st;
    SortedView cv = sorted_view_of(cat);
    const double pre = density_units ? 1. / g.vol_cell : 1.;   // S/field.cpp:996
    const int threads = 128;
    const int blocks = div_up(g.nmesh, threads);
    // Tile form (sorted candidate lists) for every scheme, warp-cooperative merge for the
    // wide stencils on smaller meshes; the union of home cells of a tile (plus the overhang
    // of a partial tile) must not wrap onto itself.
    const char* env_tile = getenv("TRV_DET_NO_TILE");
    const bool tile_form = !(env_tile && env_tile[0] == '1')
      && g.n[0] >= 2 * DT_X + 3 && g.n[1] >= 2 * DT_Y + 3 && g.n[2] >= 2 * DT_Z + 3;
    const bool warp_form = ORDER >= 3
      && g.n[0] >= 2 * GT_X + 3 && g.n[1] >= 2 * GT_Y + 3 && g.n[2] >= 2 * GT_Z + 3;
    const int nt[3] = {(g.n[0] + GT_X - 1) / GT_X, (g.n[1] + GT_Y - 1) / GT_Y,
                       (g.n[2] + GT_Z - 1) / GT_Z};
    const long long ntiles = (long long)nt[0] * nt[1] * nt[2];
    if (tile_form) {
      // Sorted candidate lists per 128-cell tile; the few tiles with too many candidates
      // go to the merge kernel afterwards.
      const int bt[3] = {(g.n[0] + DT_X - 1) / DT_X, (g.n[1] + DT_Y - 1) / DT_Y,
                         (g.n[2] + DT_Z - 1) / DT_Z};
      const long long nbig = (long long)bt[0] * bt[1] * bt[2];
      TRVB_REQUIRE(nbig < 2147483647LL, "mesh too large for int tile indices");
      int* big = nullptr;
      TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&big, sizeof(int) * (size_t)(nbig + 1)));
      int* big_count = big + nbig;
      TRVB_CUDA(cudaMemsetAsync(big_count, 0, sizeof(int), ctx->stream));
      const int tblocks = (int)div_up(nbig, 4);
      const int fblocks = ctx->num_sms * 8;
      if (cplx_mesh) {
        k_assign_gather_tile<ORDER, true><<<tblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          bt[1], bt[2], nbig, big, big_count, (double*)mesh.data);
        k_assign_gather_warp<ORDER, true><<<fblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          nt[1], nt[2], ntiles, big, big_count, bt[1], bt[2], (double*)mesh.data);
      } else {
        k_assign_gather_tile<ORDER, false><<<tblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          bt[1], bt[2], nbig, big, big_count, (double*)mesh.data);
        k_assign_gather_warp<ORDER, false><<<fblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          nt[1], nt[2], ntiles, big, big_count, bt[1], bt[2], (double*)mesh.data);
      }
      TRVB_LAUNCH_CHECK();
      TRVB_CUDA(trvb_dev_free_raw(ctx, big));
    } else if (warp_form) {
      const int wblocks = (int)std::min<long long>(div_up(ntiles, 4), 1 << 30);
      if (cplx_mesh) {
        k_assign_gather_warp<ORDER, true><<<wblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          nt[1], nt[2], ntiles, nullptr, nullptr, 0, 0, (double*)mesh.data);
      } else {
        k_assign_gather_warp<ORDER, false><<<wblocks, 128, 0, ctx->stream>>>(
          cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre, accumulate,
          nt[1], nt[2], ntiles, nullptr, nullptr, 0, 0, (double*)mesh.data);
      }
    } else if (cplx_mesh) {
      k_assign_gather<ORDER, true><<<blocks, threads, 0, ctx->stream>>>(
        cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre,
        accumulate, (double*)mesh.data);
    } else {
      k_assign_gather<ORDER, false><<<blocks, threads, 0, ctx->stream>>>(
        cv, cat->order, cat->cell_start, g, shifted, kind, yc, scale, pre,
        accumulate, (double*)mesh.data);
    }
    TRVB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

// =====================================================================
// C ABI
// =====================================================================

extern "C" int trvb_cat_create(trvb_ctx* ctx, trvb_cat** out, long long n,
                               const double* x, const double* y, const double* z,
                               const double* w, const double* los, int src_on_device) {
  TRVB_REQUIRE(ctx && out && x && y && z && n > 0, "trvb_cat_create: bad argument");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  trvb_cat* cat = new trvb_cat();
  cat->owner = ctx; cat->n = n;
  const size_t nb = sizeof(double) * (size_t)n;
  const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  int peer = -1;   // >= 0: the device arrays live on that other GPU and are copied over
  if (src_on_device) {
    // Device sources normally live on the context's GPU.  In a single-process multi-GPU run
    // the caller's arrays belong to one GPU only: the other contexts copy them over
    // (cudaMemcpyPeer).  Anything else (host memory passed as device memory) is an error.
    const double* srcs[5] = {x, y, z, w, los};
    for (const double* p : srcs) {
      if (!p) continue;
      cudaPointerAttributes attr;
      TRVB_CUDA(cudaPointerGetAttributes(&attr, p));
      if (attr.type == cudaMemoryTypeManaged) continue;
      if (attr.type != cudaMemoryTypeDevice) {
        delete cat;
        trvb_set_error("trvb_cat_create: array %p passed as device memory is not (type %d)",
                       (const void*)p, (int)attr.type);
        return 2;
      }
      if (attr.device != ctx->device) {
        if (peer >= 0 && peer != attr.device) {
          delete cat;
          trvb_set_error("trvb_cat_create: device arrays spread over GPUs %d and %d", peer, attr.device);
          return 2;
        }
        peer = attr.device;
      }
    }
    if (peer >= 0) {
      src_on_device = 1;   // copy, never borrow
      // direct NVLink copies instead of staging through the host
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx->device, peer) == cudaSuccess && can) {
        const cudaError_t pe = cudaDeviceEnablePeerAccess(peer, 0);
        if (pe != cudaSuccess) cudaGetLastError();   // already enabled
      } else {
        cudaGetLastError();
      }
    }
  }
  auto copy_in = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (peer >= 0) return cudaMemcpyPeerAsync(dst, ctx->device, src, peer, bytes, ctx->stream);
    return cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream);
  };
  if (src_on_device == 2) {
    // Borrowed: the caller keeps the coordinate arrays alive and unchanged for the
    // life of the catalogue (one estimator call); no copy, never freed here.
    cat->x = const_cast<double*>(x); cat->y = const_cast<double*>(y); cat->z = const_cast<double*>(z);
    cat->borrowed_xyz = true;
  } else {
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->x, nb));
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->y, nb));
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->z, nb));
    TRVB_CUDA(copy_in(cat->x, x, nb));
    TRVB_CUDA(copy_in(cat->y, y, nb));
    TRVB_CUDA(copy_in(cat->z, z, nb));
  }
  if (w) {
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->w, nb));
    TRVB_CUDA(copy_in(cat->w, w, nb));
  }
  if (los) {
    double* tmp = nullptr;
    TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->los, 3 * nb));
    const bool los_in_place = src_on_device && peer < 0;
    if (los_in_place) {
      tmp = const_cast<double*>(los);
    } else {
      TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&tmp, 3 * nb));
      TRVB_CUDA(copy_in(tmp, los, 3 * nb));
    }
    k_los_to_soa<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(tmp, n, cat->los);
    TRVB_LAUNCH_CHECK();
    if (!los_in_place) {
      TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
      TRVB_CUDA(trvb_dev_free_raw(ctx, tmp));
    }
  }
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = cat;
  return 0;
}

namespace {

// Pinned staging ring for uploads from pageable host memory: the driver's own
// pageable path moved the 50M-particle random catalogue of BASELINE config 3 at
// ~11 GB/s (367 ms, two thirds of the estimator call); packing only the columns the
// device needs into pinned chunks with a few host threads while the previous
// chunk is on the wire is bound by the host's memory bandwidth instead.
constexpr long long STAGE_CHUNK = 1 << 20;          // particles per chunk
constexpr int STAGE_DOUBLES = 7;                    // {x, y, z, w} + {lx, ly, lz}
struct PinnedStage {
  double* host[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool ready = false;
};
struct UploadLane { cudaStream_t stream = nullptr; std::vector<cudaEvent_t> done; };
// Upload state of ONE device: events and streams belong to the device that was current
// when they were made, and in single-process multi-GPU mode (one host thread per GPU)
// every device uploads its own copy of the catalogue at the same time.
struct DeviceUpload {
  std::mutex mutex;      // one upload at a time per device
  PinnedStage stage;
  UploadLane lane;
};
std::mutex g_upload_map_mutex;
std::map<int, DeviceUpload> g_upload;   // nodes of a std::map never move

DeviceUpload& upload_state(int device) {
  std::lock_guard<std::mutex> lock(g_upload_map_mutex);
  return g_upload[device];
}

// current device = the one `st` belongs to
cudaError_t ensure_stage(PinnedStage& st) {
  if (st.ready) return cudaSuccess;
  for (int b = 0; b < 2; b++) {
    if (!st.host[b]) {
      cudaError_t e = cudaHostAlloc((void**)&st.host[b],
                                    sizeof(double) * STAGE_DOUBLES * STAGE_CHUNK,
                                    cudaHostAllocPortable);
      if (e != cudaSuccess) return e;
    }
    if (!st.done[b]) {
      cudaError_t e = cudaEventCreateWithFlags(&st.done[b], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
  }
  st.ready = true;
  return cudaSuccess;
}

// Frees a half-built catalogue and the staging blocks when an upload fails part-way.
struct UploadGuard {
  trvb_ctx* ctx; trvb_cat* cat; void* blocks[3] = {nullptr, nullptr, nullptr};
  bool armed = true;
  ~UploadGuard() {
    if (!armed) return;
    cudaStreamSynchronize(ctx->stream);
    for (void* b : blocks) if (b) trvb_dev_free_raw(ctx, b);
    trvb_cat_destroy(cat);
  }
};

template <class F>
void host_parallel(long long m, F body) {
  const unsigned hw = std::thread::hardware_concurrency();
  const char* env_nt = getenv("TRV_UPLOAD_THREADS");
  const unsigned cap = env_nt ? (unsigned)std::max(1, atoi(env_nt)) : 8u;
  const int nt = (int)std::max(1u, std::min(cap, hw ? hw : 1u));
  if (nt == 1 || m < 65536) { body(0, m); return; }
  std::vector<std::thread> pool;
  const long long per = (m + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const long long lo = t * per, hi = std::min(m, lo + per);
    if (lo < hi) pool.emplace_back([=]() { body(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

// One staged chunk -> the SoA columns: [4 m] records {x, y, z, w}, then [3 m] LOS.
__global__ void k_unpack_chunk(const double* __restrict__ chunk, long long m, long long off,
                               long long n, double* x, double* y, double* z, double* w,
                               double* los /* may be null */) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
       i += (long long)gridDim.x * blockDim.x) {
    const double4 r = reinterpret_cast<const double4*>(chunk)[i];
    x[off + i] = r.x; y[off + i] = r.y; z[off + i] = r.z; w[off + i] = r.w;
    if (los) {
      const double* l = chunk + 4 * m + 3 * i;
      los[off + i] = l[0]; los[n + off + i] = l[1]; los[2 * n + off + i] = l[2];
    }
  }
}

}  // namespace

extern "C" int trvb_cat_create_aos(trvb_ctx* ctx, trvb_cat** out, long long n,
                                   const double* pdata, const double* los) {
  TRVB_REQUIRE(ctx && out && pdata && n > 0, "trvb_cat_create_aos: bad argument");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  DeviceUpload& up = upload_state(ctx->device);
  std::lock_guard<std::mutex> lock(up.mutex);
  PinnedStage& stage = up.stage;
  TRVB_CUDA(ensure_stage(stage));
  const size_t nb = sizeof(double) * (size_t)n;
  trvb_cat* cat = new trvb_cat();
  cat->owner = ctx; cat->n = n;
  UploadGuard guard{ctx, cat};
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->x, nb));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->y, nb));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->z, nb));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->w, nb));
  if (los) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->los, 3 * nb));
  const size_t chunk_bytes = sizeof(double) * STAGE_DOUBLES * STAGE_CHUNK;
  double* d_chunk[2] = {nullptr, nullptr};
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_chunk[0], chunk_bytes));
  guard.blocks[0] = d_chunk[0];
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&d_chunk[1], chunk_bytes));
  guard.blocks[1] = d_chunk[1];
  const int width = los ? 7 : 4;
  long long c = 0;
  for (long long off = 0; off < n; off += STAGE_CHUNK, c++) {
    const long long m = std::min<long long>(STAGE_CHUNK, n - off);
    const int b = (int)(c & 1);
    // the DMA that last read this pinned buffer must have finished
    if (c >= 2) TRVB_CUDA(cudaEventSynchronize(stage.done[b]));
    double* h = stage.host[b];
    host_parallel(m, [=](long long lo, long long hi) {
      for (long long i = lo; i < hi; i++) {
        // ParticleData {pos[3], nz, ws, wc, w} (I/particles.hpp:63-69)
        const double* p = pdata + 7 * (off + i);
        double* r = h + 4 * i;
        r[0] = p[0]; r[1] = p[1]; r[2] = p[2]; r[3] = p[6];
      }
      if (los) std::memcpy(h + 4 * m + 3 * lo, los + 3 * (off + lo), sizeof(double) * 3 * (hi - lo));
    });
    TRVB_CUDA(cudaMemcpyAsync(d_chunk[b], h, sizeof(double) * width * m, cudaMemcpyHostToDevice,
                              ctx->stream));
    TRVB_CUDA(cudaEventRecord(stage.done[b], ctx->stream));
    const int blocks = (int)std::min<long long>(div_up(m, 256), (long long)ctx->num_sms * 8);
    k_unpack_chunk<<<blocks, 256, 0, ctx->stream>>>(d_chunk[b], m, off, n, cat->x, cat->y, cat->z,
                                                   cat->w, los ? cat->los : nullptr);
    TRVB_LAUNCH_CHECK();
  }
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  guard.armed = false;
  TRVB_CUDA(trvb_dev_free_raw(ctx, d_chunk[0]));
  TRVB_CUDA(trvb_dev_free_raw(ctx, d_chunk[1]));
  *out = cat;
  return 0;
}

// ---------------------------------------------------------------------
// Streamed upload + assignment of a host catalogue (unit weights).
// ---------------------------------------------------------------------
namespace {

template <int ORDER>
int streamed_assign(trvb_ctx* ctx, trvb_cat* cat, UploadLane& lane, const double* hx, const double* hy,
                    const double* hz, double scale, trvb_mesh mesh) {
  const GridDesc& g = ctx->g;
  const long long n = cat->n;
  const bool cplx_mesh = mesh.layout == TRVB_COMPLEX;
  if (!lane.stream) TRVB_CUDA(cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
  // Every chunk costs one sweep over the mesh (its REDs touch sectors all over it: the
  // four-chunk form spends 2.6 ms in the scatter kernels instead of 1.4), so chunks are
  // few and shrink geometrically: 1/2, 1/4, 1/4 of the catalogue by default -- the work
  // left after the last byte has arrived is a quarter of the sort + assignment.
  // TRV_STREAM_LEVELS = number of halvings (C2 end to end: 1: 10.66, 2: 10.40, 3: 10.58,
  // 4: 10.53 ms).
  std::vector<long long> bounds(1, 0);
  {
    const char* env_levels = getenv("TRV_STREAM_LEVELS");
    const int TRVB_STREAM_LEVELS = env_levels ? atoi(env_levels) : 2;   // measured best on C2
    const long long min_chunk = 1LL << 20;
    long long left = n;
    for (int level = 0; level < TRVB_STREAM_LEVELS && left > 2 * min_chunk; level++) {
      const long long take = std::max(min_chunk, left / 2);
      bounds.push_back(bounds.back() + take);
      left -= take;
    }
    bounds.push_back(n);
  }
  const int nchunks = (int)bounds.size() - 1;
  while ((int)lane.done.size() < nchunks) {
    cudaEvent_t e; TRVB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    lane.done.push_back(e);
  }
  SortDesc d;
  for (int a = 0; a < 3; a++) {
    d.n[a] = g.n[a]; d.L[a] = g.L[a];
    const int tile = (a == 0) ? TILE_X : (a == 1) ? TILE_Y : TILE_Z;
    d.nk[a] = (g.n[a] + tile - 1) / tile;
  }
  d.shifted = 0; d.by_cell = 0;
  const long long nkeys = (long long)d.nk[0] * d.nk[1] * d.nk[2] * KEYS_PER_TILE;
  TRVB_REQUIRE(nkeys < 2147483647LL, "mesh too large for int sort keys");
  { int st = alloc_sorted(ctx, cat); if (st) return st; }
  int* offsets = nullptr; int* cursor = nullptr; int* chunk_sums = nullptr;
  struct Temporaries {   // returned to the arena on every exit, error paths included
    trvb_ctx* ctx; int** p[3];
    ~Temporaries() { for (int** q : p) if (*q) { trvb_dev_free_raw(ctx, *q); *q = nullptr; } }
  } temporaries{ctx, {&offsets, &cursor, &chunk_sums}};
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&offsets, sizeof(int) * (size_t)(nkeys + 1)));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cursor, sizeof(int) * (size_t)nkeys));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&chunk_sums,
                               sizeof(int) * (size_t)div_up(nkeys + 1, SCAN_CHUNK)));
  // The upload lane must not overtake work still queued on the compute stream that
  // may use the freshly allocated blocks (stream-ordered arena).
  cudaEvent_t start = lane.done[0];
  TRVB_CUDA(cudaEventRecord(start, ctx->stream));
  TRVB_CUDA(cudaStreamWaitEvent(lane.stream, start, 0));
  TRVB_CUDA(cudaMemsetAsync(mesh.data, 0, trvb_mesh_bytes(ctx, mesh.layout), ctx->stream));
  const YlmCoef yc = ylm_coef(0, 0);
  for (int c = 0; c < nchunks; c++) {
    const long long off = bounds[c], m = bounds[c + 1] - off;
    const size_t mb = sizeof(double) * (size_t)m;
    TRVB_CUDA(cudaMemcpyAsync(cat->x + off, hx + off, mb, cudaMemcpyHostToDevice, lane.stream));
    TRVB_CUDA(cudaMemcpyAsync(cat->y + off, hy + off, mb, cudaMemcpyHostToDevice, lane.stream));
    TRVB_CUDA(cudaMemcpyAsync(cat->z + off, hz + off, mb, cudaMemcpyHostToDevice, lane.stream));
    TRVB_CUDA(cudaEventRecord(lane.done[c], lane.stream));
    TRVB_CUDA(cudaStreamWaitEvent(ctx->stream, lane.done[c], 0));
    // sort this chunk by (tile, column) and spread it while the next one is on the wire
    CatView cv; cv.x = cat->x + off; cv.y = cat->y + off; cv.z = cat->z + off;
    cv.w = nullptr; cv.lx = cv.ly = cv.lz = nullptr; cv.cw = nullptr; cv.n = m;
    int st = enqueue_counting_sort(ctx, cv, d, nkeys, offsets, cursor, chunk_sums, nullptr,
                                   cat->s4 + off);
    if (st) return st;
    SortedView sv; sv.p4 = cat->s4 + off; sv.lx = sv.ly = sv.lz = nullptr; sv.cw = nullptr; sv.n = m;
    launch_throughput_scatter<ORDER>(ctx, sv, TRVB_W_UNIT, yc, scale, 0, cplx_mesh,
                                     (double*)mesh.data);
    TRVB_LAUNCH_CHECK();
  }
  for (int a = 0; a < 3; a++) { cat->sort_n[a] = g.n[a]; cat->sort_L[a] = g.L[a]; }
  cat->sort_shifted = 0; cat->sort_kind = 0;
  cat->order_valid = false; cat->chunked = true; cat->scw_valid = false;
  // the host arrays are the caller's: every copy must have left them
  TRVB_CUDA(cudaStreamSynchronize(lane.stream));
  return 0;
}

}  // namespace

extern "C" int trvb_cat_create_assign(trvb_ctx* ctx, trvb_cat** out, long long n,
                                      const double* x, const double* y, const double* z,
                                      double scale, trvb_mesh mesh) {
  TRVB_REQUIRE(ctx && out && x && y && z && n > 0 && mesh.data,
               "trvb_cat_create_assign: bad argument");
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_cat_create_assign: configuration-space mesh required");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_cat_create_assign: root context only");
  TRVB_REQUIRE(n < 2147483647LL, "catalogue too large for int indices");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  DeviceUpload& up = upload_state(ctx->device);
  std::lock_guard<std::mutex> lock(up.mutex);
  UploadLane& lane = up.lane;
  trvb_cat* cat = new trvb_cat();
  cat->owner = ctx; cat->n = n;
  UploadGuard guard{ctx, cat};
  const size_t nb = sizeof(double) * (size_t)n;
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->x, nb));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->y, nb));
  TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->z, nb));
  int st = 0;
  switch (ctx->g.order) {
    case 1: st = streamed_assign<1>(ctx, cat, lane, x, y, z, scale, mesh); break;
    case 2: st = streamed_assign<2>(ctx, cat, lane, x, y, z, scale, mesh); break;
    case 3: st = streamed_assign<3>(ctx, cat, lane, x, y, z, scale, mesh); break;
    case 4: st = streamed_assign<4>(ctx, cat, lane, x, y, z, scale, mesh); break;
    default: st = 2;
  }
  if (st) return st;   // the guard frees the half-built catalogue
  guard.armed = false;
  *out = cat;
  return 0;
}

extern "C" int trvb_cat_set_custom_weights(trvb_ctx* ctx, trvb_cat* cat,
                                           const double* weights) {
  TRVB_REQUIRE(ctx && cat && weights, "trvb_cat_set_custom_weights: null argument");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  const size_t nb = 2 * sizeof(double) * (size_t)cat->n;
  if (!cat->cw) TRVB_CUDA(trvb_dev_alloc_raw(ctx, (void**)&cat->cw, nb));
  TRVB_CUDA(cudaMemcpyAsync(cat->cw, weights, nb, cudaMemcpyHostToDevice, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  cat->scw_valid = false;   // re-gather into sorted order on next use
  return 0;
}

extern "C" void trvb_cat_destroy(trvb_cat* cat) {
  if (!cat) return;
  if (cat->owner) cudaSetDevice(cat->owner->device);
  trvb_ctx* o = cat->owner;
  if (!cat->borrowed_xyz) {
    trvb_dev_free_raw(o, cat->x); trvb_dev_free_raw(o, cat->y); trvb_dev_free_raw(o, cat->z);
  }
  trvb_dev_free_raw(o, cat->w); trvb_dev_free_raw(o, cat->los); trvb_dev_free_raw(o, cat->cw);
  trvb_dev_free_raw(o, cat->order); trvb_dev_free_raw(o, cat->cell_start);
  trvb_dev_free_raw(o, cat->s4); trvb_dev_free_raw(o, cat->slos); trvb_dev_free_raw(o, cat->scw);
  trvb_dev_free_raw(o, cat->own_rec); trvb_dev_free_raw(o, cat->own_pk);
  trvb_dev_free_raw(o, cat->own_src); trvb_dev_free_raw(o, cat->own_offsets);
  trvb_dev_free_raw(o, cat->own_irr);
  delete cat;
}

extern "C" void trvb_cat_invalidate_sort(trvb_cat* cat) {
  if (cat) { cat->sort_kind = -1; cat->sort_shifted = -1; cat->own_shifted = -1; }
}

extern "C" long long trvb_cat_size(const trvb_cat* cat) { return cat ? cat->n : 0; }

extern "C" int trvb_cat_positions(const trvb_cat* cat, const double** x, const double** y,
                                  const double** z) {
  TRVB_REQUIRE(cat && x && y && z, "trvb_cat_positions: null argument");
  *x = cat->x; *y = cat->y; *z = cat->z;
  return 0;
}

extern "C" int trvb_cat_sum(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M,
                            double out[2]) {
  TRVB_REQUIRE(ctx && cat && out, "trvb_cat_sum: null argument");
  TRVB_REQUIRE(kind >= TRVB_W_UNIT && kind <= TRVB_W_YLM_W3, "trvb_cat_sum: kind %d", kind);
  TRVB_REQUIRE(kind < TRVB_W_YLM_W || (L == 0 && M == 0) || cat->los,
               "trvb_cat_sum: y_LM weights need lines of sight");
  const int threads = 256;
  const int blocks = (int)std::min<long long>(div_up(cat->n, threads), (long long)ctx->num_sms * 8);
  double* scratch;
  int st = trvb_scratch(ctx, sizeof(double) * (2 * (size_t)blocks + 2), &scratch);
  if (st) return st;
  k_cat_sum<<<blocks, threads, 0, ctx->stream>>>(view_of(cat), kind, L, M, scratch);
  TRVB_LAUNCH_CHECK();
  k_sum_partials<<<1, 32, 0, ctx->stream>>>(scratch, blocks, 2, scratch + 2 * blocks);
  TRVB_LAUNCH_CHECK();
  TRVB_CUDA(cudaMemcpyAsync(out, scratch + 2 * blocks, 2 * sizeof(double),
                            cudaMemcpyDeviceToHost, ctx->stream));
  TRVB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int trvb_assign(trvb_ctx* ctx, trvb_cat* cat, int kind, int L, int M,
                           double scale, int density_units, int accumulate, int shifted,
                           int mode, trvb_mesh mesh) {
  TRVB_REQUIRE(ctx && cat && mesh.data, "trvb_assign: null argument");
  TRVB_REQUIRE(ctx->parent == nullptr, "trvb_assign: not available on a sub-grid context");
  TRVB_REQUIRE((kind >= TRVB_W_UNIT && kind <= TRVB_W_CYLM_W2) || kind == TRVB_W_CUSTOM,
               "trvb_assign: weight kind %d", kind);
  TRVB_REQUIRE(mesh.layout == TRVB_REAL || mesh.layout == TRVB_COMPLEX,
               "trvb_assign: mesh layout must be REAL or COMPLEX");
  TRVB_REQUIRE(kind != TRVB_W_CUSTOM || cat->cw,
               "trvb_assign: no custom weights attached (trvb_cat_set_custom_weights)");
  // y_L0 is real-valued, so M == 0 weights may target a REAL mesh.
  const bool real_weights = kind <= TRVB_W_W || (kind != TRVB_W_CUSTOM && M == 0);
  TRVB_REQUIRE(real_weights || mesh.layout == TRVB_COMPLEX,
               "trvb_assign: complex weights need a COMPLEX mesh");
  const bool needs_los = (kind == TRVB_W_YLM_W || kind == TRVB_W_CYLM_W2) && !(L == 0 && M == 0);
  TRVB_REQUIRE(!needs_los || cat->los, "trvb_assign: y_LM weights need lines of sight");
  TRVB_CUDA(cudaSetDevice(ctx->device));
  switch (ctx->g.order) {
    case 1: return launch_assign<1>(ctx, cat, kind, L, M, scale, density_units, accumulate, shifted, mode, mesh);
    case 2: return launch_assign<2>(ctx, cat, kind, L, M, scale, density_units, accumulate, shifted, mode, mesh);
    case 3: return launch_assign<3>(ctx, cat, kind, L, M, scale, density_units, accumulate, shifted, mode, mesh);
    case 4: return launch_assign<4>(ctx, cat, kind, L, M, scale, density_units, accumulate, shifted, mode, mesh);
  }
  trvb_set_error("trvb_assign: unsupported order %d", ctx->g.order);
  return 2;
}
