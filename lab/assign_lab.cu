// Development harness (not part of the product): particle-to-mesh assignment
// variants timed side by side on one GPU.  nvcc -O3 -arch=sm_100a lab/assign_lab.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void pcs(double loc, int n, int* ijk, double* win) {
  int idx = __double2int_rz(loc);
  ijk[0] = (idx == 0) ? n - 1 : idx - 1;
  ijk[1] = idx;
  ijk[2] = (idx == n - 1) ? 0 : idx + 1;
  ijk[3] = (ijk[2] == n - 1) ? 0 : ijk[2] + 1;
  const double c6 = 1. / 6;
  double s = loc - idx, u = 1. - s;
  win[0] = c6 * u * u * u;
  win[1] = c6 * (4. - 6. * s * s + 3. * s * s * s);
  win[2] = c6 * (4. - 6. * u * u + 3. * u * u * u);
  win[3] = c6 * s * s * s;
}

// ---------------- sort by sub-tile of 2^SH cells ----------------
template <int SH>
__global__ void k_count(const double* x, const double* y, const double* z, long long np, int n, double L, int* cnt) {
  int nk = n >> SH;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < np; p += (long long)gridDim.x * blockDim.x) {
    int i = min(max((int)(n * x[p] / L), 0), n - 1) >> SH;
    int j = min(max((int)(n * y[p] / L), 0), n - 1) >> SH;
    int k = min(max((int)(n * z[p] / L), 0), n - 1) >> SH;
    atomicAdd(&cnt[(i * nk + j) * nk + k], 1);
  }
}
template <int SH>
__global__ void k_scatter(const double* x, const double* y, const double* z, long long np, int n, double L, int* cur,
                          double* xs, double* ys, double* zs, int* order) {
  int nk = n >> SH;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < np; p += (long long)gridDim.x * blockDim.x) {
    double px = x[p], py = y[p], pz = z[p];
    int i = min(max((int)(n * px / L), 0), n - 1) >> SH;
    int j = min(max((int)(n * py / L), 0), n - 1) >> SH;
    int k = min(max((int)(n * pz / L), 0), n - 1) >> SH;
    int pos = atomicAdd(&cur[(i * nk + j) * nk + k], 1);
    xs[pos] = px; ys[pos] = py; zs[pos] = pz; order[pos] = (int)p;
  }
}
// simple exclusive scan (single block, serial chunks) -- lab only
__global__ void k_scan(int* a, long long n) {
  __shared__ int sm[1024]; __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024 * 8) {
    int v[8]; int loc = 0;
    for (int q = 0; q < 8; q++) { long long id = base + threadIdx.x * 8 + q; v[q] = id < n ? a[id] : 0; loc += v[q]; }
    sm[threadIdx.x] = loc; __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0; __syncthreads(); sm[threadIdx.x] += t; __syncthreads(); }
    int run = carry + sm[threadIdx.x] - loc;
    for (int q = 0; q < 8; q++) { long long id = base + threadIdx.x * 8 + q; if (id < n) a[id] = run; run += v[q]; }
    __syncthreads();
    if (threadIdx.x == 1023) carry += sm[1023];
    __syncthreads();
  }
}

// ---------------- V0: per-thread global REDs through order[] ----------------
__global__ void __launch_bounds__(256) v0(const double* x, const double* y, const double* z, const int* order, long long np, int n, double L, double* mesh) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= np) return;
  long long p = order ? order[t] : t;
  int ia[4], ib[4], ic[4]; double wa[4], wb[4], wc[4];
  pcs(n * x[p] / L, n, ia, wa); pcs(n * y[p] / L, n, ib, wb); pcs(n * z[p] / L, n, ic, wc);
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      long long row = ((long long)ia[a] * n + ib[b]) * n; double w = wa[a] * wb[b];
#pragma unroll
      for (int c = 0; c < 4; c++) atomicAdd(&mesh[row + ic[c]], w * wc[c]);
    }
}

// ---------------- V2: warp-cooperative global REDs (8 rows x 4 z per instruction) ----------------
__global__ void __launch_bounds__(256) v2(const double* xs, const double* ys, const double* zs, long long np, int n, double L, double* mesh) {
  __shared__ double s_w[8][32][12];
  __shared__ int s_i[8][32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool ok = t < np;
  {
    int ia[4], ib[4], ic[4]; double wa[4], wb[4], wc[4];
    double px = ok ? xs[t] : 0., py = ok ? ys[t] : 0., pz = ok ? zs[t] : 0.;
    pcs(n * px / L, n, ia, wa); pcs(n * py / L, n, ib, wb); pcs(n * pz / L, n, ic, wc);
    for (int q = 0; q < 4; q++) {
      s_w[warp][lane][q] = ok ? wa[q] : 0.; s_w[warp][lane][4 + q] = wb[q]; s_w[warp][lane][8 + q] = wc[q];
      s_i[warp][lane][q] = ia[q]; s_i[warp][lane][4 + q] = ib[q]; s_i[warp][lane][8 + q] = ic[q];
    }
  }
  __syncwarp();
  const int c = lane & 3, r = lane >> 2;   // 8 rows x 4 cells
  unsigned mask = __ballot_sync(0xffffffffu, ok);
  for (int s = 0; s < 32; s++) {
    if (!((mask >> s) & 1u)) continue;
    const double* w = s_w[warp][s]; const int* id = s_i[warp][s];
    const double wz = w[8 + c]; const int kz = id[8 + c];
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      int row = r + 8 * pass; int a = row >> 2, b = row & 3;
      long long gid = ((long long)id[a] * n + id[4 + b]) * n + kz;
      atomicAdd(&mesh[gid], w[a] * w[4 + b] * wz);
    }
  }
}

// ---------------- V3: output tile in shared memory, CAS atomics, plain stores ----------------
// Particles sorted by 2^SH-cell sub-tiles.  Tile = TT^3 output cells.  Homes that reach the tile: [t0-2, t0+TT].
template <int SH, int TT>
__global__ void __launch_bounds__(256) v3(const double* xs, const double* ys, const double* zs, const int* start, int n, double L, double* mesh) {
  constexpr int SUB = 1 << SH;
  constexpr int NS = (TT + 2 + SUB - 1) / SUB + 1;     // sub-tiles per axis covering [t0-2, t0+TT]
  extern __shared__ double tile[];                      // TT^3
  __shared__ int r_begin[NS * NS * 2 + 1], r_off[NS * NS * 2 + 2];
  const int nk = n >> SH, ntile = n / TT;
  const int tz = blockIdx.x % ntile, ty = (blockIdx.x / ntile) % ntile, tx = blockIdx.x / (ntile * ntile);
  const int t0x = tx * TT, t0y = ty * TT, t0z = tz * TT;
  for (int q = threadIdx.x; q < TT * TT * TT; q += blockDim.x) tile[q] = 0.;
  // ranges: for each (si, sj) sub-tile row, k-range of sub-tiles [s0, s0+NS) with wrap -> up to 2 ranges
  const int s0x = ((t0x - 2 + n) % n) >> SH, s0y = ((t0y - 2 + n) % n) >> SH, s0z = ((t0z - 2 + n) % n) >> SH;
  // (t0-2) is a multiple of SUB only if SUB | 2 ... for SH=1 yes. for SH=2: t0-2 = 4m+2 -> floor gives sub-tile containing t0-2.
  const int nr = NS * NS;
  for (int q = threadIdx.x; q < nr; q += blockDim.x) {
    int si = (s0x + q / NS) % nk, sj = (s0y + q % NS) % nk;
    long long base = ((long long)si * nk + sj) * nk;
    int k0 = s0z, k1 = s0z + NS;     // [k0, k1) possibly beyond nk
    int b0, e0, b1 = 0, e1 = 0;
    if (k1 <= nk) { b0 = start[base + k0]; e0 = start[base + k1]; }
    else { b0 = start[base + k0]; e0 = start[base + nk]; b1 = start[base]; e1 = start[base + (k1 - nk)]; }
    r_begin[2 * q] = b0; r_off[2 * q + 1] = e0 - b0;       // temporarily lengths in r_off[.+1]
    r_begin[2 * q + 1] = b1; r_off[2 * q + 2] = e1 - b1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0; r_off[0] = 0;
    for (int q = 0; q < 2 * nr; q++) { run += r_off[q + 1]; r_off[q + 1] = run; }
  }
  __syncthreads();
  const int total = r_off[2 * nr];
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int lo = 0, hi = 2 * nr;          // find q with r_off[q] <= t < r_off[q+1]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (r_off[mid] <= t) lo = mid; else hi = mid; }
    const int p = r_begin[lo] + (t - r_off[lo]);
    int ia[4], ib[4], ic[4]; double wa[4], wb[4], wc[4];
    pcs(n * xs[p] / L, n, ia, wa); pcs(n * ys[p] / L, n, ib, wb); pcs(n * zs[p] / L, n, ic, wc);
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int li = ia[a] - t0x;
      if ((unsigned)li >= (unsigned)TT) continue;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int lj = ib[b] - t0y;
        if ((unsigned)lj >= (unsigned)TT) continue;
        const double w = wa[a] * wb[b];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int lk = ic[c] - t0z;
          if ((unsigned)lk < (unsigned)TT) atomicAdd(&tile[(li * TT + lj) * TT + lk], w * wc[c]);
        }
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < TT * TT * TT; q += blockDim.x) {
    int lk = q % TT, lj = (q / TT) % TT, li = q / (TT * TT);
    mesh[((long long)(t0x + li) * n + (t0y + lj)) * n + t0z + lk] = tile[q];
  }
}

int main(int argc, char** argv) {
  long long np = argc > 1 ? atoll(argv[1]) : 10000000LL;
  int n = argc > 2 ? atoi(argv[2]) : 512;
  double L = 1000.;
  std::vector<double> hx(np), hy(np), hz(np);
  unsigned long long s = 88172645463325252ULL;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (s >> 11) * (1.0 / 9007199254740992.0); };
  for (long long i = 0; i < np; i++) { hx[i] = rnd() * L; hy[i] = rnd() * L; hz[i] = rnd() * L; }
  double *x, *y, *z, *xs, *ys, *zs, *m0, *m1; int *order, *cnt, *cur;
  size_t nb = np * sizeof(double); long long nm = (long long)n * n * n;
  CK(cudaMalloc(&x, nb)); CK(cudaMalloc(&y, nb)); CK(cudaMalloc(&z, nb));
  CK(cudaMalloc(&xs, nb)); CK(cudaMalloc(&ys, nb)); CK(cudaMalloc(&zs, nb));
  CK(cudaMalloc(&m0, nm * 8)); CK(cudaMalloc(&m1, nm * 8)); CK(cudaMalloc(&order, np * 4));
  long long maxkeys = (long long)(n / 2) * (n / 2) * (n / 2) + 1;
  CK(cudaMalloc(&cnt, maxkeys * 4)); CK(cudaMalloc(&cur, maxkeys * 4));
  CK(cudaMemcpy(x, hx.data(), nb, cudaMemcpyHostToDevice)); CK(cudaMemcpy(y, hy.data(), nb, cudaMemcpyHostToDevice)); CK(cudaMemcpy(z, hz.data(), nb, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn, int reps = 5) {
    fn(); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for (int r = 0; r < reps; r++) fn(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%-44s %8.3f ms\n", name, ms / reps); fflush(stdout);
  };
  auto sort = [&](int SH) {
    int nk = n >> SH; long long nkeys = (long long)nk * nk * nk;
    CK(cudaMemsetAsync(cnt, 0, (nkeys + 1) * 4));
    if (SH == 2) k_count<2><<<148 * 16, 256>>>(x, y, z, np, n, L, cnt); else k_count<1><<<148 * 16, 256>>>(x, y, z, np, n, L, cnt);
    k_scan<<<1, 1024>>>(cnt, nkeys + 1);
    CK(cudaMemcpyAsync(cur, cnt, nkeys * 4, cudaMemcpyDeviceToDevice));
    if (SH == 2) k_scatter<2><<<148 * 16, 256>>>(x, y, z, np, n, L, cur, xs, ys, zs, order); else k_scatter<1><<<148 * 16, 256>>>(x, y, z, np, n, L, cur, xs, ys, zs, order);
  };
  auto compare = [&](const char* name) {
    std::vector<double> a(1 << 20), b(1 << 20);
    double maxd = 0, maxv = 0, sum = 0;
    for (long long off = 0; off < nm; off += (1 << 20) * 16) {
      CK(cudaMemcpy(a.data(), m0 + off, sizeof(double) << 20, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), m1 + off, sizeof(double) << 20, cudaMemcpyDeviceToHost));
      for (int i = 0; i < (1 << 20); i++) { maxd = std::max(maxd, fabs(a[i] - b[i])); maxv = std::max(maxv, fabs(a[i])); sum += b[i]; }
    }
    printf("   check %-30s max|diff| %.3e (max %.3e) partial sum %.6e\n", name, maxd, maxv, sum);
  };
  timeit("memset mesh", [&]() { CK(cudaMemsetAsync(m0, 0, nm * 8)); });
  timeit("sort SH=2 (4^3 sub-tiles) incl gather", [&]() { sort(2); });
  timeit("V0 order[] indirection + memset", [&]() { CK(cudaMemsetAsync(m0, 0, nm * 8)); v0<<<(np + 255) / 256, 256>>>(x, y, z, order, np, n, L, m0); });
  timeit("V1 gathered positions + memset", [&]() { CK(cudaMemsetAsync(m1, 0, nm * 8)); v0<<<(np + 255) / 256, 256>>>(xs, ys, zs, nullptr, np, n, L, m1); });
  compare("V1");
  timeit("V2 warp-coop REDs + memset", [&]() { CK(cudaMemsetAsync(m1, 0, nm * 8)); v2<<<(np + 255) / 256, 256>>>(xs, ys, zs, np, n, L, m1); });
  compare("V2");
  {
    constexpr int TT = 16; int nt = n / TT; size_t sm = TT * TT * TT * 8;
    CK(cudaFuncSetAttribute(v3<2, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    timeit("V3 smem tile 16^3, SH=2", [&]() { v3<2, TT><<<nt * nt * nt, 256, sm>>>(xs, ys, zs, cnt, n, L, m1); });
    compare("V3 16 SH2");
  }
  timeit("sort SH=1 (2^3 sub-tiles) incl gather", [&]() { sort(1); });
  {
    constexpr int TT = 16; int nt = n / TT; size_t sm = TT * TT * TT * 8;
    CK(cudaFuncSetAttribute(v3<1, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    timeit("V3 smem tile 16^3, SH=1", [&]() { v3<1, TT><<<nt * nt * nt, 256, sm>>>(xs, ys, zs, cnt, n, L, m1); });
    compare("V3 16 SH1");
  }
  {
    constexpr int TT = 32; int nt = n / TT; size_t sm = (size_t)TT * TT * TT * 8;
    CK(cudaFuncSetAttribute(v3<1, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    timeit("V3 smem tile 32^3 (cube 256KB?)", [&]() { });
  }
  {
    constexpr int TT = 8; int nt = n / TT; size_t sm = TT * TT * TT * 8;
    timeit("V3 smem tile 8^3, SH=1", [&]() { v3<1, TT><<<nt * nt * nt, 256, sm>>>(xs, ys, zs, cnt, n, L, m1); });
    compare("V3 8 SH1");
  }
  return 0;
}
