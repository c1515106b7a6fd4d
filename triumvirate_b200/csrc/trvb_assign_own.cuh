// trvb_assign_own.cuh -- tile-owned, store-once particle-to-mesh assignment
// (throughput mode; included by trvb_assign.cu inside its anonymous namespace).
//
// Replaces the scatter loops of S/field.cpp:618-1112 for the throughput mode.  The
// direct scatter (k_assign_coop) zero-fills the mesh and then read-modify-writes it
// through L2 REDs: 2.65x the algorithmic DRAM traffic and an L2 RED-sector bound
// (profiles/r01d_ncu_full_k_assign_coop.csv).  Here every output cell has ONE owner:
//
//   task   = OWN_PX x OWN_PY output pencils x OWN_ZS output planes, one warp (= one CTA)
//   list   = the particles whose stencil touches the task, counting-sorted by
//            (task, first plane touched); a particle near a task border appears in
//            every task it touches (1.48 copies on average for PCS)
//   ring   = OWN_RING planes of the task's pencils in SHARED memory.  The warp walks its
//            list in plane order, one particle per step, lanes = cells of the stencil,
//            plain LDS / DADD / STS (no atomics: the warp is the only writer, and the
//            cells of one stencil are distinct).  A group of four planes is complete as
//            soon as the walk has passed it: it is STORED to the mesh once (or added to
//            it when accumulating) and its ring slots are zeroed for reuse.
//
// No zero-fill of the mesh, no RED, each mesh cell written exactly once; DRAM traffic
// = the lists (36 B per copy) + one mesh write.  The 1-D window weights are evaluated
// per particle with the reference's own operation order (window_weights below), and a
// cell receives `((scale w Wx) Wy) Wz` exactly as S/field.cpp:1044 -- only the order of
// the additions differs from the single-threaded reference.
//
// Particles whose grid index falls outside [0, n) on some axis (positions on or beyond
// the box edge: the reference does not wrap them, it only guards 0 <= gid < nmesh,
// S/field.cpp:1042) are set aside by the sort and added afterwards by k_assign_irregular
// with the reference's index arithmetic.

constexpr int OWN_PX = 16, OWN_PY = 8;         // output pencils of a task along x, y
constexpr int OWN_ZS = 128;                    // output planes of a task
constexpr int OWN_RING = 12;                   // ring depth in planes
constexpr int OWN_G = 4;                       // planes per flush group
// Ring pitches in doubles.  z pitch = ring depth = 12: the four stencil rows b = 0..3 of a
// half-warp (fixed x-plane a0) start 12 doubles apart = {0, 12, 8, 4} (mod 16) eight-byte
// banks, each covering four consecutive ring slots: 16 lanes on 16 distinct banks.
constexpr int OWN_PITCH_Y = OWN_RING;
constexpr int OWN_PITCH_X = OWN_PY * OWN_PITCH_Y + 2;
constexpr int OWN_RING_WORDS = OWN_PX * OWN_PITCH_X;
constexpr int OWN_BATCH = 32;                  // particles staged per round (one per lane)

struct OwnDesc {
  int n[3]; double L[3];
  int shifted;
  int nch[3];          // tasks per axis
  int kpt;             // sort keys per task = OWN_ZS + order - 1
  int pk_in_w;         // unit-weight catalogue: the key word travels in the record's w slot
};

// Stencil base cell b (periodic, in [0, n)), the fraction s the window weights are
// evaluated from, and -- TSC only -- which branch of S/field.cpp:869-895 applies.
// Returns false when the particle is irregular along this axis.
template <int ORDER>
__device__ __forceinline__ bool own_axis_base(double loc, int n, int& b, double& s, int& hi) {
  hi = 0;
  if (!(loc >= 0.)) return false;
  const int idx = __double2int_rz(loc);
  if (idx >= n) return false;
  s = __dsub_rn(loc, (double)idx);
  if (ORDER == 1) {
    b = (s >= 0.5) ? ((idx == n - 1) ? 0 : idx + 1) : idx;
  } else if (ORDER == 2) {
    b = idx;
  } else if (ORDER == 3) {
    hi = !(s < 0.5);
    b = hi ? idx : idx - 1;
  } else {
    b = idx - 1;
  }
  if (b < 0) b += n;
  return true;
}

// The 1-D weights of the ORDER cells b, b+1, ... exactly as S/field.cpp computes them.
template <int ORDER>
__device__ __forceinline__ void window_weights(double s, int hi, double* win) {
  if (ORDER == 1) {
    win[0] = 1.;
  } else if (ORDER == 2) {
    win[0] = __dsub_rn(1., s);
    win[1] = s;
  } else if (ORDER == 3) {
    if (!hi) {
      const double a = __dsub_rn(0.5, s), c = __dadd_rn(0.5, s);
      win[0] = __dmul_rn(__dmul_rn(0.5, a), a);
      win[1] = __dsub_rn(0.75, __dmul_rn(s, s));
      win[2] = __dmul_rn(__dmul_rn(0.5, c), c);
    } else {
      s = __dsub_rn(1., s);
      const double a = __dsub_rn(0.5, s), c = __dadd_rn(0.5, s);
      win[0] = __dmul_rn(__dmul_rn(0.5, c), c);
      win[1] = __dsub_rn(0.75, __dmul_rn(s, s));
      win[2] = __dmul_rn(__dmul_rn(0.5, a), a);
    }
  } else {
    const double c6 = 1. / 6;
    const double u = __dsub_rn(1., s);
    win[0] = __dmul_rn(__dmul_rn(__dmul_rn(c6, u), u), u);
    win[1] = __dmul_rn(c6, __dadd_rn(
      __dsub_rn(4., __dmul_rn(__dmul_rn(6., s), s)),
      __dmul_rn(__dmul_rn(__dmul_rn(3., s), s), s)));
    win[2] = __dmul_rn(c6, __dadd_rn(
      __dsub_rn(4., __dmul_rn(__dmul_rn(6., u), u)),
      __dmul_rn(__dmul_rn(__dmul_rn(3., u), u), u)));
    win[3] = __dmul_rn(__dmul_rn(__dmul_rn(c6, s), s), s);
  }
}

// The chunks (tasks along one axis, P cells each) that the cells b .. b+ORDER-1 (mod n)
// fall into, each with the base RELATIVE to the chunk start: cell j of the stencil has
// relative index rb + j and belongs to that copy iff 0 <= rb + j < chunk extent.  At most
// two copies per axis are produced; a stencil that needs more (a chunk narrower than the
// stencil, or a mesh narrower than it) reports 0 and the particle is handled as irregular.
template <int ORDER>
__device__ __forceinline__ int own_axis_copies(int b, int n, int P, int& q0, int& r0,
                                               int& q1, int& r1) {
  int cnt = 0;
  q0 = r0 = q1 = r1 = 0;
#pragma unroll
  for (int j = 0; j < ORDER; j++) {
    int c = b + j;
    if (c >= n) c -= n;
    if (c >= n) return 0;                       // mesh narrower than the stencil
    const int qq = c / P, r = c - qq * P - j;
    if (cnt == 0) { q0 = qq; r0 = r; cnt = 1; }
    else if (qq == q0 && r == r0) {}
    else if (cnt == 1) { q1 = qq; r1 = r; cnt = 2; }
    else if (qq == q1 && r == r1) {}
    else return 0;
  }
  return cnt;
}

// key word of a copy: relative bases (+3) in x and y, first-plane key r, TSC branches.
__device__ __forceinline__ int own_pack(int rbx, int rby, int r, int hx, int hy, int hz) {
  return (rbx + 3) | ((rby + 3) << 5) | (r << 10) | (hx << 18) | (hy << 19) | (hz << 20);
}

// Pass 1 (cursor == nullptr): histogram of the copies over (task, r) keys and count of
// irregular particles.  Pass 2: place the copies at their sorted slots.
template <int ORDER>
__global__ void __launch_bounds__(256)
k_own_sort(CatView c, OwnDesc d, int* __restrict__ counts, int* __restrict__ cursor,
           int* __restrict__ irr_count, double4* __restrict__ rec, int* __restrict__ pk,
           int* __restrict__ src, int* __restrict__ irr_list) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < c.n;
       i += (long long)gridDim.x * blockDim.x) {
    int b[3], hi[3]; double s[3];
    const double px = c.x[i], py = c.y[i], pz = c.z[i];
    bool ok = own_axis_base<ORDER>(grid_loc(px, d.n[0], d.L[0], d.shifted), d.n[0], b[0], s[0], hi[0]);
    ok = own_axis_base<ORDER>(grid_loc(py, d.n[1], d.L[1], d.shifted), d.n[1], b[1], s[1], hi[1]) && ok;
    ok = own_axis_base<ORDER>(grid_loc(pz, d.n[2], d.L[2], d.shifted), d.n[2], b[2], s[2], hi[2]) && ok;
    int qx0, rx0, qx1, rx1, qy0, ry0, qy1, ry1, qz0, rz0, qz1, rz1;
    int cx = 0, cy = 0, cz = 0;
    if (ok) {
      cx = own_axis_copies<ORDER>(b[0], d.n[0], OWN_PX, qx0, rx0, qx1, rx1);
      cy = own_axis_copies<ORDER>(b[1], d.n[1], OWN_PY, qy0, ry0, qy1, ry1);
      cz = own_axis_copies<ORDER>(b[2], d.n[2], OWN_ZS, qz0, rz0, qz1, rz1);
    }
    if (cx * cy * cz == 0) {
      const int slot = atomicAdd(irr_count, 1);
      if (cursor) irr_list[slot] = (int)i;
      continue;
    }
    const double w = c.w ? c.w[i] : 1.;
    for (int ix = 0; ix < cx; ix++) {
      const int qx = ix ? qx1 : qx0, rx = ix ? rx1 : rx0;
      for (int iy = 0; iy < cy; iy++) {
        const int qy = iy ? qy1 : qy0, ry = iy ? ry1 : ry0;
        for (int iz = 0; iz < cz; iz++) {
          const int qz = iz ? qz1 : qz0, rz = iz ? rz1 : rz0;
          const int task = (qx * d.nch[1] + qy) * d.nch[2] + qz;
          const int r = rz + (ORDER - 1);
          const int key = task * d.kpt + r;
          if (!cursor) {
            atomicAdd(&counts[key], 1);
          } else {
            const int pos = atomicAdd(&cursor[key], 1);
            const int word = own_pack(rx, ry, r, hi[0], hi[1], hi[2]);
            rec[pos] = make_double4(s[0], s[1], s[2],
                                    d.pk_in_w ? __longlong_as_double((long long)word) : w);
            if (!d.pk_in_w) pk[pos] = word;
            if (src) src[pos] = (int)i;
          }
        }
      }
    }
  }
}

struct OwnView {
  const double4* rec; const int* pk; const int* src;   // pk null when it travels in rec.w
  const int* offsets;                                   // nkeys + 1
  const double* lx; const double* ly; const double* lz; // catalogue order, may be null
  const double* cw;                                     // catalogue order, may be null
};

template <int ORDER> struct OwnGeom {
  static constexpr int NA0 = (ORDER == 4) ? 2 : ORDER;          // stencil x-planes per pass
  static constexpr int NPASS = ORDER / NA0;
  static constexpr int LANES = NA0 * ORDER * ORDER;             // active lanes
  static constexpr int WXY_BYTES = ((ORDER * ORDER * 8 + 15) / 16) * 16;
  static constexpr int REC_RAW = 16 + 32 + WXY_BYTES;           // header, wz[4], wxy
  // odd multiple of 16 bytes: the 16-byte stage stores of a quarter-warp (one record per
  // lane) fall on distinct bank groups
  static constexpr int REC = ((REC_RAW / 16) % 2 == 1) ? REC_RAW : REC_RAW + 16;
  static constexpr size_t SMEM = sizeof(double) * OWN_RING_WORDS + (size_t)OWN_BATCH * REC;
};

// Shared-memory accesses by 32-bit shared address.  The walk loop is written with these so
// that (i) no generic-to-shared conversion is redone per step, (ii) the operands of the next
// particle are requested before the read-modify-write of the current one (program order of
// the volatile statements), and (iii) lanes whose cell lies outside the task are switched
// off by predication instead of a divergent branch.
__device__ __forceinline__ unsigned own_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
template <int IMM>
__device__ __forceinline__ int4 own_lds_v4(unsigned a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4+%5];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ double own_lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ double2 own_lds_v2f64(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(IMM));
  return v;
}
// ring[a + IMM] += v unless (inv & mask) != 0: LOP3 -> P, @P LDS, DADD, @P STS
template <int IMM>
__device__ __forceinline__ void own_add_unless(unsigned a, double v, int inv, int mask) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f64 d;\n\t"
               "and.b32 t, %2, %3;\n\tsetp.eq.s32 p, t, 0;\n\t"
               "@p ld.shared.f64 d, [%0+%4];\n\tadd.rn.f64 d, d, %1;\n\t"
               "@p st.shared.f64 [%0+%4], d;\n\t}"
               :: "r"(a), "d"(v), "r"(inv), "r"(mask), "n"(IMM) : "memory");
}
// keep a per-lane constant in its register (the compiler otherwise rebuilds it from the
// lane index inside the walk loop)
__device__ __forceinline__ void own_pin(unsigned& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void own_pin(int& x) { asm volatile("" : "+r"(x)); }

// Store (or add) the planes [4 grp, 4 grp + 4) of the task to the mesh and clear their
// ring slots.  Out of line: it runs once per ~100 particles and would otherwise be
// replicated into every copy of the unrolled walk loop.
template <bool COMPLEX>
__device__ __noinline__ void own_flush(double* ring, int grp, int lane, int pxe, int pye,
                                       int zse, long long cell0, int n1, int n2,
                                       int comp, int accumulate, double* __restrict__ mesh) {
  const int slot0 = (OWN_G * grp) % OWN_RING;
  const int nv = min(OWN_G, zse - OWN_G * grp);
  const bool vec_ok = (n2 % 2) == 0;
#pragma unroll 4
  for (int it = 0; it < OWN_PX * OWN_PY / 32; it++) {
    const int p = it * 32 + lane;
    const int xi = p / OWN_PY, yi = p - xi * OWN_PY;
    if (xi >= pxe || yi >= pye) continue;
    double* rp = ring + xi * OWN_PITCH_X + yi * OWN_PITCH_Y + slot0;
    double2 v0 = *reinterpret_cast<double2*>(rp);
    double2 v1 = *reinterpret_cast<double2*>(rp + 2);
    *reinterpret_cast<double2*>(rp) = make_double2(0., 0.);
    *reinterpret_cast<double2*>(rp + 2) = make_double2(0., 0.);
    const long long cell = cell0 + ((long long)xi * n1 + yi) * n2 + OWN_G * grp;
    if (!COMPLEX && vec_ok && nv == OWN_G) {
      double2* mp = reinterpret_cast<double2*>(mesh + cell);
      if (accumulate) {
        const double2 o0 = mp[0], o1 = mp[1];
        v0.x += o0.x; v0.y += o0.y; v1.x += o1.x; v1.y += o1.y;
      }
      mp[0] = v0; mp[1] = v1;
    } else {
      const double v[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int k = 0; k < OWN_G; k++) {
        if (k < nv) {
          double* mp = COMPLEX ? mesh + 2 * (cell + k) + comp : mesh + cell + k;
          *mp = accumulate ? *mp + v[k] : v[k];
        }
      }
    }
  }
}

// Operands of one walk step (one particle).
struct OwnStep { int4 hdr; double wz; double2 wxy; };

// One warp per task (and, for COMPLEX meshes, per component: the real and the imaginary
// parts of the weights are spread by two tasks onto the interleaved mesh).  Seven such
// one-warp CTAs share an SM (30 KB of shared memory each).  HEAVY: weights that need the
// lines of sight or the custom weight column (kept out of the unit-weight instantiation).
template <int ORDER, bool COMPLEX, bool HEAVY>
__global__ void __launch_bounds__(32)
k_assign_own(OwnView v, GridDesc g, OwnDesc d, int kind, YlmCoef yc, double scale,
             int accumulate, double* __restrict__ mesh) {
  typedef OwnGeom<ORDER> G;
  extern __shared__ __align__(16) unsigned char own_smem[];
  double* ring = reinterpret_cast<double*>(own_smem);
  unsigned char* stage = own_smem + sizeof(double) * OWN_RING_WORDS;
  const int lane = threadIdx.x;
  const int comp = COMPLEX ? (blockIdx.x & 1) : 0;
  const int task = COMPLEX ? (blockIdx.x >> 1) : blockIdx.x;
  const int qz = task % d.nch[2], qy = (task / d.nch[2]) % d.nch[1], qx = task / (d.nch[2] * d.nch[1]);
  const int X0 = qx * OWN_PX, Y0 = qy * OWN_PY, Z0 = qz * OWN_ZS;
  const int pxe = min(OWN_PX, g.n[0] - X0), pye = min(OWN_PY, g.n[1] - Y0), zse = min(OWN_ZS, g.n[2] - Z0);
  const int lo = v.offsets[task * d.kpt], hi = v.offsets[(task + 1) * d.kpt];
  if (lo == hi && accumulate) return;
  const long long cell0 = ((long long)X0 * g.n[1] + Y0) * g.n[2] + Z0;
  const int n1 = g.n[1], n2 = g.n[2];

  for (int t = lane; t < OWN_RING_WORDS / 2; t += 32) {
    reinterpret_cast<double2*>(ring)[t] = make_double2(0., 0.);
  }
  // walking role: lane -> cell (a0 [+ 2 in the second PCS pass], b, cz) of the stencil
  const bool lane_on = lane < G::LANES;
  const int a0 = lane_on ? lane / (ORDER * ORDER) : 0;
  const int b = lane_on ? (lane / ORDER) % ORDER : 0;
  const int cz = lane_on ? lane % ORDER : 0;
  // the lane's cell is skipped when the particle's header flags its x-plane, y-row or plane
  // as outside the task; idle lanes test the always-set bit 31
  int mask0 = lane_on ? ((1 << a0) | (1 << (4 + b)) | (1 << (8 + cz))) : (int)0x80000000;
  int mask1 = lane_on ? ((1 << (a0 + 2)) | (1 << (4 + b)) | (1 << (8 + cz))) : (int)0x80000000;
  const unsigned ring_u = own_smem_u32(ring), stage_u = own_smem_u32(stage);
  unsigned cell_u = ring_u + 8u * (unsigned)(a0 * OWN_PITCH_X + b * OWN_PITCH_Y);
  unsigned wz_u = stage_u + 16u + 8u * (unsigned)cz;
  unsigned wxy_u = stage_u + 48u
    + (ORDER == 4 ? 16u * (unsigned)(a0 * 4 + b) : 8u * (unsigned)(a0 * ORDER + b));
  unsigned slot_sel = 0x4440u | (unsigned)cz;    // prmt: byte cz of the slot word
  own_pin(mask0); own_pin(mask1); own_pin(cell_u); own_pin(wz_u); own_pin(wxy_u); own_pin(slot_sel);
  const int ngroups = (zse + OWN_G - 1) / OWN_G;
  int next_flush = 0;

  // staging role: particle `lane` of the batch
  double4 rec = make_double4(0., 0., 0., 0.);
  int pkw = 0, srcw = 0;
  if (lo + lane < hi) {
    rec = v.rec[lo + lane];
    pkw = v.pk ? v.pk[lo + lane] : (int)__double_as_longlong(rec.w);
    if (HEAVY && v.src) srcw = v.src[lo + lane];
  }
  __syncwarp();

  for (int base = lo; base < hi; base += OWN_BATCH) {
    // ---- stage ---------------------------------------------------------------------
    if (base + lane < hi) {
      unsigned char* sp = stage + lane * G::REC;
      double wx[ORDER], wy[ORDER];
      window_weights<ORDER>(rec.x, (pkw >> 18) & 1, wx);
      window_weights<ORDER>(rec.y, (pkw >> 19) & 1, wy);
      const double wgt = v.pk ? rec.w : 1.;
      double wt;
      if (!HEAVY) {
        wt = comp ? 0. : (kind == TRVB_W_W ? wgt : 1.);
      } else if (kind == TRVB_W_CUSTOM) {
        wt = v.cw[2 * (long long)srcw + comp];
      } else {
        cplx y; y.re = 1.; y.im = 0.;
        if (!(yc.ell == 0 && yc.m == 0)) y = ylm_eval(yc, v.lx[srcw], v.ly[srcw], v.lz[srcw]);
        if (kind == TRVB_W_YLM_W) {
          wt = comp ? y.im * wgt : y.re * wgt;
        } else {
          const double w2 = wgt * wgt;
          wt = comp ? -y.im * w2 : y.re * w2;
        }
      }
      const double bw = __dmul_rn(scale, wt);
      double* wxyp = reinterpret_cast<double*>(sp + 48);
#pragma unroll
      for (int ax = 0; ax < ORDER; ax++) {
        const double wa = __dmul_rn(bw, wx[ax]);
#pragma unroll
        for (int bb = 0; bb < ORDER; bb++) {
          // PCS: rows a and a + 2 adjacent, so one 16-byte load serves both passes
          const int at = (ORDER == 4) ? ((ax & 1) * 4 + bb) * 2 + (ax >> 1) : ax * ORDER + bb;
          wxyp[at] = __dmul_rn(wa, wy[bb]);
        }
      }
      double wz[4] = {0., 0., 0., 0.};
      window_weights<ORDER>(rec.z, (pkw >> 20) & 1, wz);
      const int rbx = (pkw & 31) - 3, rby = ((pkw >> 5) & 31) - 3;
      const int zi0 = ((pkw >> 10) & 255) - (ORDER - 1);
      const int s0 = (zi0 + 2 * OWN_RING) % OWN_RING;
      int inv = (int)0x80000000, slots = 0;
#pragma unroll
      for (int t = 0; t < ORDER; t++) {
        if ((unsigned)(rbx + t) >= (unsigned)pxe) inv |= 1 << t;
        if ((unsigned)(rby + t) >= (unsigned)pye) inv |= 1 << (4 + t);
        if ((unsigned)(zi0 + t) >= (unsigned)zse) inv |= 1 << (8 + t);
        int sl = s0 + t;
        if (sl >= OWN_RING) sl -= OWN_RING;
        slots |= (8 * sl) << (8 * t);
      }
      int4 hdr;
      hdr.x = 8 * (rbx * OWN_PITCH_X + rby * OWN_PITCH_Y);   // byte offset of cell (a, b) = (0, 0)
      hdr.y = inv;
      // number of flush groups that are complete before this particle: the planes below
      // zi0 are; group q holds the planes 4 q .. 4 q + 3
      hdr.z = zi0 > 0 ? zi0 / OWN_G : 0;
      hdr.w = slots;
      *reinterpret_cast<int4*>(sp) = hdr;
      *reinterpret_cast<double2*>(sp + 16) = make_double2(wz[0], wz[1]);
      *reinterpret_cast<double2*>(sp + 32) = make_double2(wz[2], wz[3]);
    }
    // next batch's records on their way while this one is walked
    {
      const int nxt = base + OWN_BATCH + lane;
      double4 rec_n = make_double4(0., 0., 0., 0.);
      int pk_n = 0, src_n = 0;
      if (nxt < hi) {
        rec_n = v.rec[nxt];
        pk_n = v.pk ? v.pk[nxt] : (int)__double_as_longlong(rec_n.w);
        if (HEAVY && v.src) src_n = v.src[nxt];
      }
      rec = rec_n; pkw = pk_n; srcw = src_n;
    }
    __syncwarp();

    // ---- walk: one particle per step, lanes = cells; two steps per iteration with the
    // operands of step s + 1 requested before the read-modify-write of step s ---------
    const int count = min(OWN_BATCH, hi - base);
    auto load = [&](unsigned hp, unsigned zp, unsigned xp, OwnStep& o) {
      o.hdr = own_lds_v4<0>(hp);
      o.wz = own_lds_f64<0>(zp);
      if (ORDER == 4) o.wxy = own_lds_v2f64<0>(xp);
      else { o.wxy.x = own_lds_f64<0>(xp); o.wxy.y = 0.; }
    };
    auto step = [&](const OwnStep& o) {
      // every plane below the particle's first one is complete: store those groups
      if (next_flush < o.hdr.z) {
        const int upto = min(o.hdr.z, ngroups);
        __syncwarp();
        for (; next_flush < upto; next_flush++) {
          own_flush<COMPLEX>(ring, next_flush, lane, pxe, pye, zse, cell0, n1, n2, comp,
                             accumulate, mesh);
        }
        next_flush = max(next_flush, o.hdr.z);
        __syncwarp();
      }
      const unsigned addr = cell_u + (unsigned)o.hdr.x + __byte_perm((unsigned)o.hdr.w, 0u, slot_sel);
      own_add_unless<0>(addr, __dmul_rn(o.wxy.x, o.wz), o.hdr.y, mask0);
      if (ORDER == 4) {
        own_add_unless<16 * OWN_PITCH_X>(addr, __dmul_rn(o.wxy.y, o.wz), o.hdr.y, mask1);
      }
    };
    OwnStep A, B;
    unsigned hp = stage_u, zp = wz_u, xp = wxy_u;
    load(hp, zp, xp, A);
#pragma unroll 1
    for (int s = 0; s < count; s += 2) {
      const bool two = s + 1 < count;
      if (two) load(hp + G::REC, zp + G::REC, xp + G::REC, B);
      step(A);
      hp += 2 * G::REC; zp += 2 * G::REC; xp += 2 * G::REC;
      if (s + 2 < count) load(hp, zp, xp, A);
      if (two) step(B);
    }
    __syncwarp();
  }
  for (; next_flush < ngroups; next_flush++) {
    own_flush<COMPLEX>(ring, next_flush, lane, pxe, pye, zse, cell0, n1, n2, comp, accumulate, mesh);
  }
}

// Irregular particles (an index outside [0, n) on some axis): the reference's own index
// arithmetic and guard (S/field.cpp:1042), added with REDs after the owner pass.
template <int ORDER, bool COMPLEX>
__global__ void __launch_bounds__(128)
k_assign_irregular(CatView c, const int* __restrict__ list, int nlist, GridDesc g, int shifted,
                   int kind, int L, int M, double scale, double* __restrict__ mesh) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nlist; t += gridDim.x * blockDim.x) {
    const long long i = list[t];
    int ijk[3][ORDER];
    double win[3][ORDER];
    window_1d<ORDER>(grid_loc(c.x[i], g.n[0], g.L[0], shifted), g.n[0], ijk[0], win[0]);
    window_1d<ORDER>(grid_loc(c.y[i], g.n[1], g.L[1], shifted), g.n[1], ijk[1], win[1]);
    window_1d<ORDER>(grid_loc(c.z[i], g.n[2], g.L[2], shifted), g.n[2], ijk[2], win[2]);
    const cplx wt = particle_weight(c, i, kind, L, M);
    const double bre = __dmul_rn(scale, wt.re);
    const double bim = COMPLEX ? __dmul_rn(scale, wt.im) : 0.;
    for (int a = 0; a < ORDER; a++) {
      const double wa_re = __dmul_rn(bre, win[0][a]);
      const double wa_im = COMPLEX ? __dmul_rn(bim, win[0][a]) : 0.;
      for (int b = 0; b < ORDER; b++) {
        const double wb_re = __dmul_rn(wa_re, win[1][b]);
        const double wb_im = COMPLEX ? __dmul_rn(wa_im, win[1][b]) : 0.;
        const long long row = ((long long)ijk[0][a] * g.n[1] + ijk[1][b]) * g.n[2];
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
