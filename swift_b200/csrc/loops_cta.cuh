/*
 * loops_cta.cuh - CTA-cooperative neighbour loops ("cell-pair tiles").
 *
 * One CTA of 8 warps owns up to 64 TARGET particles of one target cell. The
 * source cells of the group's items are staged ONCE per CTA, in batches, into
 * a shared-memory pool (frame floats, payload, global index), together with
 * an axis-aligned bounding box per OCTET of 8 consecutive sources. Each warp
 * then works for 8 targets with the lane layout (t = lane & 7, s = lane >> 3):
 *
 *   CULL      lane-parallel box-box distance test of the warp's 8-target box
 *             against every staged octet box (ballot -> accepted octets);
 *   TEST      per accepted octet every lane tests its target against sources
 *             2s and 2s+1 of the octet (FMA r2, inflated radius, sorted-axis
 *             key condition) and appends candidates to its private sub-list;
 *   INTERACT  the 4 sub-lists of a target are merged on the fly and drained by
 *             its 4 lanes; each candidate is re-evaluated with the reference's
 *             exact arithmetic (see loops.cuh) before the interaction. The 4
 *             partial sums of a target are combined by two shuffles at the end.
 *
 * Compared with the warp-per-64-targets kernels of loops.cuh the culling
 * granularity is 8 targets x 8 sources instead of 64 targets x 32 sources and
 * the staging cost is shared by 64 targets.
 */
#ifndef SWIFTGPU_LOOPS_CTA_CUH
#define SWIFTGPU_LOOPS_CTA_CUH

#include "loops.cuh"

namespace swiftgpu {

#define CTA_THREADS 256
#define CTA_TARGETS 64
#define POOL 1024       /* staged sources per batch */
#define NOCT (POOL / 8) /* octets per batch */
#define BATCH_ITEMS 16  /* item fragments per batch */
#define SUBCAP 24       /* sub-list capacity per lane */

/* Constants of one staged item fragment. */
struct __align__(16) ItemInfoS {
  double ot[3]; /* drain: subtracted from the target double */
  double fs[3]; /* frame origin of the staged source floats */
  double rshift, lim_a, lim_b;
  float d[3];   /* cull: item-frame position of the target-cell origin */
  float dx_max; /* pair: ci.dx_max_sort + cj.dx_max_sort; subset pair: cj.dx_max_sort */
  float margin; /* cull / prefilter widening */
  int32_t mode, sid, kc, dbl;
  int32_t min_depth, max_depth;
  int32_t src_first; /* global index of the fragment's first source */
  int32_t src_n;     /* sources in the fragment */
  int32_t pool_off;  /* first pool slot */
  int32_t pad_;
};

template <int NP, bool STAGE_D>
struct CtaSmem {
  static constexpr int kF = 0;
  static constexpr int kP = kF + POOL * 16;
  static constexpr int kGI = kP + NP * POOL * 16;
  static constexpr int kD = kGI + POOL * 4;
  static constexpr int kOB = kD + (STAGE_D ? POOL * 24 : 0);
  static constexpr int kTP = kOB + NOCT * 32;
  static constexpr int kII = kTP + BATCH_ITEMS * CTA_TARGETS * 16;
  static constexpr int kTS = kII + BATCH_ITEMS * (int)sizeof(ItemInfoS);
  static constexpr int kList = kTS + CTA_TARGETS * 48;
  static constexpr int kCtl = kList + SUBCAP * CTA_THREADS * 2;
  static constexpr int kBytes = kCtl + 64;
};

/* Target data every thread may need while filling the TP table. */
struct __align__(16) TgtS {
  double x, y, z;
  float thg, thg2;
  int32_t depth, valid;
  int32_t pad0_, pad1_;
};
static_assert(sizeof(TgtS) == 48, "TgtS");

template <int LOOP, bool SUBSET>
__global__ void __launch_bounds__(CTA_THREADS) k_cta1(const LoopArgs A) {
  constexpr int NP = (LOOP == LOOP_GRADIENT ? 2 : 1);
  typedef CtaSmem<NP, SUBSET> SM;
  extern __shared__ __align__(16) char smem[];
  float4 *const sF = (float4 *)(smem + SM::kF);
  float4 *const sP0 = (float4 *)(smem + SM::kP);
  float4 *const sP1 = sP0 + POOL;
  int32_t *const sGI = (int32_t *)(smem + SM::kGI);
  double *const sD = (double *)(smem + SM::kD);
  float4 *const sOBlo = (float4 *)(smem + SM::kOB);
  float4 *const sOBhi = sOBlo + NOCT;
  float4 *const sTP = (float4 *)(smem + SM::kTP);
  ItemInfoS *const sII = (ItemInfoS *)(smem + SM::kII);
  TgtS *const sTS = (TgtS *)(smem + SM::kTS);
  uint16_t *const sList = (uint16_t *)(smem + SM::kList);
  int32_t *const sCtl = (int32_t *)(smem + SM::kCtl); /* [0] fragments in batch, [1] octets in batch, [2] next item, [3] next src offset */

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int t8 = lane & 7;
  const int s4 = lane >> 3;
  const int task = blockIdx.x;

  const int g = A.task_group[task];
  const int chunk = A.task_chunk[task];
  const int nt = A.tgt_count[g];
  if (chunk * CTA_TARGETS >= nt) return;
  const Group G = A.groups[g];
  const DevCell tcell = A.cells[G.tcell];
  const double T0x = tcell.loc[0], T0y = tcell.loc[1], T0z = tcell.loc[2];

  /* ---- my target (4 lanes share one) ---- */
  const int slot_t = chunk * CTA_TARGETS + warp * 8 + t8;
  const bool tvalid = slot_t < nt;
  const int ti = tvalid ? A.tgt_list[A.tgt_first[g] + slot_t] : -1;
  double tx = 0., ty = 0., tz = 0.;
  float th = 1.f, tvx = 0.f, tvy = 0.f, tvz = 0.f, tu = 0.f, tcs = 0.f;
  int tdepth = 0;
  if (tvalid) {
    tx = A.x[3 * (size_t)ti];
    ty = A.x[3 * (size_t)ti + 1];
    tz = A.x[3 * (size_t)ti + 2];
    th = A.h[ti];
    const float4 q = A.mv[ti];
    tvx = q.y;
    tvy = q.z;
    tvz = q.w;
    tdepth = A.depth_h[ti];
    if (LOOP == LOOP_GRADIENT) {
      tu = A.fq2[ti].z;
      tcs = A.fq1[ti].w;
    }
  }
  const float thg2 = hg2_exact(th);
  const float th_inv = 1.f / th;
  const float thg = __fmul_rn(th, KERNEL_GAMMA);
  if (s4 == 0) {
    TgtS ts;
    ts.x = tx;
    ts.y = ty;
    ts.z = tz;
    ts.thg = thg;
    ts.thg2 = thg2;
    ts.depth = tdepth;
    ts.valid = tvalid ? 1 : 0;
    ts.pad0_ = ts.pad1_ = 0;
    sTS[warp * 8 + t8] = ts;
  }
  /* the warp's target box, relative to the target cell origin, and its reach */
  float blo[3], bhi[3], rmax;
  {
    const float qx = dsubf(tx, T0x), qy = dsubf(ty, T0y), qz = dsubf(tz, T0z);
    blo[0] = tvalid ? qx : 3.0e30f;
    blo[1] = tvalid ? qy : 3.0e30f;
    blo[2] = tvalid ? qz : 3.0e30f;
    bhi[0] = tvalid ? qx : -3.0e30f;
    bhi[1] = tvalid ? qy : -3.0e30f;
    bhi[2] = tvalid ? qz : -3.0e30f;
    rmax = tvalid ? thg : 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      blo[k] = warp_min(blo[k]);
      bhi[k] = warp_max(bhi[k]);
    }
    rmax = warp_max(rmax);
  }
  const bool warp_has_targets = __any_sync(FULL_MASK, tvalid);

  DensityAcc dacc;
  dacc.zero();
  GradientAcc gacc;
  gacc.v_sig = 0.f;
  gacc.laplace_u = 0.f;
  gacc.alpha_max = 0.f;
  int nhit = 0;
  int ntests = 0;
  int nsub = 0; /* entries in my sub-list */
  uint16_t *const mylist = sList + tid;

  /* ---- INTERACT: merge the 4 sub-lists of each target and drain ---- */
  auto drain = [&]() {
    __syncwarp();
    const int n0 = __shfl_sync(FULL_MASK, nsub, t8);
    const int n1 = __shfl_sync(FULL_MASK, nsub, t8 + 8);
    const int n2 = __shfl_sync(FULL_MASK, nsub, t8 + 16);
    const int n3 = __shfl_sync(FULL_MASK, nsub, t8 + 24);
    const int c1 = n0 + n1, c2 = c1 + n2, total = c2 + n3;
    const int steps = (__reduce_max_sync(FULL_MASK, total) + 3) >> 2;
    for (int j = 0; j < steps; j++) {
      const int m = 4 * j + s4;
      const bool act = m < total;
      const int q = (m >= n0) + (m >= c1) + (m >= c2);
      const int base = q == 0 ? 0 : (q == 1 ? n0 : (q == 2 ? c1 : c2));
      const int kk = act ? m - base : 0;
      const int slot = act ? (int)sList[kk * CTA_THREADS + warp * 32 + t8 + 8 * q] : 0;
      const int item = __float_as_int(sOBlo[slot >> 3].w);
      const ItemInfoS &ii = sII[item];
      const int gi = sGI[slot];
      const float4 s = sF[slot];
      const bool dbl = ii.dbl != 0;
      double sxd = 0., syd = 0., szd = 0.;
      if (act && dbl) {
        if (SUBSET) {
          sxd = sD[slot];
          syd = sD[POOL + slot];
          szd = sD[2 * POOL + slot];
        } else {
          sxd = A.x[3 * (size_t)gi];
          syd = A.x[3 * (size_t)gi + 1];
          szd = A.x[3 * (size_t)gi + 2];
        }
      }
      const float spx = dbl ? 0.f : s.x, spy = dbl ? 0.f : s.y, spz = dbl ? 0.f : s.z;
      const float dx = __fsub_rn(dsubf(__dsub_rn(tx, ii.ot[0]), sxd), spx);
      const float dy = __fsub_rn(dsubf(__dsub_rn(ty, ii.ot[1]), syd), spy);
      const float dz = __fsub_rn(dsubf(__dsub_rn(tz, ii.ot[2]), szd), spz);
      const float r2 = r2_exact(dx, dy, dz);
      if (act && (r2 < thg2) && (gi != ti)) {
        const float4 f0 = sP0[slot];
        if (LOOP == LOOP_DENSITY) {
          iact_density(dacc, r2, dx, dy, dz, th_inv, tvx, tvy, tvz, f0.x, f0.y, f0.z, f0.w);
        } else {
          const float4 f1 = sP1[slot];
          iact_gradient(gacc, r2, dx, dy, dz, th, tvx, tvy, tvz, tu, tcs, f0.x, f0.y, f0.z, f0.w,
                        f1.x, f1.y, f1.z, f1.w, A.a2_Hubble);
        }
        nhit++;
      }
    }
    nsub = 0;
    __syncwarp();
  };

  if (tid == 0) {
    sCtl[2] = 0; /* next item */
    sCtl[3] = 0; /* source offset inside it */
  }
  __syncthreads();

  for (;;) {
    /* ---------------- batch: which item fragments fit the pool ---------------- */
    if (warp == 0) {
      /* lane i looks at item it0 + i: one round of loads, then a warp scan */
      const int it0 = sCtl[2];
      const int soff_in = sCtl[3];
      const int it = it0 + lane;
      int cnt = 0;
      if (it < G.item_count) {
        const Item I = A.items[G.item_first + it];
        cnt = A.cells[I.scell].count - (lane == 0 ? soff_in : 0);
      }
      const int padded = (cnt + 7) & ~7;
      int incl = padded;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += v;
      }
      const bool fits = (it < G.item_count) && (incl <= POOL) && (lane < BATCH_ITEMS);
      const unsigned fm = __ballot_sync(FULL_MASK, fits);
      /* fragments = leading run of items that fit */
      int nb = __ffs(~fm) - 1;
      if (nb < 0) nb = 32;
      int used = 0;
      if (nb == 0 && it0 < G.item_count) {
        /* a single source cell larger than the pool: take a slice of it */
        nb = 1;
        if (lane == 0) {
          sII[0].pool_off = 0;
          sII[0].src_first = soff_in;
          sII[0].src_n = POOL;
          sII[0].pad_ = it0;
          sCtl[2] = it0;
          sCtl[3] = soff_in + POOL;
        }
        used = POOL;
      } else {
        if (lane < nb) {
          sII[lane].pool_off = incl - padded;
          sII[lane].src_first = (lane == 0 ? soff_in : 0);
          sII[lane].src_n = cnt;
          sII[lane].pad_ = it;
        }
        used = __shfl_sync(FULL_MASK, incl, nb > 0 ? nb - 1 : 0);
        if (nb == 0) used = 0;
        if (lane == 0) {
          sCtl[2] = it0 + nb;
          sCtl[3] = 0;
        }
      }
      if (lane == 0) {
        sCtl[0] = nb;
        sCtl[1] = used >> 3;
      }
    }
    __syncthreads();
    const int nb = sCtl[0];
    const int noct = sCtl[1];
    if (nb == 0) break;

    /* ---------------- item constants (one thread per fragment) ---------------- */
    if (tid < nb) {
      ItemInfoS ii = sII[tid];
      const Item I = A.items[G.item_first + ii.pad_];
      const DevCell sc = A.cells[I.scell];
      const int mode = I.mode, sid = I.sid;
      const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
      ii.mode = mode;
      ii.sid = sid;
      ii.min_depth = I.min_depth;
      ii.max_depth = I.max_depth;
      ii.src_first = sc.first + ii.src_first;
      ii.rshift = ii.lim_a = ii.lim_b = 0.;
      ii.dx_max = 0.f;
      ii.kc = 0;
      ii.dbl = 0;
      ii.margin = 2.0e-6f * sc.width + 2.0e-6f * tcell.width;
      double otx = 0., oty = 0., otz = 0., fsx, fsy, fsz;
      if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
        const DevCell &ci = (mode == MODE_PAIR_L) ? tcell : sc;
        const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tcell;
        ii.rshift = __dadd_rn(
            __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
            __dmul_rn(shz, c_runner_shift[sid][2]));
        const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
        ii.dx_max = __fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
        const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                     oiz = __dadd_rn(cj.loc[2], shz);
        if (mode == MODE_PAIR_L) {
          ii.lim_a = __dsub_rn((double)__fmul_rn(fminf(h_max_lim, ci.h_max_active), KERNEL_GAMMA), ii.rshift);
          const int j0 = cj.first + (int)A.sort_idx[sort_offset(cj, sid)];
          ii.lim_b = (double)sort_key(A.x[3 * (size_t)j0], A.x[3 * (size_t)j0 + 1],
                                      A.x[3 * (size_t)j0 + 2], sid);
          otx = oix; oty = oiy; otz = oiz;
          fsx = cj.loc[0]; fsy = cj.loc[1]; fsz = cj.loc[2];
          ii.kc = 1;
        } else {
          ii.lim_a = (double)__fmul_rn(fminf(h_max_lim, cj.h_max_active), KERNEL_GAMMA);
          const int i1 = ci.first + (int)A.sort_idx[sort_offset(ci, sid) + ci.count - 1];
          ii.lim_b = __dsub_rn((double)sort_key(A.x[3 * (size_t)i1], A.x[3 * (size_t)i1 + 1],
                                                A.x[3 * (size_t)i1 + 2], sid),
                               ii.rshift);
          otx = cj.loc[0]; oty = cj.loc[1]; otz = cj.loc[2];
          fsx = oix; fsy = oiy; fsz = oiz;
          ii.kc = 2;
        }
        ii.d[0] = dsubf(T0x, otx);
        ii.d[1] = dsubf(T0y, oty);
        ii.d[2] = dsubf(T0z, otz);
      } else if (mode == MODE_SUB_SELF) {
        otx = fsx = sc.loc[0];
        oty = fsy = sc.loc[1];
        otz = fsz = sc.loc[2];
        ii.d[0] = dsubf(T0x, otx);
        ii.d[1] = dsubf(T0y, oty);
        ii.d[2] = dsubf(T0z, otz);
      } else {
        ii.dbl = 1;
        if (mode != MODE_SELF) {
          otx = shx; oty = shy; otz = shz;
          ii.kc = (mode == MODE_SUB_PAIR) ? 1 : 2;
          ii.dx_max = sc.dx_max_sort;
        }
        fsx = sc.loc[0]; fsy = sc.loc[1]; fsz = sc.loc[2];
        ii.d[0] = dsubf(__dsub_rn(T0x, otx), fsx);
        ii.d[1] = dsubf(__dsub_rn(T0y, oty), fsy);
        ii.d[2] = dsubf(__dsub_rn(T0z, otz), fsz);
      }
      ii.ot[0] = otx; ii.ot[1] = oty; ii.ot[2] = otz;
      ii.fs[0] = fsx; ii.fs[1] = fsy; ii.fs[2] = fsz;
      sII[tid] = ii;
    }
    __syncthreads();

    /* ---------------- TP table: per (target, fragment) prefilter position + key threshold ---------------- */
    for (int e = tid; e < nb * CTA_TARGETS; e += CTA_THREADS) {
      const int tt = e & (CTA_TARGETS - 1), fi = e >> 6;
      const ItemInfoS &ii = sII[fi];
      const TgtS ts = sTS[tt];
      bool part = ts.valid && ts.depth >= ii.min_depth && ts.depth <= ii.max_depth;
      float thr = 0.f, px, py, pz;
      const int mode = ii.mode, sid = ii.sid;
      if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
        const float tkey = sort_key(ts.x, ts.y, ts.z, sid);
        if (mode == MODE_PAIR_L) {
          const bool in_loop = __dadd_rn(__dadd_rn((double)tkey, ii.lim_a), (double)ii.dx_max) > ii.lim_b;
          const double di = __dsub_rn((double)__fadd_rn(__fadd_rn(tkey, ts.thg), ii.dx_max), ii.rshift);
          part = part && in_loop && !(di < ii.lim_b);
          thr = __double2float_ru(di);
        } else {
          const bool in_loop = __dsub_rn(__dsub_rn((double)tkey, ii.lim_a), (double)ii.dx_max) < ii.lim_b;
          const double dj = __dadd_rn((double)__fsub_rn(__fsub_rn(tkey, ts.thg), ii.dx_max), ii.rshift);
          part = part && in_loop && !(__dsub_rn(dj, ii.rshift) > ii.lim_b);
          thr = __double2float_rd(dj);
        }
        px = dsubf(ts.x, ii.ot[0]);
        py = dsubf(ts.y, ii.ot[1]);
        pz = dsubf(ts.z, ii.ot[2]);
      } else if (mode == MODE_SUB_SELF) {
        px = dsubf(ts.x, ii.ot[0]);
        py = dsubf(ts.y, ii.ot[1]);
        pz = dsubf(ts.z, ii.ot[2]);
      } else {
        const double tdx = __dsub_rn(ts.x, ii.ot[0]), tdy = __dsub_rn(ts.y, ii.ot[1]),
                     tdz = __dsub_rn(ts.z, ii.ot[2]);
        px = dsubf(tdx, ii.fs[0]);
        py = dsubf(tdy, ii.fs[1]);
        pz = dsubf(tdz, ii.fs[2]);
        if (mode != MODE_SELF) {
          const float f0 = (mode == MODE_SUB_PAIR) ? __fadd_rn(ts.thg, ii.dx_max) : __fsub_rn(-ts.thg, ii.dx_max);
          const double di =
              __dadd_rn(__dadd_rn(__dadd_rn((double)f0, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                  __dmul_rn(tdy, c_runner_shift[sid][1])),
                        __dmul_rn(tdz, c_runner_shift[sid][2]));
          thr = (mode == MODE_SUB_PAIR) ? __double2float_ru(di) : __double2float_rd(di);
        }
      }
      if (!part) px = 3.0e30f;
      sTP[fi * CTA_TARGETS + tt] = make_float4(px, py, pz, thr);
    }

    /* ---------------- stage the sources of the batch ---------------- */
    for (int slot = tid; slot < ((noct * 8 + 31) & ~31); slot += CTA_THREADS) {
      int fi = 0;
#pragma unroll 1
      for (int k = 1; k < nb; k++)
        if (slot >= sII[k].pool_off) fi = k;
      const ItemInfoS &ii = sII[fi];
      const int k = slot - ii.pool_off;
      const bool valid = k < ii.src_n;
      float fx = 0.f, fy = 0.f, fz = 0.f, key = 0.f;
      int sj = -1;
      if (valid) {
        sj = ii.src_first + k;
        const double sx = A.x[3 * (size_t)sj], sy = A.x[3 * (size_t)sj + 1], sz = A.x[3 * (size_t)sj + 2];
        fx = dsubf(sx, ii.fs[0]);
        fy = dsubf(sy, ii.fs[1]);
        fz = dsubf(sz, ii.fs[2]);
        if (ii.kc) key = sort_key(sx, sy, sz, ii.sid);
        sP0[slot] = A.mv[sj];
        if (LOOP == LOOP_GRADIENT) {
          const float4 q1 = A.fq1[sj];
          const float4 q2 = A.fq2[sj];
          const float4 q3 = A.fq3[sj];
          sP1[slot] = make_float4(q2.z /*u*/, q1.x /*rho*/, q1.w /*cs*/, q3.x /*alpha*/);
        }
        if (SUBSET) {
          sD[slot] = sx;
          sD[POOL + slot] = sy;
          sD[2 * POOL + slot] = sz;
        }
      }
      sF[slot] = valid ? make_float4(fx, fy, fz, key) : make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
      sGI[slot] = sj;
      /* octet box: 8 consecutive threads hold one octet */
      float lo0 = valid ? fx : 3.0e30f, lo1 = valid ? fy : 3.0e30f, lo2 = valid ? fz : 3.0e30f;
      float hi0 = valid ? fx : -3.0e30f, hi1 = valid ? fy : -3.0e30f, hi2 = valid ? fz : -3.0e30f;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        lo0 = fminf(lo0, __shfl_xor_sync(FULL_MASK, lo0, o));
        lo1 = fminf(lo1, __shfl_xor_sync(FULL_MASK, lo1, o));
        lo2 = fminf(lo2, __shfl_xor_sync(FULL_MASK, lo2, o));
        hi0 = fmaxf(hi0, __shfl_xor_sync(FULL_MASK, hi0, o));
        hi1 = fmaxf(hi1, __shfl_xor_sync(FULL_MASK, hi1, o));
        hi2 = fmaxf(hi2, __shfl_xor_sync(FULL_MASK, hi2, o));
      }
      if ((slot & 7) == 0) {
        sOBlo[slot >> 3] = make_float4(lo0, lo1, lo2, __int_as_float(fi));
        sOBhi[slot >> 3] = make_float4(hi0, hi1, hi2, 0.f);
      }
    }
    __syncthreads();

    /* ---------------- cull + test (per warp) ---------------- */
    if (warp_has_targets) {
      int cur = -1;
      float tpx = 3.0e30f, tpy = 0.f, tpz = 0.f, thr_lo = -3.4e38f, thr_hi = 3.4e38f, r2e = 0.f;
      bool skip = true;
      for (int ob = 0; ob < noct; ob += 32) {
        const int o = ob + lane;
        bool acc = false;
        if (o < noct) {
          const float4 lo = sOBlo[o], hi = sOBhi[o];
          const ItemInfoS &ii = sII[__float_as_int(lo.w)];
          const float r = fmaf(rmax, PREFILTER_REL, ii.margin);
          float d2 = 0.f;
          {
            const float a = lo.x - (bhi[0] + ii.d[0]), b = (blo[0] + ii.d[0]) - hi.x;
            const float gx = fmaxf(0.f, fmaxf(a, b));
            d2 = gx * gx;
          }
          {
            const float a = lo.y - (bhi[1] + ii.d[1]), b = (blo[1] + ii.d[1]) - hi.y;
            const float gy = fmaxf(0.f, fmaxf(a, b));
            d2 = fmaf(gy, gy, d2);
          }
          {
            const float a = lo.z - (bhi[2] + ii.d[2]), b = (blo[2] + ii.d[2]) - hi.z;
            const float gz = fmaxf(0.f, fmaxf(a, b));
            d2 = fmaf(gz, gz, d2);
          }
          acc = d2 < r * r;
        }
        unsigned m = __ballot_sync(FULL_MASK, acc);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1u;
          const int o2 = ob + b;
          const int item = __float_as_int(sOBlo[o2].w);
          if (item != cur) {
            cur = item;
            const ItemInfoS &ii = sII[item];
            const float4 tp = sTP[item * CTA_TARGETS + warp * 8 + t8];
            tpx = tp.x;
            tpy = tp.y;
            tpz = tp.z;
            thr_lo = ii.kc == 2 ? tp.w : -3.4e38f;
            thr_hi = ii.kc == 1 ? tp.w : 3.4e38f;
            if (ii.dbl) {
              const float re = fmaf(thg, PREFILTER_REL, ii.margin);
              r2e = re * re;
            } else {
              r2e = __fmul_rn(thg2, PREFILTER_REL);
            }
            skip = !__any_sync(FULL_MASK, tpx < 1.0e30f);
          }
          if (skip) continue;
          if (__any_sync(FULL_MASK, nsub > SUBCAP - 2)) drain();
          const int sl = o2 * 8 + 2 * s4;
          const float4 a = sF[sl], c = sF[sl + 1];
          ntests += 2;
          {
            const float dx = tpx - a.x, dy = tpy - a.y, dz = tpz - a.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (r2 < r2e && a.w < thr_hi && a.w > thr_lo) {
              mylist[nsub * CTA_THREADS] = (uint16_t)sl;
              nsub++;
            }
          }
          {
            const float dx = tpx - c.x, dy = tpy - c.y, dz = tpz - c.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (r2 < r2e && c.w < thr_hi && c.w > thr_lo) {
              mylist[nsub * CTA_THREADS] = (uint16_t)(sl + 1);
              nsub++;
            }
          }
        }
      }
      drain();
    }
    __syncthreads();
  }

  /* ---- combine the 4 partial sums of each target and flush ---- */
  if (LOOP == LOOP_DENSITY) {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      dacc.rho += __shfl_xor_sync(FULL_MASK, dacc.rho, o);
      dacc.rho_dh += __shfl_xor_sync(FULL_MASK, dacc.rho_dh, o);
      dacc.wcount += __shfl_xor_sync(FULL_MASK, dacc.wcount, o);
      dacc.wcount_dh += __shfl_xor_sync(FULL_MASK, dacc.wcount_dh, o);
      dacc.div_v += __shfl_xor_sync(FULL_MASK, dacc.div_v, o);
      dacc.rot[0] += __shfl_xor_sync(FULL_MASK, dacc.rot[0], o);
      dacc.rot[1] += __shfl_xor_sync(FULL_MASK, dacc.rot[1], o);
      dacc.rot[2] += __shfl_xor_sync(FULL_MASK, dacc.rot[2], o);
    }
  } else {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      gacc.v_sig = fmaxf(gacc.v_sig, __shfl_xor_sync(FULL_MASK, gacc.v_sig, o));
      gacc.laplace_u += __shfl_xor_sync(FULL_MASK, gacc.laplace_u, o);
      gacc.alpha_max = fmaxf(gacc.alpha_max, __shfl_xor_sync(FULL_MASK, gacc.alpha_max, o));
    }
  }
  int nh = nhit;
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) nh += __shfl_xor_sync(FULL_MASK, nh, o);
  if (tvalid && s4 == 0) {
    if (LOOP == LOOP_DENSITY) {
      float *pa = (float *)&A.dA[ti];
      float *pb = (float *)&A.dB[ti];
      atomicAdd(pa + 0, dacc.rho);
      atomicAdd(pa + 1, dacc.rho_dh);
      atomicAdd(pa + 2, dacc.wcount);
      atomicAdd(pa + 3, dacc.wcount_dh);
      atomicAdd(pb + 0, dacc.div_v);
      atomicAdd(pb + 1, dacc.rot[0]);
      atomicAdd(pb + 2, dacc.rot[1]);
      atomicAdd(pb + 3, dacc.rot[2]);
    } else {
      atomic_max_pos(&A.g_vsig[ti], gacc.v_sig);
      atomicAdd(&A.g_lap[ti], gacc.laplace_u);
      atomic_max_pos(&A.g_amax[ti], gacc.alpha_max);
    }
    if (nh) atomicAdd(&A.count[ti], nh);
  }
  int tot = nhit, tt = ntests;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(FULL_MASK, tot, o);
    tt += __shfl_xor_sync(FULL_MASK, tt, o);
  }
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && tt) atomicAdd(A.tests, (unsigned long long)tt);
}

}  // namespace swiftgpu
#endif
