/*
 * loops_cta.cuh - CTA-cooperative neighbour loops ("cell-pair tiles").
 *
 * One CTA of 8 warps owns up to 64 TARGET particles of one target cell.
 *
 *   PLAN      once per CTA: the constants of every item of the group
 *             (frames, sorted-axis limits; one thread per item), an item-level
 *             cull of source cells that are out of reach of the CTA's target
 *             box, and the split of the surviving items into batches that fit
 *             the shared-memory pool.
 *   STAGE     per batch, all threads: the source particles go ONCE into the
 *             pool (exact frame floats, payload, global index) together with
 *             an axis-aligned box per OCTET of 8 consecutive sources, and the
 *             per-(target, item) prefilter positions / key thresholds (TP).
 *   CULL      per warp (8 targets, lane = (t = lane & 7, s = lane >> 3)):
 *             lane-parallel box-box distance test of the warp's target box
 *             against every staged octet box (ballot -> accepted octets).
 *   TEST      per accepted octet every lane tests its target against sources
 *             2s and 2s+1 (FMA r2, inflated radius, sorted-axis key condition)
 *             and appends candidates to its private sub-list.
 *   INTERACT  the 4 sub-lists of a target are merged on the fly and drained by
 *             its 4 lanes; each candidate is re-evaluated with the reference's
 *             exact arithmetic (loops.cuh) before the interaction. The 4
 *             partial sums of a target are combined by two shuffles at the end.
 *
 * LOOP_DENSITY / LOOP_GRADIENT (+SUBSET = ghost re-runs) are type-1 loops
 * (r2 < h_t^2 gamma^2); LOOP_FORCE is the type-2 loop (DOSELF2 / DOPAIR2).
 */
#ifndef SWIFTGPU_LOOPS_CTA_CUH
#define SWIFTGPU_LOOPS_CTA_CUH

#include "loops_warp.cuh"

namespace swiftgpu {

#define CTA_THREADS 256
#define CTA_TARGETS 64
#define POOL 1024        /* staged sources per batch */
#define NOCT (POOL / 8)  /* octets per batch */
#define GROUP_ITEMS 32   /* items planned at once */
#define MAX_FRAGS 48     /* item fragments of one plan */
#define BATCH_FRAGS 16   /* fragments per batch (TP table rows) */
#define SUBCAP 16        /* sub-list capacity per lane */

/* Constants of one item of the group. */
struct __align__(16) ItemInfoS {
  double ot[3]; /* drain: subtracted from the target double */
  double fs[3]; /* frame origin of the staged source floats */
  double rshift, lim_a, lim_b; /* type-1: rshift, hi_max|hj_max, dj_min|di_max */
  double hi_max_g, hj_max_g, dx_max_d; /* force: DOPAIR2 constants (lim_a = di_max_sh, lim_b = dj_min) */
  float d[3];   /* cull: item-frame position of the target-cell origin */
  float dx_max; /* pair: ci.dx_max_sort + cj.dx_max_sort; subset pair: cj.dx_max_sort */
  float margin; /* cull / prefilter widening */
  float rsrc;   /* force: h_max * gamma of the source cell */
  int32_t mode, sid, kc, dbl;
  int32_t min_depth, max_depth;
  int32_t src_first, src_n;
};

struct Frag {
  int32_t item;     /* index into sII */
  int32_t src_off;  /* first source of the fragment inside the item */
  int32_t n;        /* sources */
  int32_t pool_off; /* first pool slot */
};

template <int NP, bool STAGE_D, bool KEYS>
struct CtaSmem {
  static constexpr int kF = 0;
  static constexpr int kP = kF + POOL * 16;
  static constexpr int kGI = kP + NP * POOL * 16;
  static constexpr int kK = kGI + POOL * 4;
  static constexpr int kD = kK + (KEYS ? POOL * 4 : 0);
  static constexpr int kOB = kD + (STAGE_D ? POOL * 24 : 0);
  static constexpr int kO2F = kOB + NOCT * 32;
  static constexpr int kTP = kO2F + NOCT;
  static constexpr int kII = kTP + BATCH_FRAGS * CTA_TARGETS * 16;
  static constexpr int kFr = kII + GROUP_ITEMS * (int)sizeof(ItemInfoS);
  static constexpr int kTS = kFr + MAX_FRAGS * (int)sizeof(Frag);
  static constexpr int kList = kTS + CTA_TARGETS * 48;
  static constexpr int kCtl = kList + SUBCAP * CTA_THREADS * 2;
  static constexpr int kBytes = kCtl + 512;
};

/* Target data every thread may need while filling the TP table. */
struct __align__(16) TgtS {
  double x, y, z;
  float thg, thg2;
  int32_t depth, valid;
  int32_t pad0_, pad1_;
};
static_assert(sizeof(TgtS) == 48, "TgtS");

template <int LOOP, bool SUBSET, int SCHEME>
__global__ void __launch_bounds__(CTA_THREADS, (LOOP == LOOP_FORCE ? 2 : 3)) k_cta(const LoopArgs A) {
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  constexpr bool STAGE_D = false; /* doubles of dbl-mode candidates come from global memory (L1/L2) */
  typedef CtaSmem<NP, STAGE_D, FORCE> SM;
  extern __shared__ __align__(16) char smem[];
  float4 *const sF = (float4 *)(smem + SM::kF);
  float4 *const sP0 = (float4 *)(smem + SM::kP);
  float4 *const sP1 = sP0 + POOL;
  float4 *const sP2 = sP0 + 2 * POOL;
  float4 *const sP3 = sP0 + 3 * POOL;
  int32_t *const sGI = (int32_t *)(smem + SM::kGI);
  float *const sK = (float *)(smem + SM::kK);
  double *const sD = (double *)(smem + SM::kD);
  float4 *const sOBlo = (float4 *)(smem + SM::kOB);
  float4 *const sOBhi = sOBlo + NOCT;
  uint8_t *const sO2F = (uint8_t *)(smem + SM::kO2F);
  float4 *const sTP = (float4 *)(smem + SM::kTP);
  ItemInfoS *const sII = (ItemInfoS *)(smem + SM::kII);
  Frag *const sFr = (Frag *)(smem + SM::kFr);
  TgtS *const sTS = (TgtS *)(smem + SM::kTS);
  uint16_t *const sList = (uint16_t *)(smem + SM::kList);
  int32_t *const sCtl = (int32_t *)(smem + SM::kCtl);
  /* sCtl: [0] nfrags, [1] nbatches, [2..2+MAX_FRAGS] first fragment of batch b, then the warps' boxes */
  float *const sBox = (float *)(sCtl + 64);

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
  ForceQ tq;
  tq.m = tq.vx = tq.vy = tq.vz = 0.f;
  tq.rho = 1.f;
  tq.P = tq.f = tq.cs = tq.balsara = 0.f;
  tq.h = 1.f;
  tq.u = tq.alpha_visc = tq.alpha_diff = 0.f;
  tq.time_bin = 0;
  int tdepth = 0;
  if (tvalid) {
    tx = A.x[3 * (size_t)ti];
    ty = A.x[3 * (size_t)ti + 1];
    tz = A.x[3 * (size_t)ti + 2];
    const float4 q = A.mv[ti];
    tvx = q.y;
    tvy = q.z;
    tvz = q.w;
    tdepth = A.depth_h[ti];
    if (FORCE) {
      const float4 q1 = A.fq1[ti], q2 = A.fq2[ti];
      tq.m = q.x; tq.vx = q.y; tq.vy = q.z; tq.vz = q.w;
      tq.rho = q1.x; tq.P = q1.y; tq.f = q1.z; tq.cs = q1.w;
      tq.balsara = q2.x; tq.h = q2.y; tq.u = q2.z; tq.time_bin = __float_as_int(q2.w);
      if (SCHEME == SCH_SPHENIX) {
        const float4 q3 = A.fq3[ti];
        tq.alpha_visc = q3.x;
        tq.alpha_diff = q3.y;
      }
      th = tq.h;
    } else {
      th = A.h[ti];
      if (LOOP == LOOP_GRADIENT) {
        tu = A.fq2[ti].z;
        tcs = A.fq1[ti].w;
      }
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
    if (lane == 0) {
      float *b = sBox + warp * 8;
      b[0] = blo[0]; b[1] = blo[1]; b[2] = blo[2];
      b[3] = bhi[0]; b[4] = bhi[1]; b[5] = bhi[2];
      b[6] = rmax;
    }
  }
  const bool warp_has_targets = __any_sync(FULL_MASK, tvalid);
  __syncthreads();
  /* the CTA's target box */
  float clo[3], chi[3], crmax = 0.f;
  {
    clo[0] = clo[1] = clo[2] = 3.0e30f;
    chi[0] = chi[1] = chi[2] = -3.0e30f;
#pragma unroll
    for (int w = 0; w < CTA_THREADS / 32; w++) {
      const float *b = sBox + w * 8;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        clo[k] = fminf(clo[k], b[k]);
        chi[k] = fmaxf(chi[k], b[3 + k]);
      }
      crmax = fmaxf(crmax, b[6]);
    }
  }

  DensityAcc dacc;
  dacc.zero();
  GradientAcc gacc;
  gacc.v_sig = 0.f;
  gacc.laplace_u = 0.f;
  gacc.alpha_max = 0.f;
  ForceAcc facc;
  facc.ax = facc.ay = facc.az = facc.u_dt = facc.h_dt = 0.f;
  facc.v_sig = 0.f;
  facc.min_ngb = NUM_TIME_BINS + 1;
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
      const int fl = sO2F[slot >> 3];
      const ItemInfoS &ii = sII[__float_as_int(sOBlo[slot >> 3].w)];
      const int gi = sGI[slot];
      const float4 s = sF[slot];
      const bool dbl = act && (ii.dbl != 0);
      float dx, dy, dz;
      {
        /* float modes: the reference's dx is tp - sp on the very floats of the test */
        const float4 tp = sTP[fl * CTA_TARGETS + warp * 8 + t8];
        dx = __fsub_rn(tp.x, s.x);
        dy = __fsub_rn(tp.y, s.y);
        dz = __fsub_rn(tp.z, s.z);
      }
      if (__any_sync(FULL_MASK, dbl)) {
        if (dbl) {
          double sxd, syd, szd;
          if (STAGE_D) {
            sxd = sD[slot];
            syd = sD[POOL + slot];
            szd = sD[2 * POOL + slot];
          } else {
            sxd = A.x[3 * (size_t)gi];
            syd = A.x[3 * (size_t)gi + 1];
            szd = A.x[3 * (size_t)gi + 2];
          }
          dx = dsubf(__dsub_rn(tx, ii.ot[0]), sxd);
          dy = dsubf(__dsub_rn(ty, ii.ot[1]), syd);
          dz = dsubf(__dsub_rn(tz, ii.ot[2]), szd);
        }
      }
      const float r2 = r2_exact(dx, dy, dz);
      if (!FORCE) {
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
      } else {
        const float4 q2 = sP2[slot];
        const float sh = q2.y;
        const float shg2 = hg2_exact(sh);
        bool ok;
        if (ii.dbl) {
          /* DOSELF2 :2792: doi = r2 < hig2 || r2 < hjg2 */
          ok = (r2 < thg2 || r2 < shg2) && (gi != ti);
        } else {
          /* DOPAIR2 pass conditions (see loops.cuh k_loop2) */
          const bool tleft = (ii.mode == MODE_PAIR_L);
          const int sid = ii.sid;
          const double rshift = ii.rshift, hi_max_g = ii.hi_max_g, hj_max_g = ii.hj_max_g,
                       dx_max = ii.dx_max_d, di_max_sh = ii.lim_a, dj_min = ii.lim_b;
          const float tkey = sort_key(tx, ty, tz, sid);
          const float skey = sK[slot];
          const float shg = __fmul_rn(sh, KERNEL_GAMMA);
          if (tleft) {
            const bool inA =
                __dsub_rn(__dadd_rn(__dadd_rn((double)tkey, hi_max_g), dx_max), rshift) > dj_min;
            const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(tkey, thg), dx_max), rshift);
            const double t_di = (inA && !(di < dj_min)) ? di : -1.0e300;
            const double t_keysh = __dsub_rn((double)tkey, rshift);
            const bool inB = __dsub_rn(__dsub_rn((double)skey, hj_max_g), dx_max) < di_max_sh;
            const double dj = __dsub_rn((double)__fsub_rn(skey, shg), dx_max);
            const double s_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
            const bool c1 = ((double)skey < t_di) && (r2 < thg2);
            const bool c2 = (t_keysh > s_dj) && (r2 < shg2) && !(r2 < thg2);
            ok = c1 || c2;
          } else {
            const bool inB = __dsub_rn(__dsub_rn((double)tkey, hj_max_g), dx_max) < di_max_sh;
            const double dj = __dsub_rn((double)__fsub_rn(tkey, thg), dx_max);
            const double t_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
            const bool inA =
                __dsub_rn(__dadd_rn(__dadd_rn((double)skey, hi_max_g), dx_max), rshift) > dj_min;
            const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(skey, shg), dx_max), rshift);
            const double s_di = (inA && !(di < dj_min)) ? di : -1.0e300;
            const double s_keysh = __dsub_rn((double)skey, rshift);
            const bool c1 = ((double)tkey < s_di) && (r2 < shg2);
            const bool c2 = (s_keysh > t_dj) && (r2 < thg2) && !(r2 < shg2);
            ok = c1 || c2;
          }
        }
        if (act && ok) {
          ForceQ sq;
          const float4 q0 = sP0[slot], q1 = sP1[slot];
          sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
          sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
          sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
          sq.alpha_visc = sq.alpha_diff = 0.f;
          if (SCHEME == SCH_SPHENIX) {
            const float4 q3 = sP3[slot];
            sq.alpha_visc = q3.x;
            sq.alpha_diff = q3.y;
          }
          iact_force<SCHEME>(facc, r2, dx, dy, dz, tq, sq, A.a2_Hubble);
          nhit++;
        }
      }
    }
    nsub = 0;
    __syncwarp();
  };

  if (tid == 0) {
    sCtl[60] = 0; /* next item of the group to plan */
    sCtl[61] = 0; /* ... starting at this source offset */
  }
  for (;;) {
    __syncthreads(); /* previous plan fully consumed */
    const int sb = sCtl[60];
    const int offb = sCtl[61];
    if (sb >= G.item_count) break;
    const int nitems = min(GROUP_ITEMS, G.item_count - sb);
    __syncthreads();

    /* ================= PLAN (warp 0: lane i = item sb + i) ================= */
    if (warp == 0) {
      bool keep = false;
      ItemInfoS ii;
      int scount = 0;
      if (lane < nitems) {
        const Item I = A.items[G.item_first + sb + lane];
        const DevCell sc = A.cells[I.scell];
        const int mode = I.mode, sid = I.sid;
        scount = sc.count;
        const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
        ii.mode = mode;
        ii.sid = sid;
        ii.min_depth = I.min_depth;
        ii.max_depth = I.max_depth;
        ii.src_first = sc.first;
        ii.src_n = sc.count;
        ii.rshift = ii.lim_a = ii.lim_b = ii.hi_max_g = ii.hj_max_g = ii.dx_max_d = 0.;
        ii.dx_max = 0.f;
        ii.kc = 0;
        ii.dbl = 0;
        ii.margin = 2.0e-6f * sc.width + 2.0e-6f * tcell.width;
        ii.rsrc = FORCE ? __fmul_rn(sc.h_max, KERNEL_GAMMA) : 0.f;
        double otx = 0., oty = 0., otz = 0., fsx, fsy, fsz;
        if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
          const DevCell &ci = (mode == MODE_PAIR_L) ? tcell : sc;
          const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tcell;
          ii.rshift = __dadd_rn(
              __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
              __dmul_rn(shz, c_runner_shift[sid][2]));
          const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                       oiz = __dadd_rn(cj.loc[2], shz);
          /* sort_j[0].d and sort_i[count-1].d of the reference = key extrema */
          const double dj_min = (double)A.ext[seg_index(cj, sid)].x;
          const double di_max = (double)A.ext[seg_index(ci, sid)].y;
          if (FORCE) {
            ii.hi_max_g = __dmul_rn((double)ci.h_max, (double)KERNEL_GAMMA);
            ii.hj_max_g = __dmul_rn((double)cj.h_max, (double)KERNEL_GAMMA);
            ii.dx_max_d = (double)__fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
            ii.lim_a = __dsub_rn(di_max, ii.rshift); /* di_max_sh */
            ii.lim_b = dj_min;
          } else {
            const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
            ii.dx_max = __fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
            if (mode == MODE_PAIR_L) {
              ii.lim_a = __dsub_rn((double)__fmul_rn(fminf(h_max_lim, ci.h_max_active), KERNEL_GAMMA), ii.rshift);
              ii.lim_b = dj_min;
              ii.kc = 1;
            } else {
              ii.lim_a = (double)__fmul_rn(fminf(h_max_lim, cj.h_max_active), KERNEL_GAMMA);
              ii.lim_b = __dsub_rn(di_max, ii.rshift);
              ii.kc = 2;
            }
          }
          if (mode == MODE_PAIR_L) {
            otx = oix; oty = oiy; otz = oiz;
            fsx = cj.loc[0]; fsy = cj.loc[1]; fsz = cj.loc[2];
          } else {
            otx = cj.loc[0]; oty = cj.loc[1]; otz = cj.loc[2];
            fsx = oix; fsy = oiy; fsz = oiz;
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
        /* item-level cull: the source cell's box (staged-float frame) against
         * the CTA's target box */
        {
          const float r = fmaf(fmaxf(crmax, ii.rsrc), PREFILTER_REL, ii.margin) + sc.dx_max_part;
          const float c0[3] = {dsubf(sc.loc[0], fsx), dsubf(sc.loc[1], fsy), dsubf(sc.loc[2], fsz)};
          float d2 = 0.f;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float a = c0[k] - (chi[k] + ii.d[k]), b = (clo[k] + ii.d[k]) - (c0[k] + sc.width);
            const float gk = fmaxf(0.f, fmaxf(a, b));
            d2 = fmaf(gk, gk, d2);
          }
          keep = (d2 < r * r) && scount > 0;
        }
        sII[lane] = ii;
      }
      __syncwarp();
      /* surviving items -> fragments -> batches (lane 0, shared memory only) */
      const unsigned km = __ballot_sync(FULL_MASK, keep);
      int nfr = 0, nbat = 0;
      if (lane == 0) {
        int used = POOL + 1, inb = BATCH_FRAGS; /* forces a new batch first */
        unsigned m = km;
        int next_item = sb + nitems, next_off = 0;
        while (m) {
          const int it = __ffs(m) - 1;
          m &= m - 1u;
          int off = (it == 0) ? offb : 0;
          int left = sII[it].src_n - off;
          if (nfr >= MAX_FRAGS) { /* out of fragments: resume here in the next plan */
            next_item = sb + it;
            next_off = off;
            break;
          }
          while (left > 0) {
            if (nfr >= MAX_FRAGS) {
              next_item = sb + it;
              next_off = off;
              m = 0;
              break;
            }
            const int take = min(left, POOL);
            const int padded = (take + 7) & ~7;
            if (used + padded > POOL || inb >= BATCH_FRAGS) {
              sCtl[2 + nbat] = nfr;
              nbat++;
              used = 0;
              inb = 0;
            }
            Frag f;
            f.item = it;
            f.src_off = off;
            f.n = take;
            f.pool_off = used;
            sFr[nfr++] = f;
            used += padded;
            inb++;
            left -= take;
            off += take;
          }
        }
        sCtl[2 + nbat] = nfr;
        sCtl[0] = nfr;
        sCtl[1] = nbat;
        sCtl[60] = next_item;
        sCtl[61] = next_off;
      }
    }
    __syncthreads();
    const int nbat = sCtl[1];

    for (int bt = 0; bt < nbat; bt++) {
      const int f0 = sCtl[2 + bt], f1 = sCtl[3 + bt];
      const int nb = f1 - f0;
      const int noct = (sFr[f1 - 1].pool_off + ((sFr[f1 - 1].n + 7) & ~7)) >> 3;

      /* ---------------- octet -> fragment map ---------------- */
      if (tid < nb) {
        const Frag f = sFr[f0 + tid];
        const int o0 = f.pool_off >> 3, o1 = (f.pool_off + f.n + 7) >> 3;
        for (int o = o0; o < o1; o++) sO2F[o] = (uint8_t)tid;
      }
      /* ---------------- TP table: per (target, fragment) prefilter position + key threshold ---------------- */
      for (int e = tid; e < nb * CTA_TARGETS; e += CTA_THREADS) {
        const int tt = e & (CTA_TARGETS - 1), fl = e >> 6;
        const ItemInfoS &ii = sII[sFr[f0 + fl].item];
        const TgtS ts = sTS[tt];
        bool part = ts.valid && ts.depth >= ii.min_depth && ts.depth <= ii.max_depth;
        float thr = 0.f, px, py, pz;
        const int mode = ii.mode, sid = ii.sid;
        if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
          if (!FORCE) {
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
            const float f0_ = (mode == MODE_SUB_PAIR) ? __fadd_rn(ts.thg, ii.dx_max) : __fsub_rn(-ts.thg, ii.dx_max);
            const double di =
                __dadd_rn(__dadd_rn(__dadd_rn((double)f0_, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                    __dmul_rn(tdy, c_runner_shift[sid][1])),
                          __dmul_rn(tdz, c_runner_shift[sid][2]));
            thr = (mode == MODE_SUB_PAIR) ? __double2float_ru(di) : __double2float_rd(di);
          }
        }
        if (!part) px = 3.0e30f;
        sTP[fl * CTA_TARGETS + tt] = make_float4(px, py, pz, thr);
      }
      __syncthreads(); /* sO2F visible */

      /* ---------------- stage the sources of the batch ---------------- */
      for (int slot = tid; slot < ((noct * 8 + 31) & ~31); slot += CTA_THREADS) {
        const bool inb = slot < noct * 8;
        const int fl = inb ? sO2F[slot >> 3] : 0;
        const Frag f = sFr[f0 + fl];
        const ItemInfoS &ii = sII[f.item];
        const int k = slot - f.pool_off;
        const bool valid = inb && k < f.n;
        float fx = 0.f, fy = 0.f, fz = 0.f, key = 0.f, w = 0.f;
        int sj = -1;
        if (valid) {
          sj = ii.src_first + f.src_off + k;
          const double sx = A.x[3 * (size_t)sj], sy = A.x[3 * (size_t)sj + 1], sz = A.x[3 * (size_t)sj + 2];
          fx = dsubf(sx, ii.fs[0]);
          fy = dsubf(sy, ii.fs[1]);
          fz = dsubf(sz, ii.fs[2]);
          if (FORCE) {
            const float4 q2 = A.fq2[sj];
            if (!ii.dbl) {
              key = sort_key(sx, sy, sz, ii.sid);
              w = __fmul_rn(hg2_exact(q2.y), PREFILTER_REL);
            } else {
              const float re = fmaf(__fmul_rn(q2.y, KERNEL_GAMMA), PREFILTER_REL, ii.margin);
              w = re * re;
            }
            sK[slot] = key;
            sP0[slot] = A.mv[sj];
            sP1[slot] = A.fq1[sj];
            sP2[slot] = q2;
            if (SCHEME == SCH_SPHENIX) sP3[slot] = A.fq3[sj];
          } else {
            /* w: the sorted-axis condition folded into ONE compare "w < thr":
             * kc 1 (key < thr): w = key; kc 2 (key > thr): w = -key, thr = -thr; none: w = -inf */
            if (ii.kc) key = sort_key(sx, sy, sz, ii.sid);
            w = ii.kc == 1 ? key : (ii.kc == 2 ? -key : -3.4e38f);
            sP0[slot] = A.mv[sj];
            if (LOOP == LOOP_GRADIENT) {
              const float4 q1 = A.fq1[sj];
              const float4 q2 = A.fq2[sj];
              const float4 q3 = A.fq3[sj];
              sP1[slot] = make_float4(q2.z /*u*/, q1.x /*rho*/, q1.w /*cs*/, q3.x /*alpha*/);
            }
          }
          if (STAGE_D) {
            sD[slot] = sx;
            sD[POOL + slot] = sy;
            sD[2 * POOL + slot] = sz;
          }
        } else if (FORCE && inb) {
          sP2[slot] = make_float4(0.f, 1.f, 0.f, 0.f);
          sK[slot] = 0.f;
        }
        if (inb) {
          sF[slot] = valid ? make_float4(fx, fy, fz, w) : make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
          sGI[slot] = sj;
        }
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
        if (inb && (slot & 7) == 0) {
          sOBlo[slot >> 3] = make_float4(lo0, lo1, lo2, __int_as_float(f.item));
          sOBhi[slot >> 3] = make_float4(hi0, hi1, hi2, 0.f);
        }
      }
      __syncthreads();

      /* ---------------- cull + test (per warp) ---------------- */
      if (warp_has_targets) {
        int cur = -1;
        float tpx = 3.0e30f, tpy = 0.f, tpz = 0.f, thr_hi = 3.4e38f, r2e = 0.f;
        int nsub_ub = nsub; /* warp-uniform upper bound of every lane's nsub */
        bool skip = true;
        for (int ob = 0; ob < noct; ob += 32) {
          const int o = ob + lane;
          bool acc = false;
          if (o < noct) {
            const float4 lo = sOBlo[o], hi = sOBhi[o];
            const ItemInfoS &ii = sII[__float_as_int(lo.w)];
            const float r = fmaf(FORCE ? fmaxf(rmax, ii.rsrc) : rmax, PREFILTER_REL, ii.margin);
            float d2;
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
            const int fl = sO2F[o2];
            if (fl != cur) {
              cur = fl;
              const ItemInfoS &ii = sII[sFr[f0 + fl].item];
              const float4 tp = sTP[fl * CTA_TARGETS + warp * 8 + t8];
              tpx = tp.x;
              tpy = tp.y;
              tpz = tp.z;
              thr_hi = ii.kc == 1 ? tp.w : (ii.kc == 2 ? -tp.w : 3.4e38f);
              if (ii.dbl) {
                const float re = fmaf(thg, PREFILTER_REL, ii.margin);
                r2e = re * re;
              } else {
                r2e = __fmul_rn(thg2, PREFILTER_REL);
              }
              skip = !__any_sync(FULL_MASK, tpx < 1.0e30f);
            }
            if (skip) continue;
            if (nsub_ub > SUBCAP - 2) {
              nsub_ub = __reduce_max_sync(FULL_MASK, nsub);
              if (nsub_ub > SUBCAP - 2) {
                drain();
                nsub_ub = 0;
              }
            }
            nsub_ub += 2;
            const int sl = o2 * 8 + 2 * s4;
            const float4 a = sF[sl], c = sF[sl + 1];
            ntests++;
            {
              const float dx = tpx - a.x, dy = tpy - a.y, dz = tpz - a.z;
              const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
              const bool ok = FORCE ? (r2 < fmaxf(r2e, a.w)) : (r2 < r2e && a.w < thr_hi);
              if (ok) {
                mylist[nsub * CTA_THREADS] = (uint16_t)sl;
                nsub++;
              }
            }
            {
              const float dx = tpx - c.x, dy = tpy - c.y, dz = tpz - c.z;
              const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
              const bool ok = FORCE ? (r2 < fmaxf(r2e, c.w)) : (r2 < r2e && c.w < thr_hi);
              if (ok) {
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
  }

  /* ---- combine the 4 partial sums of each target and flush ---- */
  int nh = nhit;
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) nh += __shfl_xor_sync(FULL_MASK, nh, o);
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
    if (tvalid && s4 == 0) {
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
    }
  } else if (LOOP == LOOP_GRADIENT) {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      gacc.v_sig = fmaxf(gacc.v_sig, __shfl_xor_sync(FULL_MASK, gacc.v_sig, o));
      gacc.laplace_u += __shfl_xor_sync(FULL_MASK, gacc.laplace_u, o);
      gacc.alpha_max = fmaxf(gacc.alpha_max, __shfl_xor_sync(FULL_MASK, gacc.alpha_max, o));
    }
    if (tvalid && s4 == 0) {
      atomic_max_pos(&A.g_vsig[ti], gacc.v_sig);
      atomicAdd(&A.g_lap[ti], gacc.laplace_u);
      atomic_max_pos(&A.g_amax[ti], gacc.alpha_max);
    }
  } else {
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
      facc.ax += __shfl_xor_sync(FULL_MASK, facc.ax, o);
      facc.ay += __shfl_xor_sync(FULL_MASK, facc.ay, o);
      facc.az += __shfl_xor_sync(FULL_MASK, facc.az, o);
      facc.u_dt += __shfl_xor_sync(FULL_MASK, facc.u_dt, o);
      facc.h_dt += __shfl_xor_sync(FULL_MASK, facc.h_dt, o);
      facc.v_sig = fmaxf(facc.v_sig, __shfl_xor_sync(FULL_MASK, facc.v_sig, o));
      facc.min_ngb = min(facc.min_ngb, __shfl_xor_sync(FULL_MASK, facc.min_ngb, o));
    }
    if (tvalid && s4 == 0) {
      float *po = (float *)&A.fo1[ti];
      atomicAdd(po + 0, facc.ax);
      atomicAdd(po + 1, facc.ay);
      atomicAdd(po + 2, facc.az);
      atomicAdd(po + 3, facc.u_dt);
      atomicAdd(&A.f_hdt[ti], facc.h_dt);
      if (SCHEME != SCH_SPHENIX) atomic_max_pos(&A.f_vsig[ti], facc.v_sig);
      atomicMin(&A.f_minngb[ti], facc.min_ngb);
    }
  }
  if (tvalid && s4 == 0 && nh) atomicAdd(&A.count[ti], nh);
  int tot = nhit, tt = ntests;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(FULL_MASK, tot, o);
    tt += __shfl_xor_sync(FULL_MASK, tt, o);
  }
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && tt) atomicAdd(A.tests, 2ull * (unsigned long long)tt);
}

}  // namespace swiftgpu
#endif
