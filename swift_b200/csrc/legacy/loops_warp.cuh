/*
 * legacy/loops_warp.cuh - FIRST generation of the neighbour-loop kernels (kept for reference; built only
 * with -DSWIFTGPU_LEGACY_LOOPS, `make legacy`): the neighbour-loop kernels (K1/K2/K3 of SURVEY 2.1).
 *
 * One warp owns up to 32 TARGET particles of one target cell (a "task") and
 * walks every directed item of that cell's group (worklist.hpp). The work is
 * split in two phases so that both run at full SIMT efficiency:
 *
 *  TEST   The source cells of the items are streamed, in sorted-axis order, in
 *         chunks of 32 particles into a warp-private shared-memory tile that
 *         SPANS ITEMS (up to SCAP slots). Every lane tests its own target
 *         against the 32 freshly staged sources with a cheap conservative
 *         prefilter (FMA r2 against an inflated h^2 gamma^2, broadcast LDS,
 *         ~9 instructions per pair) and appends the slot numbers of the
 *         candidates to its private hit list in shared memory. Pseudo-Verlet
 *         pruning: a chunk whose first sort key is beyond the reach of every
 *         lane (warp max/min by shuffles) ends the item.
 *
 *  INTERACT  When the tile or a list is full (and at the end of the task) the
 *         lists are drained lane-parallel: lane t pops its k-th candidate,
 *         re-evaluates the reference's EXACT arithmetic for that pair (frame
 *         positions from doubles, un-fused r2, sorted-axis key conditions) and
 *         applies the non-symmetric interaction. Because a drain covers many
 *         items, lanes hold similar numbers of candidates (mean/max ~0.75
 *         instead of ~0.3 when draining per 32-source chunk).
 *
 * Accumulators stay in registers for the whole task and are flushed once (no
 * atomics in the inner loops).
 *
 * Reference semantics reproduced (runner_doiact_functions_hydro.h):
 *   MODE_SELF      DOSELF1 :2299 / DOSELF2 :2624   dx = (float)(x_t - x_s) on doubles
 *   MODE_PAIR_L/R  DOPAIR1 :1234 / DOPAIR2 :1601   floats in the frame cj->loc (+shift)
 *   MODE_SUB_SELF  DOSELF_SUBSET :1108             floats relative to c->loc
 *   MODE_SUB_PAIR  DOPAIR_SUBSET :855              (float)((x_t - shift) - x_s) on doubles
 * Both dx forms are evaluated by ONE branch-free expression in the drain:
 *   dx = (float)((x_t - ot) - xs_d) - xs_f
 * with (ot, xs_d, xs_f) = (frame origin, 0, staged frame float) for the float
 * modes and (shift or 0, source double, 0) for the double modes; subtracting
 * an exact zero does not round, so both reproduce the reference bit for bit.
 */
#ifndef SWIFTGPU_LOOPS_WARP_CUH
#define SWIFTGPU_LOOPS_WARP_CUH

#include "../loops_common.cuh"

namespace swiftgpu {

#define LCAP 64 /* hit-list capacity per target */
#define TPL 2   /* targets per lane: every staged source is tested against 2 targets */

/* What the drain needs to know about a staged chunk of 32 slots. */
struct __align__(16) ChunkInfo {
  double ot[3]; /* subtracted from the target's double position */
  int32_t dbl;  /* 1: dx from doubles (MODE_SELF, MODE_SUB_PAIR*) */
  int32_t item;
};

/* Type-2 (force) chunks also carry the constants of their pair item
 * (DOPAIR2 :1601-1735), so that the drain does not have to re-derive them. */
struct __align__(16) ChunkInfoF {
  double ot[3];
  double rshift, hi_max_g, hj_max_g, dx_max, di_max_sh, dj_min;
  int32_t dbl;   /* 1: MODE_SELF */
  int32_t sid;   /* bit 8: targets are in the left cell (MODE_PAIR_L) */
};

/* Shared-memory tile of one warp. NP = payload float4 per source. */
template <int SCAP, int NP, bool STAGE_D, bool KEYS, typename CIT>
struct Tile {
  static constexpr int kF = 0;
  static constexpr int kP = kF + SCAP * 16;
  static constexpr int kGI = kP + NP * SCAP * 16;
  static constexpr int kK = kGI + SCAP * 4;
  static constexpr int kD = kK + (KEYS ? SCAP * 4 : 0);
  static constexpr int kList = kD + (STAGE_D ? SCAP * 24 : 0);
  static constexpr int kChunk = kList + LCAP * TASK_TARGETS * 2;
  static constexpr int kBytes = kChunk + (SCAP / 32) * (int)sizeof(CIT);
  char *base;
  __device__ __forceinline__ float4 *F() const { return (float4 *)(base + kF); }
  __device__ __forceinline__ float4 *P(int k) const { return (float4 *)(base + kP) + k * SCAP; }
  __device__ __forceinline__ int32_t *GI() const { return (int32_t *)(base + kGI); }
  __device__ __forceinline__ float *K() const { return (float *)(base + kK); }
  __device__ __forceinline__ double *D() const { return (double *)(base + kD); }
  __device__ __forceinline__ uint16_t *list() const { return (uint16_t *)(base + kList); }
  __device__ __forceinline__ CIT *chunk() const { return (CIT *)(base + kChunk); }
};

/* The prefilter over one freshly staged chunk: TPL lane-private targets
 * against 32 broadcast sources. KC: 0 no key condition, 1 key < thr,
 * 2 key > thr, 3 type-2 (radius^2 = max(target, source)). */
template <int KC>
__device__ __forceinline__ void test_chunk(const float4 *__restrict__ F, int nst,
                                           const float (&tp)[TPL][3], const float (&r2e)[TPL],
                                           const float (&thr)[TPL], uint16_t *(&lp)[TPL],
                                           int (&nlist)[TPL]) {
  unsigned mask[TPL];
#pragma unroll
  for (int u = 0; u < TPL; u++) mask[u] = 0u;
#pragma unroll 16
  for (int q = 0; q < 32; q++) {
    const float4 s = F[nst + q];
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      const float dx = tp[u][0] - s.x, dy = tp[u][1] - s.y, dz = tp[u][2] - s.z;
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      bool ok;
      if (KC == 3)
        ok = r2 < fmaxf(r2e[u], s.w);
      else
        ok = r2 < r2e[u];
      if (KC == 1) ok = ok && (s.w < thr[u]);
      if (KC == 2) ok = ok && (s.w > thr[u]);
      if (ok) mask[u] |= 1u << q;
    }
  }
  /* decode the (sparse) masks into the hit lists */
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    unsigned m = mask[u];
    while (m) {
      const int q = __ffs(m) - 1;
      m &= m - 1u;
      *lp[u] = (uint16_t)(nst + q);
      lp[u] += TASK_TARGETS;
      nlist[u]++;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Type-1 loops: density (all schemes), gradient (SPHENIX), and (SUBSET) the  */
/* density re-runs of the ghost. Hit criterion r2 < h_t^2 gamma^2.            */
/* ------------------------------------------------------------------------ */
#define SCAP1 256
template <int LOOP, bool SUBSET>
struct Tile1 : Tile<SCAP1, (LOOP == LOOP_GRADIENT ? 2 : 1), SUBSET, false, ChunkInfo> {};

template <int LOOP, bool SUBSET>
__global__ void __launch_bounds__(32) k_loop1(const LoopArgs A) {
  extern __shared__ __align__(16) char smem_raw[];
  typedef Tile1<LOOP, SUBSET> TT;
  TT T;
  T.base = smem_raw;
  const int lane = threadIdx.x;
  const int task = blockIdx.x;

  const int g = A.task_group[task];
  const int chunk = A.task_chunk[task];
  const int nt = A.tgt_count[g];
  if (chunk * TASK_TARGETS >= nt) return;
  const Group G = A.groups[g];

  /* target state */
  bool tvalid[TPL];
  int ti[TPL], tdepth[TPL];
  double tx[TPL], ty[TPL], tz[TPL];
  float th[TPL], tvx[TPL], tvy[TPL], tvz[TPL], tu[TPL], tcs[TPL], thg2[TPL], th_inv[TPL], thg[TPL];
  DensityAcc dacc[TPL];
  GradientAcc gacc[TPL];
  int nhit[TPL];
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    const int slot_t = chunk * TASK_TARGETS + u * 32 + lane;
    tvalid[u] = slot_t < nt;
    ti[u] = tvalid[u] ? A.tgt_list[A.tgt_first[g] + slot_t] : -1;
    tx[u] = ty[u] = tz[u] = 0.;
    th[u] = 1.f;
    tvx[u] = tvy[u] = tvz[u] = tu[u] = tcs[u] = 0.f;
    tdepth[u] = 0;
    if (tvalid[u]) {
      const size_t p = (size_t)ti[u];
      tx[u] = A.x[3 * p];
      ty[u] = A.x[3 * p + 1];
      tz[u] = A.x[3 * p + 2];
      th[u] = A.h[p];
      const float4 q = A.mv[p];
      tvx[u] = q.y;
      tvy[u] = q.z;
      tvz[u] = q.w;
      tdepth[u] = A.depth_h[p];
      if (LOOP == LOOP_GRADIENT) {
        tu[u] = A.fq2[p].z;
        tcs[u] = A.fq1[p].w;
      }
    }
    thg2[u] = hg2_exact(th[u]);
    th_inv[u] = 1.f / th[u];
    thg[u] = __fmul_rn(th[u], KERNEL_GAMMA); /* hi * kernel_gamma (float) */
    dacc[u].zero();
    gacc[u].v_sig = 0.f;
    gacc[u].laplace_u = 0.f;
    gacc[u].alpha_max = 0.f;
    nhit[u] = 0;
  }
  int nchunks = 0;

  /* tile state */
  int nst = 0; /* staged slots */
  int nlist[TPL];
  uint16_t *lp[TPL];
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    nlist[u] = 0;
    lp[u] = T.list() + u * 32 + lane;
  }

  /* ---- INTERACT: drain the hit lists ---- */
  auto drain = [&]() {
    __syncwarp();
    int mx = nlist[0];
#pragma unroll
    for (int u = 1; u < TPL; u++) mx = max(mx, nlist[u]);
    const int maxn = __reduce_max_sync(FULL_MASK, mx);
    const float4 *F = T.F();
    const int32_t *GI = T.GI();
    const ChunkInfo *CI = T.chunk();
    const uint16_t *l0 = T.list() + lane;
    for (int k = 0; k < maxn; k++) {
#pragma unroll
      for (int u = 0; u < TPL; u++) {
        const bool act = k < nlist[u];
        const int slot = act ? (int)l0[k * TASK_TARGETS + u * 32] : 0;
        const ChunkInfo ci = CI[slot >> 5];
        const int gi = GI[slot];
        const float4 s = F[slot];
        const bool dbl = ci.dbl != 0;
        double sxd = 0., syd = 0., szd = 0.;
        if (act && dbl) {
          if (SUBSET) {
            const double *D = T.D();
            sxd = D[slot];
            syd = D[SCAP1 + slot];
            szd = D[2 * SCAP1 + slot];
          } else {
            sxd = A.x[3 * (size_t)gi];
            syd = A.x[3 * (size_t)gi + 1];
            szd = A.x[3 * (size_t)gi + 2];
          }
        }
        const float spx = dbl ? 0.f : s.x, spy = dbl ? 0.f : s.y, spz = dbl ? 0.f : s.z;
        const float dx = __fsub_rn(dsubf(__dsub_rn(tx[u], ci.ot[0]), sxd), spx);
        const float dy = __fsub_rn(dsubf(__dsub_rn(ty[u], ci.ot[1]), syd), spy);
        const float dz = __fsub_rn(dsubf(__dsub_rn(tz[u], ci.ot[2]), szd), spz);
        const float r2 = r2_exact(dx, dy, dz);
        if (act && (r2 < thg2[u]) && (gi != ti[u])) {
          const float4 f0 = T.P(0)[slot];
          if (LOOP == LOOP_DENSITY) {
            iact_density(dacc[u], r2, dx, dy, dz, th_inv[u], tvx[u], tvy[u], tvz[u], f0.x, f0.y, f0.z,
                         f0.w);
          } else {
            const float4 f1 = T.P(1)[slot];
            iact_gradient(gacc[u], r2, dx, dy, dz, th[u], tvx[u], tvy[u], tvz[u], tu[u], tcs[u], f0.x,
                          f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, A.a2_Hubble);
          }
          nhit[u]++;
        }
      }
    }
    nst = 0;
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      nlist[u] = 0;
      lp[u] = T.list() + u * 32 + lane;
    }
    __syncwarp();
  };

  for (int it = 0; it < G.item_count; it++) {
    const Item I = A.items[G.item_first + it];
    const DevCell sc = A.cells[I.scell];
    const int mode = I.mode;
    const int sid = I.sid;
    const int scount = sc.count;

    /* item-level (warp-uniform) frame */
    bool ascending = true;
    double fsx = 0., fsy = 0., fsz = 0.; /* frame origin of the staged source floats */
    double otx = 0., oty = 0., otz = 0.; /* subtracted from the target double in the drain */
    const bool dbl_mode = (mode == MODE_SELF || mode == MODE_SUB_PAIR || mode == MODE_SUB_PAIR_F);
    const bool sorted = (mode == MODE_PAIR_L || mode == MODE_PAIR_R || mode == MODE_SUB_PAIR ||
                         mode == MODE_SUB_PAIR_F);
    const bool is_pair = (mode == MODE_PAIR_L || mode == MODE_PAIR_R);
    const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1],
                 shz = I.shift[2] * A.dim[2];
    int64_t soff = 0;
    if (sorted) soff = sort_offset(sc, sid);
    double rshift = 0., lim_a = 0., lim_b = 0.; /* pair: hi_max / dj_min or hj_max / di_max */
    float dx_max = 0.f;

    if (is_pair) {
      /* oriented pair: ci = left cell, cj = right cell */
      const DevCell tc = A.cells[I.tcell];
      const DevCell &ci = (mode == MODE_PAIR_L) ? tc : sc;
      const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tc;
      rshift = __dadd_rn(
          __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
          __dmul_rn(shz, c_runner_shift[sid][2]));
      const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
      dx_max = __fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
      /* frame origins: ci particles are shifted by cj->loc + shift, cj by cj->loc */
      const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                   oiz = __dadd_rn(cj.loc[2], shz);
      if (mode == MODE_PAIR_L) {
        /* targets in ci: functions_hydro.h:1296-1332 */
        lim_a = __dsub_rn((double)__fmul_rn(fminf(h_max_lim, ci.h_max_active), KERNEL_GAMMA), rshift);
        /* dj_min = sort_j[0].d */
        const int j0 = cj.first + (int)A.sort_idx[sort_offset(cj, sid)];
        lim_b = (double)sort_key(A.x[3 * (size_t)j0], A.x[3 * (size_t)j0 + 1], A.x[3 * (size_t)j0 + 2],
                                 sid);
        otx = oix;
        oty = oiy;
        otz = oiz;
        fsx = cj.loc[0];
        fsy = cj.loc[1];
        fsz = cj.loc[2];
        ascending = true;
      } else {
        /* targets in cj: functions_hydro.h:1420-1448 */
        lim_a = (double)__fmul_rn(fminf(h_max_lim, cj.h_max_active), KERNEL_GAMMA);
        const int i1 = ci.first + (int)A.sort_idx[sort_offset(ci, sid) + ci.count - 1];
        lim_b = __dsub_rn((double)sort_key(A.x[3 * (size_t)i1], A.x[3 * (size_t)i1 + 1],
                                           A.x[3 * (size_t)i1 + 2], sid),
                          rshift);
        otx = cj.loc[0];
        oty = cj.loc[1];
        otz = cj.loc[2];
        fsx = oix;
        fsy = oiy;
        fsz = oiz;
        ascending = false;
      }
    } else if (mode == MODE_SUB_SELF) {
      /* DOSELF_SUBSET :1127-1129: floats relative to c->loc (c = scell) */
      otx = fsx = sc.loc[0];
      oty = fsy = sc.loc[1];
      otz = fsz = sc.loc[2];
    } else {
      /* Double modes. MODE_SELF: dx = (float)(x_t - x_s). MODE_SUB_PAIR*:
       * (float)((x_t - shift) - x_s), DOPAIR_SUBSET :885-897 / :955-967. The
       * prefilter works on floats relative to the source cell; its radius is
       * widened by the rounding of those floats. */
      if (mode != MODE_SELF) {
        otx = shx;
        oty = shy;
        otz = shz;
        ascending = (mode == MODE_SUB_PAIR);
      }
      fsx = sc.loc[0];
      fsy = sc.loc[1];
      fsz = sc.loc[2];
    }

    /* per-target participation, prefilter position and key threshold */
    float tp[TPL][3], r2e[TPL], thr[TPL];
    bool anypart = false;
    float reach_l = ascending ? -3.0e38f : 3.0e38f;
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      bool part = tvalid[u] && tdepth[u] >= I.min_depth && tdepth[u] <= I.max_depth;
      thr[u] = 0.f;
      r2e[u] = __fmul_rn(thg2[u], PREFILTER_REL);
      if (is_pair) {
        const float tkey = sort_key(tx[u], ty[u], tz[u], sid);
        if (mode == MODE_PAIR_L) {
          const bool in_loop = __dadd_rn(__dadd_rn((double)tkey, lim_a), (double)dx_max) > lim_b;
          const double di = __dsub_rn((double)__fadd_rn(__fadd_rn(tkey, thg[u]), dx_max), rshift);
          part = part && in_loop && !(di < lim_b);
          thr[u] = __double2float_ru(di); /* key < di  <=>  key < ru(di) for float keys */
        } else {
          const bool in_loop = __dsub_rn(__dsub_rn((double)tkey, lim_a), (double)dx_max) < lim_b;
          const double dj = __dadd_rn((double)__fsub_rn(__fsub_rn(tkey, thg[u]), dx_max), rshift);
          part = part && in_loop && !(__dsub_rn(dj, rshift) > lim_b);
          thr[u] = __double2float_rd(dj); /* key > dj  <=>  key > rd(dj) */
        }
        tp[u][0] = dsubf(tx[u], otx);
        tp[u][1] = dsubf(ty[u], oty);
        tp[u][2] = dsubf(tz[u], otz);
      } else if (mode == MODE_SUB_SELF) {
        tp[u][0] = dsubf(tx[u], otx);
        tp[u][1] = dsubf(ty[u], oty);
        tp[u][2] = dsubf(tz[u], otz);
      } else {
        const double tdx = __dsub_rn(tx[u], otx), tdy = __dsub_rn(ty[u], oty),
                     tdz = __dsub_rn(tz[u], otz);
        tp[u][0] = dsubf(tdx, fsx);
        tp[u][1] = dsubf(tdy, fsy);
        tp[u][2] = dsubf(tdz, fsz);
        const float re = fmaf(thg[u], PREFILTER_REL, 4.0e-6f * sc.width);
        r2e[u] = re * re;
        if (mode != MODE_SELF) {
          /* di = hi*kernel_gamma + dxj + pix*rs0 + piy*rs1 + piz*rs2, left to right */
          const float dxj = sc.dx_max_sort;
          const float f0 = (mode == MODE_SUB_PAIR) ? __fadd_rn(thg[u], dxj) : __fsub_rn(-thg[u], dxj);
          const double di =
              __dadd_rn(__dadd_rn(__dadd_rn((double)f0, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                  __dmul_rn(tdy, c_runner_shift[sid][1])),
                        __dmul_rn(tdz, c_runner_shift[sid][2]));
          thr[u] = (mode == MODE_SUB_PAIR) ? __double2float_ru(di) : __double2float_rd(di);
        }
      }
      if (part) {
        anypart = true;
        reach_l = ascending ? fmaxf(reach_l, thr[u]) : fminf(reach_l, thr[u]);
      } else {
        tp[u][0] = 3.0e30f; /* never passes the prefilter */
      }
    }
    if (!__any_sync(FULL_MASK, anypart)) continue;
    /* reach of the warp along the axis, for the sorted early exit */
    float reach = 0.f;
    if (sorted) reach = ascending ? warp_max(reach_l) : warp_min(reach_l);

    for (int base = 0; base < scount; base += 32) {
      /* ---- stage 32 sources ---- */
      const int k = base + lane;
      int sj = -1;
      float skey = ascending ? 3.0e38f : -3.0e38f; /* padding never passes the prune */
      double sx = 0., sy = 0., sz = 0.;
      if (k < scount) {
        int local = k;
        if (sorted) local = (int)A.sort_idx[soff + (ascending ? k : scount - 1 - k)];
        sj = sc.first + local;
        sx = A.x[3 * (size_t)sj];
        sy = A.x[3 * (size_t)sj + 1];
        sz = A.x[3 * (size_t)sj + 2];
        if (sorted) skey = sort_key(sx, sy, sz, sid);
      }
      /* sorted early exit: first key of the chunk already out of everyone's reach */
      if (sorted) {
        const float first_key = __shfl_sync(FULL_MASK, skey, 0);
        if (ascending ? !(first_key < reach) : !(first_key > reach)) break;
      }
      {
        bool full = nst + 32 > SCAP1;
#pragma unroll
        for (int u = 0; u < TPL; u++) full = full || (nlist[u] > LCAP - 32);
        if (__any_sync(FULL_MASK, full)) drain();
      }
      const int slot = nst + lane;
      if (k < scount) {
        T.F()[slot] = make_float4(dsubf(sx, fsx), dsubf(sy, fsy), dsubf(sz, fsz), skey);
        T.P(0)[slot] = A.mv[sj];
        if (LOOP == LOOP_GRADIENT) {
          const float4 q1 = A.fq1[sj];
          const float4 q2 = A.fq2[sj];
          const float4 q3 = A.fq3[sj];
          T.P(1)[slot] = make_float4(q2.z /*u*/, q1.x /*rho*/, q1.w /*cs*/, q3.x /*alpha*/);
        }
        if (SUBSET) {
          double *D = T.D();
          D[slot] = sx;
          D[SCAP1 + slot] = sy;
          D[2 * SCAP1 + slot] = sz;
        }
      } else {
        T.F()[slot] = make_float4(-3.0e30f, -3.0e30f, -3.0e30f, skey);
      }
      T.GI()[slot] = sj;
      if (lane == 0) {
        ChunkInfo ci;
        ci.ot[0] = otx;
        ci.ot[1] = oty;
        ci.ot[2] = otz;
        ci.dbl = dbl_mode ? 1 : 0;
        ci.item = G.item_first + it;
        T.chunk()[nst >> 5] = ci;
      }
      __syncwarp();

      /* ---- test ---- */
      nchunks++;
      if (!sorted)
        test_chunk<0>(T.F(), nst, tp, r2e, thr, lp, nlist);
      else if (ascending)
        test_chunk<1>(T.F(), nst, tp, r2e, thr, lp, nlist);
      else
        test_chunk<2>(T.F(), nst, tp, r2e, thr, lp, nlist);
      nst += 32;
    }
  }
  drain();

  /* ---- flush (once per task) ---- */
  int tot = 0;
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    if (tvalid[u]) {
      const int p = ti[u];
      if (LOOP == LOOP_DENSITY) {
        float *pa = (float *)&A.dA[p];
        float *pb = (float *)&A.dB[p];
        atomicAdd(pa + 0, dacc[u].rho);
        atomicAdd(pa + 1, dacc[u].rho_dh);
        atomicAdd(pa + 2, dacc[u].wcount);
        atomicAdd(pa + 3, dacc[u].wcount_dh);
        atomicAdd(pb + 0, dacc[u].div_v);
        atomicAdd(pb + 1, dacc[u].rot[0]);
        atomicAdd(pb + 2, dacc[u].rot[1]);
        atomicAdd(pb + 3, dacc[u].rot[2]);
      } else {
        atomic_max_pos(&A.g_vsig[p], gacc[u].v_sig);
        atomicAdd(&A.g_lap[p], gacc[u].laplace_u);
        atomic_max_pos(&A.g_amax[p], gacc[u].alpha_max);
      }
      if (nhit[u]) atomicAdd(&A.count[p], nhit[u]);
    }
    tot += nhit[u];
  }
  /* global interaction counter */
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && nchunks)
    atomicAdd(A.tests, (unsigned long long)nchunks * (unsigned long long)(1024 * TPL));
}

/* ------------------------------------------------------------------------ */
/* Type-2 loop: force. Hit criterion r2 < max(h_t, h_s)^2 gamma^2, with the
 * two-pass pruning of DOPAIR2 restated per (i in ci, j in cj):
 *   pass A (:1737-1975): i in A-range, key_j < di_i, r2 < hig2
 *   pass B (:1978-2230): j in B-range, key_i - rshift > dj_j, hig2 <= r2 < hjg2
 * The prefilter only applies the radius; the pass conditions (doubles) are
 * evaluated per candidate in the drain.                                      */
/* ------------------------------------------------------------------------ */
#define SCAP2 192
template <int SCHEME>
struct Tile2 : Tile<SCAP2, (SCHEME == SCH_SPHENIX ? 4 : 3), false, true, ChunkInfoF> {};

template <int SCHEME>
__global__ void __launch_bounds__(32) k_loop2(const LoopArgs A) {
  extern __shared__ __align__(16) char smem_raw[];
  typedef Tile2<SCHEME> TT;
  TT T;
  T.base = smem_raw;
  const int lane = threadIdx.x;
  const int task = blockIdx.x;

  const int g = A.task_group[task];
  const int chunk = A.task_chunk[task];
  const int nt = A.tgt_count[g];
  if (chunk * TASK_TARGETS >= nt) return;
  const Group G = A.groups[g];

  bool tvalid[TPL];
  int ti[TPL], tdepth[TPL], nhit[TPL];
  double tx[TPL], ty[TPL], tz[TPL];
  ForceQ tq[TPL];
  float thg2[TPL], thg[TPL];
  ForceAcc acc[TPL];
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    const int slot_t = chunk * TASK_TARGETS + u * 32 + lane;
    tvalid[u] = slot_t < nt;
    ti[u] = tvalid[u] ? A.tgt_list[A.tgt_first[g] + slot_t] : -1;
    tx[u] = ty[u] = tz[u] = 0.;
    tq[u].m = tq[u].vx = tq[u].vy = tq[u].vz = 0.f;
    tq[u].rho = 1.f;
    tq[u].P = tq[u].f = tq[u].cs = tq[u].balsara = 0.f;
    tq[u].h = 1.f;
    tq[u].u = tq[u].alpha_visc = tq[u].alpha_diff = 0.f;
    tq[u].time_bin = 0;
    tdepth[u] = 0;
    if (tvalid[u]) {
      const size_t p = (size_t)ti[u];
      tx[u] = A.x[3 * p];
      ty[u] = A.x[3 * p + 1];
      tz[u] = A.x[3 * p + 2];
      const float4 q0 = A.mv[p], q1 = A.fq1[p], q2 = A.fq2[p];
      tq[u].m = q0.x; tq[u].vx = q0.y; tq[u].vy = q0.z; tq[u].vz = q0.w;
      tq[u].rho = q1.x; tq[u].P = q1.y; tq[u].f = q1.z; tq[u].cs = q1.w;
      tq[u].balsara = q2.x; tq[u].h = q2.y; tq[u].u = q2.z; tq[u].time_bin = __float_as_int(q2.w);
      if (SCHEME == SCH_SPHENIX) {
        const float4 q3 = A.fq3[p];
        tq[u].alpha_visc = q3.x;
        tq[u].alpha_diff = q3.y;
      }
      tdepth[u] = A.depth_h[p];
    }
    thg2[u] = hg2_exact(tq[u].h);
    thg[u] = __fmul_rn(tq[u].h, KERNEL_GAMMA);
    acc[u].ax = acc[u].ay = acc[u].az = acc[u].u_dt = acc[u].h_dt = 0.f;
    acc[u].v_sig = 0.f;
    acc[u].min_ngb = NUM_TIME_BINS + 1;
    nhit[u] = 0;
  }
  int nchunks = 0;

  int nst = 0;
  int nlist[TPL];
  uint16_t *lp[TPL];
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    nlist[u] = 0;
    lp[u] = T.list() + u * 32 + lane;
  }

  /* ---- INTERACT ---- */
  auto drain = [&]() {
    __syncwarp();
    int mx = nlist[0];
#pragma unroll
    for (int u = 1; u < TPL; u++) mx = max(mx, nlist[u]);
    const int maxn = __reduce_max_sync(FULL_MASK, mx);
    const float4 *F = T.F();
    const int32_t *GI = T.GI();
    const float *K = T.K();
    const ChunkInfoF *CI = T.chunk();
    const uint16_t *l0 = T.list() + lane;
    for (int k = 0; k < maxn; k++) {
#pragma unroll
      for (int u = 0; u < TPL; u++) {
        const bool act = k < nlist[u];
        const int slot = act ? (int)l0[k * TASK_TARGETS + u * 32] : 0;
        const ChunkInfoF &ci = CI[slot >> 5];
        const int gi = GI[slot];
        const float4 s = F[slot];
        const bool dbl = ci.dbl != 0;
        double sxd = 0., syd = 0., szd = 0.;
        if (act && dbl) {
          sxd = A.x[3 * (size_t)gi];
          syd = A.x[3 * (size_t)gi + 1];
          szd = A.x[3 * (size_t)gi + 2];
        }
        const float spx = dbl ? 0.f : s.x, spy = dbl ? 0.f : s.y, spz = dbl ? 0.f : s.z;
        const float dx = __fsub_rn(dsubf(__dsub_rn(tx[u], ci.ot[0]), sxd), spx);
        const float dy = __fsub_rn(dsubf(__dsub_rn(ty[u], ci.ot[1]), syd), spy);
        const float dz = __fsub_rn(dsubf(__dsub_rn(tz[u], ci.ot[2]), szd), spz);
        const float r2 = r2_exact(dx, dy, dz);
        const float4 q2 = T.P(2)[slot];
        const float sh = q2.y;
        const float shg2 = hg2_exact(sh);
        bool ok;
        if (dbl) {
          /* DOSELF2 :2792: doi = r2 < hig2 || r2 < hjg2 */
          ok = (r2 < thg2[u] || r2 < shg2) && (gi != ti[u]);
        } else {
          /* DOPAIR2: the pair constants of the item */
          const bool tleft = (ci.sid & 256) != 0;
          const int sid = ci.sid & 255;
          const double rshift = ci.rshift, hi_max_g = ci.hi_max_g, hj_max_g = ci.hj_max_g,
                       dx_max = ci.dx_max, di_max_sh = ci.di_max_sh, dj_min = ci.dj_min;
          const float tkey = sort_key(tx[u], ty[u], tz[u], sid);
          const float skey = K[slot];
          const float shg = __fmul_rn(sh, KERNEL_GAMMA);
          if (tleft) {
            /* t = i in ci, s = j in cj */
            const bool inA =
                __dsub_rn(__dadd_rn(__dadd_rn((double)tkey, hi_max_g), dx_max), rshift) > dj_min;
            const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(tkey, thg[u]), dx_max), rshift);
            const double t_di = (inA && !(di < dj_min)) ? di : -1.0e300;
            const double t_keysh = __dsub_rn((double)tkey, rshift);
            const bool inB = __dsub_rn(__dsub_rn((double)skey, hj_max_g), dx_max) < di_max_sh;
            const double dj = __dsub_rn((double)__fsub_rn(skey, shg), dx_max);
            const double s_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
            const bool c1 = ((double)skey < t_di) && (r2 < thg2[u]);
            const bool c2 = (t_keysh > s_dj) && (r2 < shg2) && !(r2 < thg2[u]);
            ok = c1 || c2;
          } else {
            /* t = j in cj, s = i in ci */
            const bool inB = __dsub_rn(__dsub_rn((double)tkey, hj_max_g), dx_max) < di_max_sh;
            const double dj = __dsub_rn((double)__fsub_rn(tkey, thg[u]), dx_max);
            const double t_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
            const bool inA =
                __dsub_rn(__dadd_rn(__dadd_rn((double)skey, hi_max_g), dx_max), rshift) > dj_min;
            const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(skey, shg), dx_max), rshift);
            const double s_di = (inA && !(di < dj_min)) ? di : -1.0e300;
            const double s_keysh = __dsub_rn((double)skey, rshift);
            const bool c1 = ((double)tkey < s_di) && (r2 < shg2);
            const bool c2 = (s_keysh > t_dj) && (r2 < thg2[u]) && !(r2 < shg2);
            ok = c1 || c2;
          }
        }
        if (act && ok) {
          ForceQ sq;
          const float4 q0 = T.P(0)[slot], q1 = T.P(1)[slot];
          sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
          sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
          sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
          sq.alpha_visc = sq.alpha_diff = 0.f;
          if (SCHEME == SCH_SPHENIX) {
            const float4 q3 = T.P(3)[slot];
            sq.alpha_visc = q3.x;
            sq.alpha_diff = q3.y;
          }
          iact_force<SCHEME>(acc[u], r2, dx, dy, dz, tq[u], sq, A.a2_Hubble);
          nhit[u]++;
        }
      }
    }
    nst = 0;
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      nlist[u] = 0;
      lp[u] = T.list() + u * 32 + lane;
    }
    __syncwarp();
  };

  for (int it = 0; it < G.item_count; it++) {
    const Item I = A.items[G.item_first + it];
    const DevCell sc = A.cells[I.scell];
    const int mode = I.mode;
    const int sid = I.sid;
    const int scount = sc.count;

    double otx = 0., oty = 0., otz = 0.;
    double fsx, fsy, fsz;
    bool sorted = false, tleft = true;
    double hi_max_g = 0., hj_max_g = 0., dx_max = 0., rshift = 0.;
    double di_max_sh = 0., dj_min = 0.;
    int64_t soff = 0;
    float wadd = 0.f; /* widening of the radii in the double mode */

    if (mode == MODE_SELF) {
      /* DOSELF2 :2624-2875 */
      fsx = sc.loc[0];
      fsy = sc.loc[1];
      fsz = sc.loc[2];
      otx = fsx; /* prefilter frame only; the drain uses ot = 0 */
      oty = fsy;
      otz = fsz;
      wadd = 4.0e-6f * sc.width;
    } else {
      /* ---- DOPAIR2 ---- */
      sorted = true;
      const DevCell tc = A.cells[I.tcell];
      tleft = (mode == MODE_PAIR_L);
      const DevCell &ci = tleft ? tc : sc;
      const DevCell &cj = tleft ? sc : tc;
      const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1],
                   shz = I.shift[2] * A.dim[2];
      rshift = __dadd_rn(
          __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
          __dmul_rn(shz, c_runner_shift[sid][2]));
      hi_max_g = __dmul_rn((double)ci.h_max, (double)KERNEL_GAMMA);
      hj_max_g = __dmul_rn((double)cj.h_max, (double)KERNEL_GAMMA);
      dx_max = (double)__fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
      const int64_t soff_i = sort_offset(ci, sid), soff_j = sort_offset(cj, sid);
      const int i1 = ci.first + (int)A.sort_idx[soff_i + ci.count - 1];
      const int j0 = cj.first + (int)A.sort_idx[soff_j];
      const double di_max = (double)sort_key(A.x[3 * (size_t)i1], A.x[3 * (size_t)i1 + 1],
                                             A.x[3 * (size_t)i1 + 2], sid);
      dj_min = (double)sort_key(A.x[3 * (size_t)j0], A.x[3 * (size_t)j0 + 1],
                                A.x[3 * (size_t)j0 + 2], sid);
      di_max_sh = __dsub_rn(di_max, rshift);
      const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                   oiz = __dadd_rn(cj.loc[2], shz);
      if (tleft) {
        otx = oix;
        oty = oiy;
        otz = oiz;
        fsx = cj.loc[0];
        fsy = cj.loc[1];
        fsz = cj.loc[2];
        soff = soff_j;
      } else {
        otx = cj.loc[0];
        oty = cj.loc[1];
        otz = cj.loc[2];
        fsx = oix;
        fsy = oiy;
        fsz = oiz;
        soff = soff_i;
      }
    }

    float tp[TPL][3], r2e[TPL], thr_unused[TPL];
    bool anypart = false;
    double reach_l = tleft ? -1.0e300 : 1.0e300;
#pragma unroll
    for (int u = 0; u < TPL; u++) {
      const bool part = tvalid[u] && tdepth[u] >= I.min_depth && tdepth[u] <= I.max_depth;
      thr_unused[u] = 0.f;
      tp[u][0] = dsubf(tx[u], otx);
      tp[u][1] = dsubf(ty[u], oty);
      tp[u][2] = dsubf(tz[u], otz);
      if (!sorted) {
        const float re = fmaf(thg[u], PREFILTER_REL, wadd);
        r2e[u] = re * re;
      } else {
        r2e[u] = __fmul_rn(thg2[u], PREFILTER_REL);
        /* Conservative reach along the axis for the sorted early exit. */
        const float tkey = sort_key(tx[u], ty[u], tz[u], sid);
        if (tleft) {
          double t_di = -1.0e300;
          const bool inA =
              __dsub_rn(__dadd_rn(__dadd_rn((double)tkey, hi_max_g), dx_max), rshift) > dj_min;
          const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(tkey, thg[u]), dx_max), rshift);
          if (inA && !(di < dj_min)) t_di = di;
          const double t_keysh = __dsub_rn((double)tkey, rshift);
          /* sources j ascending; j can matter while key_j - hj_max*g - dx_max <= max(di, keysh) */
          if (part) reach_l = fmax(reach_l, fmax(t_di, t_keysh));
        } else {
          double t_dj = 1.0e300;
          const bool inB = __dsub_rn(__dsub_rn((double)tkey, hj_max_g), dx_max) < di_max_sh;
          const double dj = __dsub_rn((double)__fsub_rn(tkey, thg[u]), dx_max);
          if (inB && !(dj > di_max_sh)) t_dj = dj;
          /* sources i descending; i can matter while key_i + hi_max*g + dx_max - rshift >= min(key_t, dj) */
          if (part) reach_l = fmin(reach_l, fmin((double)tkey, t_dj));
        }
      }
      if (part)
        anypart = true;
      else
        tp[u][0] = 3.0e30f;
    }
    if (!__any_sync(FULL_MASK, anypart)) continue;
    double reach = 0.;
    if (sorted) reach = tleft ? warp_max_d(reach_l) : warp_min_d(reach_l);
    if (!sorted) otx = oty = otz = 0.; /* MODE_SELF: the drain subtracts the doubles directly */

    for (int base = 0; base < scount; base += 32) {
      const int k = base + lane;
      int sj = -1;
      float skey = tleft ? 3.0e38f : -3.0e38f;
      double sx = 0., sy = 0., sz = 0.;
      if (k < scount) {
        int local = k;
        if (sorted) local = (int)A.sort_idx[soff + (tleft ? k : scount - 1 - k)];
        sj = sc.first + local;
        sx = A.x[3 * (size_t)sj];
        sy = A.x[3 * (size_t)sj + 1];
        sz = A.x[3 * (size_t)sj + 2];
        if (sorted) skey = sort_key(sx, sy, sz, sid);
      }
      if (sorted) {
        /* early exit on the sorted axis (conservative, with a rounding slack) */
        const double fk = (double)__shfl_sync(FULL_MASK, skey, 0);
        const double slack = 1.0e-5 * (fabs(fk) + 1.0);
        if (tleft) {
          if (fk - hj_max_g - dx_max - slack > reach) break;
        } else {
          if (fk + hi_max_g + dx_max - rshift + slack < reach) break;
        }
      }
      {
        bool full = nst + 32 > SCAP2;
#pragma unroll
        for (int u = 0; u < TPL; u++) full = full || (nlist[u] > LCAP - 32);
        if (__any_sync(FULL_MASK, full)) drain();
      }
      const int slot = nst + lane;
      if (k < scount) {
        const float4 q2 = A.fq2[sj];
        const float sh = q2.y;
        float w;
        if (sorted) {
          w = __fmul_rn(hg2_exact(sh), PREFILTER_REL);
        } else {
          const float re = fmaf(__fmul_rn(sh, KERNEL_GAMMA), PREFILTER_REL, wadd);
          w = re * re;
        }
        T.F()[slot] = make_float4(dsubf(sx, fsx), dsubf(sy, fsy), dsubf(sz, fsz), w);
        T.K()[slot] = skey;
        T.P(0)[slot] = A.mv[sj];
        T.P(1)[slot] = A.fq1[sj];
        T.P(2)[slot] = q2;
        if (SCHEME == SCH_SPHENIX) T.P(3)[slot] = A.fq3[sj];
      } else {
        T.F()[slot] = make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
        T.K()[slot] = skey;
        T.P(2)[slot] = make_float4(0.f, 1.f, 0.f, 0.f);
      }
      T.GI()[slot] = sj;
      if (lane == 0) {
        ChunkInfoF ci;
        ci.ot[0] = otx;
        ci.ot[1] = oty;
        ci.ot[2] = otz;
        ci.rshift = rshift;
        ci.hi_max_g = hi_max_g;
        ci.hj_max_g = hj_max_g;
        ci.dx_max = dx_max;
        ci.di_max_sh = di_max_sh;
        ci.dj_min = dj_min;
        ci.dbl = sorted ? 0 : 1;
        ci.sid = sid | (tleft ? 256 : 0);
        T.chunk()[nst >> 5] = ci;
      }
      __syncwarp();
      nchunks++;
      test_chunk<3>(T.F(), nst, tp, r2e, thr_unused, lp, nlist);
      nst += 32;
    }
  }
  drain();

  int tot = 0;
#pragma unroll
  for (int u = 0; u < TPL; u++) {
    if (tvalid[u]) {
      const int p = ti[u];
      float *po = (float *)&A.fo1[p];
      atomicAdd(po + 0, acc[u].ax);
      atomicAdd(po + 1, acc[u].ay);
      atomicAdd(po + 2, acc[u].az);
      atomicAdd(po + 3, acc[u].u_dt);
      atomicAdd(&A.f_hdt[p], acc[u].h_dt);
      if (SCHEME != SCH_SPHENIX) atomic_max_pos(&A.f_vsig[p], acc[u].v_sig);
      atomicMin(&A.f_minngb[p], acc[u].min_ngb);
      if (nhit[u]) atomicAdd(&A.count[p], nhit[u]);
    }
    tot += nhit[u];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && nchunks)
    atomicAdd(A.tests, (unsigned long long)nchunks * (unsigned long long)(1024 * TPL));
}

}  // namespace swiftgpu
#endif
