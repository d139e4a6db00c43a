/*
 * loops.cuh - the neighbour-loop kernels (K1/K2/K3 of SURVEY 2.1).
 *
 * One warp owns up to 32 TARGET particles of one target cell (a "task") and
 * walks every directed item of that cell's group (worklist.hpp). Per item it
 * streams the SOURCE cell in chunks of 32 particles through a warp-private
 * shared-memory tile (positions already converted to the item's float frame,
 * plus the sort key), and every lane tests its own target against the 32
 * staged sources with the reference's exact arithmetic, recording hits in a
 * 32-bit mask. The hits are then drained lane-parallel with the (FMA) non-
 * symmetric interaction; accumulators stay in registers for the whole group
 * and are flushed once per task (no atomics in the inner loop).
 *
 * Pseudo-Verlet pruning: pair items stream the source cell in sorted-axis
 * order and stop as soon as the first key of a chunk is beyond the reach of
 * every lane (a warp max / min taken with shuffles); inside a chunk the
 * reference's `sort_j[pjd].d < di` test is a single float compare against a
 * per-lane threshold rounded towards +inf (equivalent for float keys).
 *
 * Reference semantics reproduced (runner_doiact_functions_hydro.h):
 *   MODE_SELF      DOSELF1 :2299 / DOSELF2 :2624   dx = (float)(x_t - x_s) on doubles
 *   MODE_PAIR_L/R  DOPAIR1 :1234 / DOPAIR2 :1601   floats in the frame cj->loc (+shift)
 *   MODE_SUB_SELF  DOSELF_SUBSET :1108             floats relative to c->loc
 *   MODE_SUB_PAIR  DOPAIR_SUBSET :855              (float)((x_t - shift) - x_s) on doubles
 */
#ifndef SWIFTGPU_LOOPS_CUH
#define SWIFTGPU_LOOPS_CUH

#include "sph_math.cuh"
#include "worklist.hpp"

namespace swiftgpu {

#define FULL_MASK 0xffffffffu
#define WARPS_PER_BLOCK 4

/* Device view of one cell (64 bytes). */
struct DevCell {
  double loc[3];
  int32_t first;
  int32_t count;
  float h_max;
  float h_max_active;
  float dx_max_sort;
  float h_max_allowed;
  float h_min_allowed;
  int32_t parent;
  int64_t sort_base; /* offset of this cell's first sorted index array, -1 if none */
  uint16_t sort_mask; /* which sids are present */
  int8_t depth;
  uint8_t flags; /* bit0 active, bit1 local, bit2 split */
  int32_t pad_;
};
static_assert(sizeof(DevCell) == 72, "DevCell layout");

__device__ __forceinline__ int64_t sort_offset(const DevCell &c, int sid) {
  return c.sort_base + (int64_t)__popc((unsigned)c.sort_mask & ((1u << sid) - 1u)) * c.count;
}

struct LoopArgs {
  const DevCell *cells;
  const Item *items;
  const Group *groups;
  const int32_t *task_group; /* per task */
  const int32_t *task_chunk;
  int ntasks;
  const int32_t *tgt_list;  /* target particle indices */
  const int32_t *tgt_first; /* per group: offset into tgt_list */
  const int32_t *tgt_count; /* per group */
  const uint32_t *sort_idx;
  /* particle state */
  const double *x;       /* 3n */
  const float4 *mv;      /* (m, vx, vy, vz) */
  const float *h;
  const int8_t *depth_h;
  const int8_t *time_bin;
  /* gradient / force inputs */
  const float4 *fq1; /* (rho, P, f, cs) */
  const float4 *fq2; /* (balsara, h, u, time_bin) */
  const float4 *fq3; /* (alpha_visc, alpha_diff, -, -) */
  /* outputs */
  float4 *dA;      /* (rho, rho_dh, wcount, wcount_dh) */
  float4 *dB;      /* (div_v, rot_v) */
  float *g_vsig;   /* gradient: viscosity.v_sig (max) */
  float *g_lap;    /* gradient: diffusion.laplace_u (sum) */
  float *g_amax;   /* gradient: force.alpha_visc_max_ngb (max) */
  float4 *fo1;     /* (ax, ay, az, u_dt) */
  float *f_hdt;
  float *f_vsig;
  int32_t *f_minngb;
  int32_t *count; /* per-particle directed interaction counter of this loop */
  unsigned long long *total; /* global interaction counter */
  unsigned long long *tests; /* global distance-test counter */
  double dim[3];
  float a2_Hubble;
  int max_active_bin;
};

/* warp-private staging tile */
struct __align__(16) SrcTile {
  float4 pos[32];   /* frame floats + key (float modes) */
  double xd[32][3]; /* absolute / shifted doubles (double modes) */
  float4 f0[32];    /* (m, vx, vy, vz) */
  float4 f1[32];    /* loop-dependent source fields */
  float4 f2[32];
  float4 f3[32];
  double dB[32];    /* force pass-B threshold of the source */
  double dA[32];    /* force: di of the source (MODE_PAIR_R) */
  int32_t idx[32];
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

__device__ __forceinline__ void atomic_max_pos(float *addr, float v) {
  /* non-negative floats order like their bit patterns */
  atomicMax((int *)addr, __float_as_int(v));
}

/* ------------------------------------------------------------------------ */
/* Type-1 loops: density (all schemes), gradient (SPHENIX), and the density
 * subset re-runs of the ghost. Hit criterion r2 < h_t^2 gamma^2.            */
/* ------------------------------------------------------------------------ */
template <int LOOP>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
    k_loop1(const LoopArgs A) {
  __shared__ SrcTile tiles[WARPS_PER_BLOCK];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int task = blockIdx.x * WARPS_PER_BLOCK + wib;
  if (task >= A.ntasks) return;
  SrcTile &T = tiles[wib];

  const int g = A.task_group[task];
  const int chunk = A.task_chunk[task];
  const int nt = A.tgt_count[g];
  if (chunk * 32 >= nt) return;
  const Group G = A.groups[g];
  const int slot = chunk * 32 + lane;
  const bool tvalid = slot < nt;
  const int ti = tvalid ? A.tgt_list[A.tgt_first[g] + slot] : -1;

  /* target state */
  double tx = 0., ty = 0., tz = 0.;
  float th = 1.f, tvx = 0.f, tvy = 0.f, tvz = 0.f;
  float tu = 0.f, tcs = 0.f;
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
  const float thg = __fmul_rn(th, KERNEL_GAMMA); /* hi * kernel_gamma (float) */

  DensityAcc dacc;
  dacc.zero();
  GradientAcc gacc;
  gacc.v_sig = 0.f;
  gacc.laplace_u = 0.f;
  gacc.alpha_max = 0.f;
  int nhit = 0;
  int nchunks = 0;

  for (int it = 0; it < G.item_count; it++) {
    const Item I = A.items[G.item_first + it];
    const DevCell sc = A.cells[I.scell];
    const int mode = I.mode;
    const int sid = I.sid;
    const int scount = sc.count;

    /* per-lane participation and frame */
    bool part = tvalid && tdepth >= I.min_depth && tdepth <= I.max_depth;
    float tpx = 0.f, tpy = 0.f, tpz = 0.f; /* float-frame position */
    double tdx = tx, tdy = ty, tdz = tz;   /* double position (minus shift) */
    float thr = 0.f;                       /* pruning threshold on the source key */
    bool ascending = true;
    double fsx = 0., fsy = 0., fsz = 0.; /* frame origin of the sources */
    const bool dbl_mode = (mode == MODE_SELF || mode == MODE_SUB_PAIR || mode == MODE_SUB_PAIR_F);
    const bool sorted = (mode == MODE_PAIR_L || mode == MODE_PAIR_R || mode == MODE_SUB_PAIR ||
                         mode == MODE_SUB_PAIR_F);
    const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1],
                 shz = I.shift[2] * A.dim[2];
    int64_t soff = 0;
    if (sorted) soff = sort_offset(sc, sid);

    if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
      /* oriented pair: ci = left cell, cj = right cell */
      const DevCell tc = A.cells[I.tcell];
      const DevCell &ci = (mode == MODE_PAIR_L) ? tc : sc;
      const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tc;
      const double rshift = __dadd_rn(
          __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
          __dmul_rn(shz, c_runner_shift[sid][2]));
      const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
      const float dx_max = __fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
      /* frame origins: ci particles are shifted by cj->loc + shift, cj by cj->loc */
      const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                   oiz = __dadd_rn(cj.loc[2], shz);
      const float tkey = sort_key(tx, ty, tz, sid);
      if (mode == MODE_PAIR_L) {
        /* targets in ci: functions_hydro.h:1296-1332 */
        const double hi_max =
            __dsub_rn((double)__fmul_rn(fminf(h_max_lim, ci.h_max_active), KERNEL_GAMMA), rshift);
        /* dj_min = sort_j[0].d */
        const int j0 = cj.first + (int)A.sort_idx[sort_offset(cj, sid)];
        const double dj_min =
            (double)sort_key(A.x[3 * (size_t)j0], A.x[3 * (size_t)j0 + 1], A.x[3 * (size_t)j0 + 2], sid);
        const bool in_loop = __dadd_rn(__dadd_rn((double)tkey, hi_max), (double)dx_max) > dj_min;
        const double di = __dsub_rn((double)__fadd_rn(__fadd_rn(tkey, thg), dx_max), rshift);
        part = part && in_loop && !(di < dj_min);
        thr = __double2float_ru(di); /* key < di  <=>  key < ru(di) for float keys */
        tpx = dsubf(tx, oix);
        tpy = dsubf(ty, oiy);
        tpz = dsubf(tz, oiz);
        fsx = cj.loc[0];
        fsy = cj.loc[1];
        fsz = cj.loc[2];
        ascending = true;
      } else {
        /* targets in cj: functions_hydro.h:1420-1448 */
        const double hj_max = (double)__fmul_rn(fminf(h_max_lim, cj.h_max_active), KERNEL_GAMMA);
        const int i1 = ci.first + (int)A.sort_idx[sort_offset(ci, sid) + ci.count - 1];
        const double di_max = __dsub_rn(
            (double)sort_key(A.x[3 * (size_t)i1], A.x[3 * (size_t)i1 + 1], A.x[3 * (size_t)i1 + 2], sid),
            rshift);
        const bool in_loop = __dsub_rn(__dsub_rn((double)tkey, hj_max), (double)dx_max) < di_max;
        const double dj = __dadd_rn((double)__fsub_rn(__fsub_rn(tkey, thg), dx_max), rshift);
        part = part && in_loop && !(__dsub_rn(dj, rshift) > di_max);
        thr = __double2float_rd(dj); /* key > dj  <=>  key > rd(dj) */
        tpx = dsubf(tx, cj.loc[0]);
        tpy = dsubf(ty, cj.loc[1]);
        tpz = dsubf(tz, cj.loc[2]);
        fsx = oix;
        fsy = oiy;
        fsz = oiz;
        ascending = false;
      }
    } else if (mode == MODE_SUB_SELF) {
      /* DOSELF_SUBSET :1127-1129: floats relative to c->loc (c = scell) */
      tpx = dsubf(tx, sc.loc[0]);
      tpy = dsubf(ty, sc.loc[1]);
      tpz = dsubf(tz, sc.loc[2]);
      fsx = sc.loc[0];
      fsy = sc.loc[1];
      fsz = sc.loc[2];
    } else if (mode == MODE_SUB_PAIR || mode == MODE_SUB_PAIR_F) {
      /* DOPAIR_SUBSET :885-897 / :955-967 */
      tdx = __dsub_rn(tx, shx);
      tdy = __dsub_rn(ty, shy);
      tdz = __dsub_rn(tz, shz);
      const double proj = __dadd_rn(
          __dadd_rn(__dmul_rn(tdx, c_runner_shift[sid][0]), __dmul_rn(tdy, c_runner_shift[sid][1])),
          __dmul_rn(tdz, c_runner_shift[sid][2]));
      /* di = hi*kernel_gamma + dxj + pix*rs0 + piy*rs1 + piz*rs2, left to right */
      double di;
      const float dxj = sc.dx_max_sort;
      if (mode == MODE_SUB_PAIR) {
        const float f0 = __fadd_rn(thg, dxj);
        di = __dadd_rn(__dadd_rn(__dadd_rn((double)f0, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                 __dmul_rn(tdy, c_runner_shift[sid][1])),
                       __dmul_rn(tdz, c_runner_shift[sid][2]));
        thr = __double2float_ru(di);
        ascending = true;
      } else {
        const float f0 = __fsub_rn(-thg, dxj);
        di = __dadd_rn(__dadd_rn(__dadd_rn((double)f0, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                 __dmul_rn(tdy, c_runner_shift[sid][1])),
                       __dmul_rn(tdz, c_runner_shift[sid][2]));
        thr = __double2float_rd(di);
        ascending = false;
      }
      (void)proj;
    }

    if (!__any_sync(FULL_MASK, part)) continue;
    const float my_hg2 = part ? thg2 : -1.f;
    /* reach of the warp along the axis, for the sorted early exit */
    float reach = 0.f;
    if (sorted) reach = ascending ? warp_max(part ? thr : -3.0e38f) : warp_min(part ? thr : 3.0e38f);

    for (int base = 0; base < scount; base += 32) {
      /* ---- stage 32 sources ---- */
      const int k = base + lane;
      int sj = -1;
      float skey = ascending ? 3.0e38f : -3.0e38f; /* padding never passes the prune */
      __syncwarp();
      if (k < scount) {
        int local = k;
        if (sorted) local = (int)A.sort_idx[soff + (ascending ? k : scount - 1 - k)];
        sj = sc.first + local;
        const double sx = A.x[3 * (size_t)sj], sy = A.x[3 * (size_t)sj + 1],
                     sz = A.x[3 * (size_t)sj + 2];
        if (sorted) skey = sort_key(sx, sy, sz, sid);
        if (dbl_mode) {
          T.xd[lane][0] = sx;
          T.xd[lane][1] = sy;
          T.xd[lane][2] = sz;
          T.pos[lane] = make_float4(0.f, 0.f, 0.f, skey);
        } else {
          T.pos[lane] = make_float4(dsubf(sx, fsx), dsubf(sy, fsy), dsubf(sz, fsz), skey);
        }
        T.f0[lane] = A.mv[sj];
        if (LOOP == LOOP_GRADIENT) {
          const float4 q1 = A.fq1[sj];
          const float4 q2 = A.fq2[sj];
          const float4 q3 = A.fq3[sj];
          T.f1[lane] = make_float4(q2.z /*u*/, q1.x /*rho*/, q1.w /*cs*/, q3.x /*alpha*/);
        }
      } else {
        T.pos[lane] = make_float4(1.0e30f, 1.0e30f, 1.0e30f, skey);
        if (dbl_mode) T.xd[lane][0] = T.xd[lane][1] = T.xd[lane][2] = 1.0e300;
      }
      T.idx[lane] = sj;
      __syncwarp();

      /* sorted early exit: first key of the chunk already out of everyone's reach */
      if (sorted) {
        const float first_key = T.pos[0].w;
        if (ascending ? !(first_key < reach) : !(first_key > reach)) break;
      }

      /* ---- test ---- */
      nchunks++;
      unsigned mask = 0u;
      if (!dbl_mode) {
#pragma unroll 8
        for (int q = 0; q < 32; q++) {
          const float4 s = T.pos[q];
          const float dx = __fsub_rn(tpx, s.x), dy = __fsub_rn(tpy, s.y), dz = __fsub_rn(tpz, s.z);
          const float r2 = r2_exact(dx, dy, dz);
          bool ok = r2 < my_hg2;
          if (mode == MODE_PAIR_L) ok = ok && (s.w < thr);
          if (mode == MODE_PAIR_R) ok = ok && (s.w > thr);
          if (mode == MODE_SUB_SELF) ok = ok && (T.idx[q] != ti);
          mask |= (ok ? 1u : 0u) << q;
        }
      } else {
#pragma unroll 4
        for (int q = 0; q < 32; q++) {
          const float dx = dsubf(tdx, T.xd[q][0]), dy = dsubf(tdy, T.xd[q][1]),
                      dz = dsubf(tdz, T.xd[q][2]);
          const float r2 = r2_exact(dx, dy, dz);
          bool ok = r2 < my_hg2;
          const float key = T.pos[q].w;
          if (mode == MODE_SELF) ok = ok && (T.idx[q] != ti) && (T.idx[q] >= 0);
          if (mode == MODE_SUB_PAIR) ok = ok && (key < thr);
          if (mode == MODE_SUB_PAIR_F) ok = ok && (key > thr);
          mask |= (ok ? 1u : 0u) << q;
        }
      }

      /* ---- drain ---- */
      while (__any_sync(FULL_MASK, mask != 0u)) {
        if (mask) {
          const int q = __ffs(mask) - 1;
          mask &= mask - 1u;
          float dx, dy, dz;
          if (!dbl_mode) {
            const float4 s = T.pos[q];
            dx = __fsub_rn(tpx, s.x);
            dy = __fsub_rn(tpy, s.y);
            dz = __fsub_rn(tpz, s.z);
          } else {
            dx = dsubf(tdx, T.xd[q][0]);
            dy = dsubf(tdy, T.xd[q][1]);
            dz = dsubf(tdz, T.xd[q][2]);
          }
          const float r2 = r2_exact(dx, dy, dz);
          const float4 f0 = T.f0[q];
          if (LOOP == LOOP_DENSITY) {
            iact_density(dacc, r2, dx, dy, dz, th_inv, tvx, tvy, tvz, f0.x, f0.y, f0.z, f0.w);
          } else {
            const float4 f1 = T.f1[q];
            iact_gradient(gacc, r2, dx, dy, dz, th, tvx, tvy, tvz, tu, tcs, f0.x, f0.y, f0.z, f0.w,
                          f1.x, f1.y, f1.z, f1.w, A.a2_Hubble);
          }
          nhit++;
        }
      }
    }
  }

  /* ---- flush (once per task) ---- */
  if (tvalid) {
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
    if (nhit) atomicAdd(&A.count[ti], nhit);
  }
  /* global interaction counter */
  int tot = nhit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && nchunks) atomicAdd(A.tests, (unsigned long long)nchunks * 1024ull);
}

/* ------------------------------------------------------------------------ */
/* Type-2 loop: force. Hit criterion r2 < max(h_t, h_s)^2 gamma^2, with the
 * two-pass pruning of DOPAIR2 restated per (i in ci, j in cj):
 *   pass A (:1737-1975): i in A-range, key_j < di_i, r2 < hig2
 *   pass B (:1978-2230): j in B-range, key_i - rshift > dj_j, hig2 <= r2 < hjg2 */
/* ------------------------------------------------------------------------ */
template <int SCHEME>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
    k_loop2(const LoopArgs A) {
  __shared__ SrcTile tiles[WARPS_PER_BLOCK];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int task = blockIdx.x * WARPS_PER_BLOCK + wib;
  if (task >= A.ntasks) return;
  SrcTile &T = tiles[wib];

  const int g = A.task_group[task];
  const int chunk = A.task_chunk[task];
  const int nt = A.tgt_count[g];
  if (chunk * 32 >= nt) return;
  const Group G = A.groups[g];
  const int slot = chunk * 32 + lane;
  const bool tvalid = slot < nt;
  const int ti = tvalid ? A.tgt_list[A.tgt_first[g] + slot] : -1;

  double tx = 0., ty = 0., tz = 0.;
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
    const float4 q0 = A.mv[ti], q1 = A.fq1[ti], q2 = A.fq2[ti];
    tq.m = q0.x; tq.vx = q0.y; tq.vy = q0.z; tq.vz = q0.w;
    tq.rho = q1.x; tq.P = q1.y; tq.f = q1.z; tq.cs = q1.w;
    tq.balsara = q2.x; tq.h = q2.y; tq.u = q2.z; tq.time_bin = __float_as_int(q2.w);
    if (SCHEME == SCH_SPHENIX) {
      const float4 q3 = A.fq3[ti];
      tq.alpha_visc = q3.x;
      tq.alpha_diff = q3.y;
    }
    tdepth = A.depth_h[ti];
  }
  const float th = tq.h;
  const float thg2 = hg2_exact(th);
  const float thg = __fmul_rn(th, KERNEL_GAMMA);

  ForceAcc acc;
  acc.ax = acc.ay = acc.az = acc.u_dt = acc.h_dt = 0.f;
  acc.v_sig = 0.f;
  acc.min_ngb = NUM_TIME_BINS + 1;
  int nhit = 0;
  int nchunks = 0;

  for (int it = 0; it < G.item_count; it++) {
    const Item I = A.items[G.item_first + it];
    const DevCell sc = A.cells[I.scell];
    const int mode = I.mode;
    const int sid = I.sid;
    const int scount = sc.count;
    const bool part = tvalid && tdepth >= I.min_depth && tdepth <= I.max_depth;
    if (!__any_sync(FULL_MASK, part)) continue;
    const float my_hg2 = part ? thg2 : -1.f;

    if (mode == MODE_SELF) {
      /* DOSELF2 :2624-2875: doi = r2 < hig2 || r2 < hjg2 */
      for (int base = 0; base < scount; base += 32) {
        const int k = base + lane;
        __syncwarp();
        int sj = -1;
        if (k < scount) {
          sj = sc.first + k;
          T.xd[lane][0] = A.x[3 * (size_t)sj];
          T.xd[lane][1] = A.x[3 * (size_t)sj + 1];
          T.xd[lane][2] = A.x[3 * (size_t)sj + 2];
          const float4 q2 = A.fq2[sj];
          T.f0[lane] = A.mv[sj];
          T.f1[lane] = A.fq1[sj];
          T.f2[lane] = q2;
          if (SCHEME == SCH_SPHENIX) T.f3[lane] = A.fq3[sj];
          T.pos[lane] = make_float4(0.f, 0.f, 0.f, hg2_exact(q2.y));
        } else {
          T.xd[lane][0] = T.xd[lane][1] = T.xd[lane][2] = 1.0e300;
          T.pos[lane] = make_float4(0.f, 0.f, 0.f, -1.f);
        }
        T.idx[lane] = sj;
        __syncwarp();
        nchunks++;
        unsigned mask = 0u;
#pragma unroll 4
        for (int q = 0; q < 32; q++) {
          const float dx = dsubf(tx, T.xd[q][0]), dy = dsubf(ty, T.xd[q][1]),
                      dz = dsubf(tz, T.xd[q][2]);
          const float r2 = r2_exact(dx, dy, dz);
          const float shg2 = T.pos[q].w;
          const bool ok = part && (r2 < my_hg2 || r2 < shg2) && (T.idx[q] != ti) && (T.idx[q] >= 0);
          mask |= (ok ? 1u : 0u) << q;
        }
        while (__any_sync(FULL_MASK, mask != 0u)) {
          if (mask) {
            const int q = __ffs(mask) - 1;
            mask &= mask - 1u;
            const float dx = dsubf(tx, T.xd[q][0]), dy = dsubf(ty, T.xd[q][1]),
                        dz = dsubf(tz, T.xd[q][2]);
            const float r2 = r2_exact(dx, dy, dz);
            ForceQ sq;
            const float4 q0 = T.f0[q], q1 = T.f1[q], q2 = T.f2[q];
            sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
            sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
            sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
            sq.alpha_visc = sq.alpha_diff = 0.f;
            if (SCHEME == SCH_SPHENIX) {
              const float4 q3 = T.f3[q];
              sq.alpha_visc = q3.x;
              sq.alpha_diff = q3.y;
            }
            iact_force<SCHEME>(acc, r2, dx, dy, dz, tq, sq, A.a2_Hubble);
            nhit++;
          }
        }
      }
      continue;
    }

    /* ---- DOPAIR2 ---- */
    const DevCell tc = A.cells[I.tcell];
    const bool tleft = (mode == MODE_PAIR_L);
    const DevCell &ci = tleft ? tc : sc;
    const DevCell &cj = tleft ? sc : tc;
    const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1],
                 shz = I.shift[2] * A.dim[2];
    const double rshift = __dadd_rn(
        __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
        __dmul_rn(shz, c_runner_shift[sid][2]));
    const double hi_max_g = __dmul_rn((double)ci.h_max, (double)KERNEL_GAMMA);
    const double hj_max_g = __dmul_rn((double)cj.h_max, (double)KERNEL_GAMMA);
    const double dx_max = (double)__fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
    const int64_t soff_i = sort_offset(ci, sid), soff_j = sort_offset(cj, sid);
    const int i1 = ci.first + (int)A.sort_idx[soff_i + ci.count - 1];
    const int j0 = cj.first + (int)A.sort_idx[soff_j];
    const double di_max =
        (double)sort_key(A.x[3 * (size_t)i1], A.x[3 * (size_t)i1 + 1], A.x[3 * (size_t)i1 + 2], sid);
    const double dj_min =
        (double)sort_key(A.x[3 * (size_t)j0], A.x[3 * (size_t)j0 + 1], A.x[3 * (size_t)j0 + 2], sid);
    const double di_max_sh = __dsub_rn(di_max, rshift);
    const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                 oiz = __dadd_rn(cj.loc[2], shz);

    /* Per-particle pruning values. For a particle p of ci:
     *   okA, di(p) = (float)(key + h*gamma) + dx_max - rshift, keyA(p) = key - rshift
     * for a particle p of cj:
     *   okB, dj(p) = (float)(key - h*gamma) - dx_max                                  */
    const float tkey = sort_key(tx, ty, tz, sid);
    double t_di = -1.0e300, t_keysh = 0., t_dj = 1.0e300;
    float tpx, tpy, tpz;
    if (tleft) {
      const bool inA =
          __dsub_rn(__dadd_rn(__dadd_rn((double)tkey, hi_max_g), dx_max), rshift) > dj_min;
      const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(tkey, thg), dx_max), rshift);
      if (inA && !(di < dj_min)) t_di = di;
      t_keysh = __dsub_rn((double)tkey, rshift);
      tpx = dsubf(tx, oix);
      tpy = dsubf(ty, oiy);
      tpz = dsubf(tz, oiz);
    } else {
      const bool inB = __dsub_rn(__dsub_rn((double)tkey, hj_max_g), dx_max) < di_max_sh;
      const double dj = __dsub_rn((double)__fsub_rn(tkey, thg), dx_max);
      if (inB && !(dj > di_max_sh)) t_dj = dj;
      tpx = dsubf(tx, cj.loc[0]);
      tpy = dsubf(ty, cj.loc[1]);
      tpz = dsubf(tz, cj.loc[2]);
    }
    /* Conservative reach along the axis for the sorted early exit. */
    double reach;
    if (tleft) {
      /* sources j ascending; j can matter while key_j - hj_max*g - dx_max <= max(di, keysh) */
      reach = warp_max_d(part ? fmax(t_di, t_keysh) : -1.0e300);
    } else {
      /* sources i descending; i can matter while key_i + hi_max*g + dx_max - rshift >= min(key_t, dj) */
      reach = warp_min_d(part ? fmin((double)tkey, t_dj) : 1.0e300);
    }
    const int64_t soff = tleft ? soff_j : soff_i;

    for (int base = 0; base < scount; base += 32) {
      const int k = base + lane;
      __syncwarp();
      int sj = -1;
      float skey = tleft ? 3.0e38f : -3.0e38f;
      if (k < scount) {
        const int local = (int)A.sort_idx[soff + (tleft ? k : scount - 1 - k)];
        sj = sc.first + local;
        const double sx = A.x[3 * (size_t)sj], sy = A.x[3 * (size_t)sj + 1],
                     sz = A.x[3 * (size_t)sj + 2];
        skey = sort_key(sx, sy, sz, sid);
        const float4 q2 = A.fq2[sj];
        const float sh = q2.y;
        const float shg = __fmul_rn(sh, KERNEL_GAMMA);
        if (tleft) {
          /* source j in cj */
          const bool inB = __dsub_rn(__dsub_rn((double)skey, hj_max_g), dx_max) < di_max_sh;
          const double dj = __dsub_rn((double)__fsub_rn(skey, shg), dx_max);
          T.dB[lane] = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
          T.pos[lane] = make_float4(dsubf(sx, cj.loc[0]), dsubf(sy, cj.loc[1]), dsubf(sz, cj.loc[2]), skey);
        } else {
          /* source i in ci */
          const bool inA =
              __dsub_rn(__dadd_rn(__dadd_rn((double)skey, hi_max_g), dx_max), rshift) > dj_min;
          const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(skey, shg), dx_max), rshift);
          T.dA[lane] = (inA && !(di < dj_min)) ? di : -1.0e300;
          T.dB[lane] = __dsub_rn((double)skey, rshift); /* keyA of the source */
          T.pos[lane] = make_float4(dsubf(sx, oix), dsubf(sy, oiy), dsubf(sz, oiz), skey);
        }
        T.f0[lane] = A.mv[sj];
        T.f1[lane] = A.fq1[sj];
        T.f2[lane] = q2;
        if (SCHEME == SCH_SPHENIX) T.f3[lane] = A.fq3[sj];
      } else {
        T.pos[lane] = make_float4(1.0e30f, 1.0e30f, 1.0e30f, skey);
        T.dB[lane] = tleft ? 1.0e300 : -1.0e300;
        T.dA[lane] = -1.0e300;
        T.f2[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      T.idx[lane] = sj;
      __syncwarp();

      /* early exit on the sorted axis (conservative, with a rounding slack) */
      {
        const double fk = (double)T.pos[0].w;
        const double slack = 1.0e-5 * (fabs(fk) + 1.0);
        if (tleft) {
          if (fk - hj_max_g - dx_max - slack > reach) break;
        } else {
          if (fk + hi_max_g + dx_max - rshift + slack < reach) break;
        }
      }

      nchunks++;
      unsigned mask = 0u;
#pragma unroll 4
      for (int q = 0; q < 32; q++) {
        const float4 s = T.pos[q];
        const float dx = __fsub_rn(tpx, s.x), dy = __fsub_rn(tpy, s.y), dz = __fsub_rn(tpz, s.z);
        const float r2 = r2_exact(dx, dy, dz);
        const float shg2 = hg2_exact(T.f2[q].y);
        bool ok;
        if (tleft) {
          /* t = i in ci, s = j in cj */
          const bool c1 = ((double)s.w < t_di) && (r2 < thg2);
          const bool c2 = (t_keysh > T.dB[q]) && (r2 < shg2) && !(r2 < thg2);
          ok = c1 || c2;
        } else {
          /* t = j in cj, s = i in ci */
          const bool c1 = ((double)tkey < T.dA[q]) && (r2 < shg2);
          const bool c2 = (T.dB[q] > t_dj) && (r2 < thg2) && !(r2 < shg2);
          ok = c1 || c2;
        }
        ok = ok && part && (T.idx[q] >= 0);
        mask |= (ok ? 1u : 0u) << q;
      }
      while (__any_sync(FULL_MASK, mask != 0u)) {
        if (mask) {
          const int q = __ffs(mask) - 1;
          mask &= mask - 1u;
          const float4 s = T.pos[q];
          const float dx = __fsub_rn(tpx, s.x), dy = __fsub_rn(tpy, s.y), dz = __fsub_rn(tpz, s.z);
          const float r2 = r2_exact(dx, dy, dz);
          ForceQ sq;
          const float4 q0 = T.f0[q], q1 = T.f1[q], q2 = T.f2[q];
          sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
          sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
          sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
          sq.alpha_visc = sq.alpha_diff = 0.f;
          if (SCHEME == SCH_SPHENIX) {
            const float4 q3 = T.f3[q];
            sq.alpha_visc = q3.x;
            sq.alpha_diff = q3.y;
          }
          iact_force<SCHEME>(acc, r2, dx, dy, dz, tq, sq, A.a2_Hubble);
          nhit++;
        }
      }
    }
  }

  if (tvalid) {
    float *po = (float *)&A.fo1[ti];
    atomicAdd(po + 0, acc.ax);
    atomicAdd(po + 1, acc.ay);
    atomicAdd(po + 2, acc.az);
    atomicAdd(po + 3, acc.u_dt);
    atomicAdd(&A.f_hdt[ti], acc.h_dt);
    if (SCHEME != SCH_SPHENIX) atomic_max_pos(&A.f_vsig[ti], acc.v_sig);
    atomicMin(&A.f_minngb[ti], acc.min_ngb);
    if (nhit) atomicAdd(&A.count[ti], nhit);
  }
  int tot = nhit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
  if (lane == 0 && tot) atomicAdd(A.total, (unsigned long long)tot);
  if (lane == 0 && nchunks) atomicAdd(A.tests, (unsigned long long)nchunks * 1024ull);
}

}  // namespace swiftgpu
#endif
