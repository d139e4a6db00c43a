/*
 * loops_common.cuh - what every generation of the neighbour-loop kernels shares:
 * the device view of a cell, the argument block of a loop launch, warp helpers,
 * the mbarrier / bulk-TMA PTX wrappers, and the EXACT sorted-axis conditions of
 * DOPAIR1 / DOPAIR2 / DOPAIR_SUBSET (functions_hydro.h:1296-1332, :1420-1448,
 * :1652-1735, :1806-2238, :891-1000), evaluated only for pairs within keyE of
 * the cut-off (see loops_pipe.cuh).
 */
#ifndef SWIFTGPU_LOOPS_COMMON_CUH
#define SWIFTGPU_LOOPS_COMMON_CUH

#include "sph_math.cuh"
#include "worklist.hpp"

namespace swiftgpu {

#define FULL_MASK 0xffffffffu

/* Device view of one cell (80 bytes). */
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
  float width;   /* max_k width[k] */
  float dx_max_part; /* how far a particle may sit outside the cell box */
  int32_t seg_base;  /* index of this cell's first (cell, sid) segment, see seg_index() */
};
static_assert(sizeof(DevCell) == 80, "DevCell layout");

/* Index of the (cell, sid) segment: key extrema and sorted index arrays are
 * stored per requested segment, a cell's segments are consecutive. */
__device__ __forceinline__ int seg_index(const DevCell &c, int sid) {
  return c.seg_base + __popc((unsigned)c.sort_mask & ((1u << sid) - 1u));
}
__device__ __forceinline__ int64_t sort_offset(const DevCell &c, int sid) {
  return c.sort_base + (int64_t)__popc((unsigned)c.sort_mask & ((1u << sid) - 1u)) * c.count;
}

struct TaskRec;
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
  const float2 *ext; /* per (cell, sid) segment: (min, max) sort key = sort[0].d, sort[count-1].d */
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
  /* tile pipeline (loops_tile.cuh): TMA-copyable source records */
  const float4 *xf;   /* (float x, y, z of the absolute position, (h gamma REL + margin)^2) */
  const double *xs0, *xs1, *xs2; /* SoA copies of the double positions (8-byte TMA columns) */
  const float4 *gq;   /* gradient payload (u, rho, cs, alpha_visc) */
  const float4 *boxes; /* per cell octet: lo.xyz_, hi.xyz_ */
  const int32_t *cell_box_first;
  float keyE;   /* r-margin below which the sorted-axis conditions are implied */
  float margin; /* absolute widening of the float prefilter */
  int hold;     /* stages a consumer warp holds before it drains (<= NS - 1) */
  unsigned int *task_counter; /* persistent CTAs draw their tasks from here (zeroed per launch) */
  /* frame pipeline (loops_pipe.cuh) */
  const float4 *frames;            /* per-(cell, origin) arrays of (float)(x - origin) */
  const struct TaskRec *task_recs; /* compacted tasks of this launch (k_task_recs) */
  const unsigned int *ntask_dev;   /* their number */
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
  int32_t *wakeup; /* limiter loop: limiter_data.wakeup of every particle (scatter, atomicMax) */
  int32_t *count; /* per-particle directed interaction counter of this loop */
  unsigned long long *total; /* global interaction counter */
  unsigned long long *tests; /* global distance-test counter */
  double dim[3];
  float a2_Hubble;
  int max_active_bin;
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

/* Relative inflation of the prefilter radius^2: covers the difference between
 * the fused r2 of the prefilter and the reference's un-fused r2 (~2^-22). */
#define PREFILTER_REL 1.00001f
#define TASK_TARGETS 64 /* targets of one chunk of the host task list */

#define TL_CWARPS 8 /* consumer warps of the standard CTA (template CW: 8, or 4 for sparse target sets) */
#define TL_TARGETS 64 /* targets of a task chunk (host task list); a CW-warp CTA takes 8 * CW of them */
#ifndef TL_SLOTS
#define TL_SLOTS 256 /* source slots per stage (<= 256: 8-bit slot field of the list entries) */
#endif
#define TL_OCT (TL_SLOTS / 8)
#define TL_FRAGS 8   /* fragments per stage */
#ifndef TL_WAIT_HINT_NS
#define TL_WAIT_HINT_NS 2000u /* try_wait suspend-time hint */
#endif
#define TL_SUBCAP1 10 /* sub-list capacity per lane, type-1 loops (3 CTAs/SM) */
#define TL_SUBCAP2 16 /* ... force loop (2 CTAs/SM) */
#ifndef TL_DENS_BLOCKS
#define TL_DENS_BLOCKS 3 /* resident CTAs per SM the type-1 kernels are compiled for */
#endif
#ifndef TL_DENS_NS
#define TL_DENS_NS 4
#endif
#ifndef TL_SPARSE_NS
#define TL_SPARSE_NS 2 /* ring stages of the 4-warp CTAs */
#define TL_SPARSE_BLOCKS 5
#endif
#define TL_DCOL (TL_SLOTS + 2 * TL_FRAGS) /* double column: 2 spare entries per fragment (alignment) */

/* ---- mbarrier / bulk-TMA PTX ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  /* the last operand lets the hardware keep the warp suspended (no issue slots) for up to ~2 us */
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity), "r"(TL_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
/* Bounded wait: a protocol error traps (after 4 s of wall-clock time on the
 * device) instead of hanging the GPU. */
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  if (mbar_try(b, parity)) return;
  unsigned long long t0 = 0;
  for (unsigned spins = 1; !mbar_try(b, parity); spins++) {
    if ((spins & 1023u) == 0) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      if (t - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}


/* ---- exact sorted-axis conditions (rare path) ---- */

/* Type-1 loops, functions_hydro.h:1296-1332 / :1420-1448 (DOPAIR1) and
 * :891-1000 (DOPAIR_SUBSET): does the reference reach source s from target t
 * along the sorted axis? Same arithmetic as the reference, constants
 * re-derived from the cells. */
struct SlowArgs {
  const Item *items;
  const DevCell *cells;
  const float2 *ext;
  double dim[3];
};
__device__ __noinline__ bool exact_type1(const SlowArgs A, int item, double tx, double ty, double tz,
                                         float thg, double sx, double sy, double sz) {
  const Item I = A.items[item];
  const DevCell tc = A.cells[I.tcell];
  const DevCell sc = A.cells[I.scell];
  const int mode = I.mode, sid = I.sid;
  const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
  const float skey = sort_key(sx, sy, sz, sid);
  if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
    const DevCell &ci = (mode == MODE_PAIR_L) ? tc : sc;
    const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tc;
    const double rshift = __dadd_rn(
        __dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
        __dmul_rn(shz, c_runner_shift[sid][2]));
    const double dj_min = (double)A.ext[seg_index(cj, sid)].x;
    const double di_max = (double)A.ext[seg_index(ci, sid)].y;
    const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
    const float dx_max = __fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
    const float tkey = sort_key(tx, ty, tz, sid);
    if (mode == MODE_PAIR_L) {
      const double lim_a =
          __dsub_rn((double)__fmul_rn(fminf(h_max_lim, ci.h_max_active), KERNEL_GAMMA), rshift);
      const bool in_loop = __dadd_rn(__dadd_rn((double)tkey, lim_a), (double)dx_max) > dj_min;
      const double di = __dsub_rn((double)__fadd_rn(__fadd_rn(tkey, thg), dx_max), rshift);
      return in_loop && !(di < dj_min) && ((double)skey < di);
    } else {
      const double lim_a = (double)__fmul_rn(fminf(h_max_lim, cj.h_max_active), KERNEL_GAMMA);
      const double lim_b = __dsub_rn(di_max, rshift);
      const bool in_loop = __dsub_rn(__dsub_rn((double)tkey, lim_a), (double)dx_max) < lim_b;
      const double dj = __dadd_rn((double)__fsub_rn(__fsub_rn(tkey, thg), dx_max), rshift);
      return in_loop && !(__dsub_rn(dj, rshift) > lim_b) && ((double)skey > dj);
    }
  }
  if (mode == MODE_SUB_PAIR || mode == MODE_SUB_PAIR_F) {
    const double tdx = __dsub_rn(tx, shx), tdy = __dsub_rn(ty, shy), tdz = __dsub_rn(tz, shz);
    const float dx_max = sc.dx_max_sort;
    const float f0 = (mode == MODE_SUB_PAIR) ? __fadd_rn(thg, dx_max) : __fsub_rn(-thg, dx_max);
    const double di = __dadd_rn(__dadd_rn(__dadd_rn((double)f0, __dmul_rn(tdx, c_runner_shift[sid][0])),
                                          __dmul_rn(tdy, c_runner_shift[sid][1])),
                                __dmul_rn(tdz, c_runner_shift[sid][2]));
    return (mode == MODE_SUB_PAIR) ? ((double)skey < di) : ((double)skey > di);
  }
  return true;
}

/* Type-2 loop, DOPAIR2 functions_hydro.h:1652-1735 (ranges) and :1806-2238
 * (the two passes): is the pair taken, given the exact r2? */
__device__ __noinline__ bool exact_type2(const SlowArgs A, int item, double tx, double ty, double tz,
                                         float thg, float thg2, double sx, double sy, double sz,
                                         float shg, float shg2, float r2) {
  const Item I = A.items[item];
  const DevCell tc = A.cells[I.tcell];
  const DevCell sc = A.cells[I.scell];
  const int mode = I.mode, sid = I.sid;
  const bool tleft = (mode == MODE_PAIR_L);
  const DevCell &ci = tleft ? tc : sc;
  const DevCell &cj = tleft ? sc : tc;
  const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
  const double rshift =
      __dadd_rn(__dadd_rn(__dmul_rn(shx, c_runner_shift[sid][0]), __dmul_rn(shy, c_runner_shift[sid][1])),
                __dmul_rn(shz, c_runner_shift[sid][2]));
  const double dj_min = (double)A.ext[seg_index(cj, sid)].x;
  const double di_max = (double)A.ext[seg_index(ci, sid)].y;
  const double hi_max_g = __dmul_rn((double)ci.h_max, (double)KERNEL_GAMMA);
  const double hj_max_g = __dmul_rn((double)cj.h_max, (double)KERNEL_GAMMA);
  const double dx_max = (double)__fadd_rn(ci.dx_max_sort, cj.dx_max_sort);
  const double di_max_sh = __dsub_rn(di_max, rshift);
  const float tkey = sort_key(tx, ty, tz, sid);
  const float skey = sort_key(sx, sy, sz, sid);
  if (tleft) {
    const bool inA = __dsub_rn(__dadd_rn(__dadd_rn((double)tkey, hi_max_g), dx_max), rshift) > dj_min;
    const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(tkey, thg), dx_max), rshift);
    const double t_di = (inA && !(di < dj_min)) ? di : -1.0e300;
    const double t_keysh = __dsub_rn((double)tkey, rshift);
    const bool inB = __dsub_rn(__dsub_rn((double)skey, hj_max_g), dx_max) < di_max_sh;
    const double dj = __dsub_rn((double)__fsub_rn(skey, shg), dx_max);
    const double s_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
    const bool c1 = ((double)skey < t_di) && (r2 < thg2);
    const bool c2 = (t_keysh > s_dj) && (r2 < shg2) && !(r2 < thg2);
    return c1 || c2;
  } else {
    const bool inB = __dsub_rn(__dsub_rn((double)tkey, hj_max_g), dx_max) < di_max_sh;
    const double dj = __dsub_rn((double)__fsub_rn(tkey, thg), dx_max);
    const double t_dj = (inB && !(dj > di_max_sh)) ? dj : 1.0e300;
    const bool inA = __dsub_rn(__dadd_rn(__dadd_rn((double)skey, hi_max_g), dx_max), rshift) > dj_min;
    const double di = __dsub_rn(__dadd_rn((double)__fadd_rn(skey, shg), dx_max), rshift);
    const double s_di = (inA && !(di < dj_min)) ? di : -1.0e300;
    const double s_keysh = __dsub_rn((double)skey, rshift);
    const bool c1 = ((double)tkey < s_di) && (r2 < shg2);
    const bool c2 = (s_keysh > t_dj) && (r2 < thg2) && !(r2 < shg2);
    return c1 || c2;
  }
}

/* (max(a - E, 0))^2 shaved by 2^-19: below it, r < a - E certainly. */
__device__ __forceinline__ float sure_r2(float a, float E) {
  const float b = a - E;
  return b > 0.f ? b * b * 0.999998f : 0.f;
}


}  // namespace swiftgpu
#endif
