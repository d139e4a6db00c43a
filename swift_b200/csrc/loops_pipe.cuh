/*
 * loops_pipe.cuh - the neighbour loops (density, ghost re-runs, gradient,
 * force, time-step limiter) as a TMA-fed producer/consumer pipeline over
 * PRECOMPUTED FRAME FLOATS.
 *
 * One persistent CTA = CW consumer warps (<= 8 TARGET particles each; a full
 * task: one Morton octet of a leaf per warp) + 1 producer warp; tasks (<= 8 CW
 * targets of one target cell, dealt evenly to the warps) are drawn from a
 * global counter and the ring of NS stages runs across task boundaries. CW = 8
 * (64-target tasks) or, for sparse target sets, 4 (32-target tasks, more CTAs
 * per SM); which of the two finds tasks is decided on the device (k_task_recs).
 * One __syncthreads() in the prologue, none after it.
 *
 * What the reference evaluates for a pair is decided in the float frame of the
 * leaf-level call, functions_hydro.h:1327-1347:
 *     pix = (float)(pi->x - (cj->loc + shift)),  pjx = (float)(pj->x - cj->loc),
 *     dx = pix - pjx,  r2 = dx*dx + dy*dy + dz*dz (un-fused),  r2 < hig2
 * Both frame floats depend on ONE particle and ONE (cell, origin) combination
 * only, so they are computed once per step by k_frames (kernels_records.cuh) into
 * per-(cell, origin) arrays of float4 - 14 per leaf of a uniform tree: the
 * cell's own frame plus the 13 directions in which it is the left cell `ci` -
 * and the loops stage exactly the array an item needs with one bulk TMA copy.
 * The distance test then IS the reference's arithmetic: no conservative
 * prefilter, no doubles in the loop, every listed candidate is a hit. Only the
 * items the reference evaluates on doubles (DOSELF1/2 :2299,:2624:
 * (float)(x_i - x_j); DOPAIR_SUBSET :891-916: (float)((x_i - shift) - x_j)) keep
 * the two-step form: a prefilter on the source cell's own-frame floats, then
 * the exact expression from staged double columns in the drain.
 *
 *   PRODUCER  per task: one 48-byte TaskRec (targets, their box and reach,
 *             built on the device by k_task_recs), then lane = item: the
 *             constants of 32 items of the target cell's group (frame origins,
 *             cull offset, caps), item-level cull against the task's target
 *             box, fragments of <= 256 source slots / 8 items per stage. The
 *             source data are copied as they lie in HBM by bulk TMA
 *             (cp.async.bulk -> mbarrier complete_tx): the frame array, the
 *             payload columns (mv | gq | fq1, fq2, fq3), the octet boxes and,
 *             for double-mode fragments only, the three double columns.
 *             While the stages of task k stream, the producer PREFETCHES task
 *             k+1 one dependent load per stage (task id -> TaskRec -> items ->
 *             source cells -> window), so that no chain of global-load
 *             latencies is exposed at a task boundary.
 *   CONSUMER  per stage: lane = octet box-box cull against the warp's target
 *             box (one ballot); for every fragment with an accepted octet the
 *             target's coordinates in that item's frame are computed ONCE
 *             ((float)(x_t - origin), 3 DADD + 3 F2F) and parked in shared
 *             memory together with the test limit; per accepted octet every
 *             lane (t = lane & 7 target, s = lane >> 3) tests its target
 *             against sources 2s, 2s+1 and appends hits to its sub-list. A
 *             drain merges the 4 sub-lists of a target over its 4 lanes,
 *             re-forms dx from the parked target floats and the staged source
 *             floats (3 FADD) and applies the interaction; then the held
 *             stages are released.
 *
 * The sorted-axis conditions of DOPAIR1/DOPAIR2 (:1296-1332, :1420-1448,
 * :1652-1735, :1806-2238) are geometrically implied by the distance condition
 * up to the rounding of the float sort keys; exact_type1/2 (loops_common.cuh)
 * evaluate them for hits within keyE of the cut-off only.
 */
#ifndef SWIFTGPU_LOOPS_PIPE_CUH
#define SWIFTGPU_LOOPS_PIPE_CUH

#include "loops_common.cuh"

namespace swiftgpu {

#ifndef PL_SLOTS
#define PL_SLOTS 256 /* source slots per stage (8-bit slot field of the list entries); multiple of 8, <= 256 */
#endif
#define PL_OCT (PL_SLOTS / 8)
#define PL_FRAGS 8 /* fragments (items) per stage */
#define TIME_BIN_NEIGHBOUR_MAX_DELTA_BIN 2 /* timeline.h:51 */
#define PL_TARGETS 64 /* targets of a task of the 8-warp kernel (8 per consumer warp) */
#ifndef PL_SPARSE_CW
#define PL_SPARSE_CW 4 /* consumer warps of the variant for sparse target sets: tasks of 8 * PL_SPARSE_CW targets */
#endif
#define PL_PRE_REL2 1.00002f /* relative widening of a prefilter limit r^2 (double modes) */

/* One task = up to 64 targets of one group. Built on the device after the
 * target lists (k_task_recs), compacted: empty tasks do not appear. */
struct __align__(16) TaskRec {
  int32_t item_first, item_count;
  int32_t tgt_off; /* offset of the task's first target in tgt_list */
  int32_t ntgt;
  int32_t tcell;
  float lo[3], hi[3]; /* box of the targets in the target cell's own frame: (float)(x - loc) */
  float rmax;         /* max h gamma of the targets */
};
static_assert(sizeof(TaskRec) == 48, "TaskRec");

/* Constants of one item, derived by the producer (lane = item) and handed to
 * the consumers with every fragment of the item. */
struct __align__(16) PipeItem {
  double ot[3]; /* test frame of the target: tp = (float)(x_t - ot) */
  double sh[3]; /* double modes: dx = (float)((x_t - sh) - x_s) */
  float d[3];   /* cull: target own-frame float - d = position in the source cell's own frame */
  float rsrc;   /* force: h_max gamma of the source cell (also the source-side cap) */
  float relq, padd; /* (8-byte aligned) force test: r2 < fma(max(hig2, hjg2), relq, padd); (1, 0) in the frame modes */
  float hcap;   /* target-side cap of h gamma under which the key conditions are implied */
  int32_t item;     /* global item index (slow path) */
  int32_t gi_base;  /* global particle index = gi_base + slot-in-stage */
  int16_t dofs;     /* slot -> index in the staged double columns */
  int8_t mode, sid, min_depth, max_depth, dbl, nokey;
};
static_assert(sizeof(PipeItem) == 96, "PipeItem");
static_assert(offsetof(PipeItem, relq) % 8 == 0, "relq/padd are read as one float2");

template <int NP, int NS, int QCAP, int CW, int DS, int SL = PL_SLOTS>
struct PipeSmem {
  static constexpr int kDCol = DS + 2 * PL_FRAGS; /* double column: 2 spare entries per fragment (alignment) */
  static constexpr int kStageF = 0;
  static constexpr int kStageP = kStageF + SL * 16;
  static constexpr int kStageD = kStageP + NP * SL * 16;
  static constexpr int kStageOB = kStageD + 3 * kDCol * 8;
  static constexpr int kStageIT = kStageOB + (SL / 8) * 32;
  static constexpr int kStageO2F = kStageIT + PL_FRAGS * (int)sizeof(PipeItem);
  static constexpr int kStageMeta = kStageO2F + (SL / 8);
  static constexpr int kStageBytes = ((kStageMeta + 64) + 127) & ~127;
  static constexpr int kList = NS * kStageBytes;
  static constexpr int kTP = kList + QCAP * 32 * CW * 2; /* float4 [CW][2][PL_FRAGS][8] */
  static constexpr int kBar = kTP + CW * 2 * PL_FRAGS * 8 * 16;
  static constexpr int kBox = kBar + 2 * NS * 8;
  static constexpr int kWin = kBox + (CW + 1) * 32; /* producer: 2 windows of 32 items */
  static constexpr int kWinAux = kWin + 2 * 32 * (int)sizeof(PipeItem);
  static constexpr int kBytes = kWinAux + 2 * 32 * 16;
};

/* stage meta: 16 words */
enum { PM_NFR = 0, PM_NOCT = 1, PM_FLAG = 2, PM_TASK = 3, PM_TGT_OFF = 4, PM_NTGT = 5, PM_LOC = 8 /* 3 doubles */ };

/* force payload lane of the exact hj^2 gamma^2 of the source (k_ghost / k_aos_to_soa keep it there) */
#define PL_HG2_COL(SCHEME) ((SCHEME) == SCH_SPHENIX ? 3 : 2)

#ifndef PL_BLOCKS_TYPE1
#define PL_BLOCKS_TYPE1 2 /* CTAs per SM the 8-warp density / gradient kernels are compiled for (3: 72 registers) */
#endif
#define PL_MIN_BLOCKS(LOOP, CW) ((CW) >= 8 ? ((LOOP) == LOOP_FORCE ? 2 : PL_BLOCKS_TYPE1) : 4)

template <int LOOP, int SCHEME, int NS, int CW, int DS, int SL = PL_SLOTS>
__global__ void __launch_bounds__(32 * (CW + 1), PL_MIN_BLOCKS(LOOP, CW)) k_pipe(const LoopArgs A) {
  static_assert(SL % 8 == 0 && SL <= 256 && SL >= 64, "stage slots: whole octets, 8-bit slot field");
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  constexpr int QCAP = FORCE ? TL_SUBCAP2 : TL_SUBCAP1;
  typedef PipeSmem<NP, NS, QCAP, CW, DS, SL> SM;
  static_assert(CW >= 1 && CW <= 8, "a task is 8 * CW targets");
  extern __shared__ __align__(128) char smem_pl[];
  char *const smem = smem_pl;
  uint16_t *const sList = (uint16_t *)(smem + SM::kList);
  uint64_t *const sFull = (uint64_t *)(smem + SM::kBar);
  uint64_t *const sEmpty = sFull + NS;
  float *const sBox = (float *)(smem + SM::kBox);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int t8 = lane & 7;
  const int s4 = lane >> 3;
  const bool consumer = warp < CW;

  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(sFull + s, 1);
      mbar_init(sEmpty + s, CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads(); /* barriers initialised: the only CTA-wide barrier */

  /* ===================================================================== */
  /* PRODUCER                                                               */
  /* ===================================================================== */
  if (!consumer) {
    PipeItem *const sWin = (PipeItem *)(smem + SM::kWin);
    int4 *const sWinAux = (int4 *)(smem + SM::kWinAux); /* (first particle, count, frame offset, first octet box) */
    const int ntask = (int)*A.ntask_dev;

    /* ---- one window of 32 items of a task: derive the constants, cull, write (lane = item) ---- */
    struct CellLite { /* what the item constants need of a cell */
      double loc[3];
      float h_max, h_max_active, h_max_allowed, width, dx_max_part;
      int32_t first, count;
    };
    auto load_cell = [&](int c) {
      const DevCell &C = A.cells[c];
      CellLite L;
      L.loc[0] = C.loc[0]; L.loc[1] = C.loc[1]; L.loc[2] = C.loc[2];
      L.h_max = C.h_max; L.h_max_active = C.h_max_active; L.h_max_allowed = C.h_max_allowed;
      L.width = C.width; L.dx_max_part = C.dx_max_part;
      L.first = C.first; L.count = C.count;
      return L;
    };
    struct WinLoad { /* registers in flight between the steps of the prefetch */
      Item I;
      CellLite sc;
      int bfirst;
      bool valid;
    };
    auto win_load_items = [&](const TaskRec &T, int win_base, WinLoad &W) {
      W.valid = win_base + lane < T.item_count;
      if (W.valid) W.I = A.items[T.item_first + win_base + lane];
    };
    auto win_load_cells = [&](WinLoad &W) {
      if (W.valid) {
        W.sc = load_cell(W.I.scell);
        W.bfirst = A.cell_box_first[W.I.scell];
      }
    };
    auto win_finish = [&](const TaskRec &T, const CellLite &tcell, int win_base, const WinLoad &W, int buf) -> unsigned {
      bool keep = false;
      if (W.valid) {
        const Item &I = W.I;
        const CellLite &sc = W.sc;
        const int mode = I.mode;
        const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
        /* written in place (a local PipeItem would live in local memory) */
        PipeItem *const wp = &sWin[buf * 32 + lane];
        wp->mode = (int8_t)mode;
        wp->sid = (int8_t)I.sid;
        wp->min_depth = I.min_depth;
        wp->max_depth = I.max_depth;
        wp->dofs = 0;
        wp->item = T.item_first + win_base + lane;
        wp->gi_base = 0;
        const float rsrc = FORCE ? __fmul_rn(sc.h_max, KERNEL_GAMMA) : 0.f;
        wp->rsrc = rsrc;
        float hcap = 3.402823466e+38f, relq = 1.f, padd = 0.f;
        int dbl = 0, nokey = 0;
        double o0, o1, o2, s0 = 0., s1 = 0., s2 = 0.;
        /* displacement target - source = (t_own + T.loc) - (s_own + S.loc) - sh_eff */
        double ex = 0., ey = 0., ez = 0.; /* sh_eff */
        if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
          if (mode == MODE_PAIR_L) { /* pix = x - (cj->loc + shift), targets are the pi, cj = source cell */
            o0 = __dadd_rn(sc.loc[0], shx);
            o1 = __dadd_rn(sc.loc[1], shy);
            o2 = __dadd_rn(sc.loc[2], shz);
            ex = shx; ey = shy; ez = shz;
          } else { /* pjx = x - cj->loc, targets are the pj, cj = target cell */
            o0 = tcell.loc[0];
            o1 = tcell.loc[1];
            o2 = tcell.loc[2];
            ex = -shx; ey = -shy; ez = -shz;
          }
          if (FORCE) {
            hcap = __fmul_rn(tcell.h_max, KERNEL_GAMMA);
          } else {
            const float ci_hma = (mode == MODE_PAIR_L) ? tcell.h_max_allowed : sc.h_max_allowed;
            const float h_max_lim = (I.flags & 1) ? ci_hma : 3.402823466e+38f;
            hcap = __fmul_rn(fminf(h_max_lim, tcell.h_max_active), KERNEL_GAMMA);
          }
        } else if (mode == MODE_SUB_SELF) { /* floats relative to c->loc, :1108 */
          o0 = sc.loc[0];
          o1 = sc.loc[1];
          o2 = sc.loc[2];
          nokey = 1;
        } else {
          /* double modes: prefilter in the source cell's own frame, exact in the drain */
          dbl = 1;
          if (mode != MODE_SELF) {
            s0 = shx; s1 = shy; s2 = shz;
            ex = shx; ey = shy; ez = shz;
          } else {
            nokey = 1;
          }
          o0 = __dadd_rn(sc.loc[0], s0);
          o1 = __dadd_rn(sc.loc[1], s1);
          o2 = __dadd_rn(sc.loc[2], s2);
          /* limit r^2 of a prefilter radius (R (1 + 1e-5) + margin)^2 <= R^2 PRE_REL2 + padd for R <= Rmax */
          const float Rmax = fmaxf(T.rmax, rsrc);
          relq = PL_PRE_REL2;
          padd = fmaf(2.f * Rmax, A.margin, A.margin * A.margin) * 1.001f;
        }
        wp->ot[0] = o0; wp->ot[1] = o1; wp->ot[2] = o2;
        wp->sh[0] = s0; wp->sh[1] = s1; wp->sh[2] = s2;
        wp->hcap = hcap;
        wp->relq = relq;
        wp->padd = padd;
        wp->dbl = (int8_t)dbl;
        wp->nokey = (int8_t)nokey;
        const float d0 = (float)__dsub_rn(__dadd_rn(sc.loc[0], ex), tcell.loc[0]);
        const float d1 = (float)__dsub_rn(__dadd_rn(sc.loc[1], ey), tcell.loc[1]);
        const float d2 = (float)__dsub_rn(__dadd_rn(sc.loc[2], ez), tcell.loc[2]);
        wp->d[0] = d0; wp->d[1] = d1; wp->d[2] = d2;
        sWinAux[buf * 32 + lane] = make_int4(sc.first, sc.count, (int)I.sframe, W.bfirst);
        /* item-level cull: the source cell's box [0, width] (+ drift) against the task's target box */
        const float r = fmaf(fmaxf(T.rmax, rsrc), PREFILTER_REL, A.margin) + sc.dx_max_part;
        const float dd[3] = {d0, d1, d2};
        float q2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float a = 0.f - (T.hi[k] - dd[k]), b = (T.lo[k] - dd[k]) - sc.width;
          const float gk = fmaxf(0.f, fmaxf(a, b));
          q2 = fmaf(gk, gk, q2);
        }
        keep = (q2 < r * r) && sc.count > 0;
      }
      __syncwarp();
      return __ballot_sync(FULL_MASK, keep);
    };

    /* ---- task fetch (bootstrap and empty-ring fallback: all loads back to back) ---- */
    auto draw = [&]() -> int {
      int t = 0;
      if (lane == 0) t = (int)atomicAdd(A.task_counter, 1u);
      return __shfl_sync(FULL_MASK, t, 0);
    };

    TaskRec cur, nxt;
    CellLite cur_cell, nxt_cell;
    unsigned cur_km = 0, nxt_km = 0;
    int cur_buf = 0;
    bool cur_valid = false, nxt_valid = false;
    {
      const int t = draw();
      if (t < ntask) {
        cur = A.task_recs[t];
        cur_cell = load_cell(cur.tcell);
        WinLoad W;
        win_load_items(cur, 0, W);
        win_load_cells(W);
        cur_km = win_finish(cur, cur_cell, 0, W, 0);
        cur_valid = true;
      }
    }
    int it = 0; /* stage counter of the whole CTA life */
    while (cur_valid) {
      /* prefetch state of the NEXT task: 0 draw, 1 record, 2 items, 3 cells, 4 window, 5 done */
      int pstate = 0;
      int pf_task = 0;
      WinLoad PW;
      PW.valid = false;
      nxt_valid = false;
      auto prefetch_step = [&]() {
        switch (pstate) {
          case 0:
            if (lane == 0) pf_task = (int)atomicAdd(A.task_counter, 1u);
            pstate = 1;
            break;
          case 1: {
            pf_task = __shfl_sync(FULL_MASK, pf_task, 0);
            if (pf_task >= ntask) {
              pstate = 5;
            } else {
              nxt = A.task_recs[pf_task];
              pstate = 2;
            }
            break;
          }
          case 2:
            nxt_cell = load_cell(nxt.tcell);
            win_load_items(nxt, 0, PW);
            pstate = 3;
            break;
          case 3:
            win_load_cells(PW);
            pstate = 4;
            break;
          case 4:
            nxt_km = win_finish(nxt, nxt_cell, 0, PW, cur_buf ^ 1);
            nxt_valid = true;
            pstate = 5;
            break;
          default:
            break;
        }
      };

      const TaskRec &T = cur;
      bool first_stage = true; /* the first published stage of the task carries the task's targets */
      unsigned km = cur_km;    /* kept items of the window */

      /* ---- publish one stage: lanes with `has` copy their fragment (lane-parallel), lane 0 arrives ---- */
      auto publish = [&](bool has, int myfrag, int nfr, int used, int my_w, int my_off, int my_n, int my_pool,
                         int my_dbase, int gap_from, int gap_to) {
        const int s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        char *const st = smem + s * SM::kStageBytes;
        int32_t *const meta = (int32_t *)(st + SM::kStageMeta);
        mbar_wait(sEmpty + s, ph ^ 1u);
        uint32_t bytes = 0;
        if (has) {
          const int4 aux = sWinAux[cur_buf * 32 + my_w];
          const int first = aux.x + my_off;
          const int dpar = first & 1; /* the double columns are copied from an even index */
          const bool pdbl = sWin[cur_buf * 32 + my_w].dbl != 0;
          {
            const int4 *const src4 = (const int4 *)&sWin[cur_buf * 32 + my_w];
            int4 *const dst4 = (int4 *)&((PipeItem *)(st + SM::kStageIT))[myfrag];
#pragma unroll
            for (int q = 0; q < (int)sizeof(PipeItem) / 16; q++) dst4[q] = src4[q];
            PipeItem *const dp = (PipeItem *)dst4;
            dp->gi_base = first - my_pool;
            dp->dofs = (int16_t)(my_dbase + dpar - my_pool);
          }
          /* octet -> fragment map, sentinel records of the padding slots, sentinel boxes of a gap */
          const int o0 = my_pool >> 3, o1 = (my_pool + my_n + 7) >> 3;
          uint8_t *o2f = (uint8_t *)(st + SM::kStageO2F);
          for (int o = o0; o < o1; o++) o2f[o] = (uint8_t)myfrag;
          float4 *F = (float4 *)(st + SM::kStageF);
          for (int k = my_pool + my_n; k < o1 * 8; k++)
            F[k] = make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
          /* bulk copies */
          const uint32_t n16 = (uint32_t)my_n * 16u;
          tma_load(F + my_pool, A.frames + ((size_t)(uint32_t)aux.z + (size_t)my_off), n16, sFull + s);
          float4 *P = (float4 *)(st + SM::kStageP);
          /* first payload column: (m, v) - the limiter loop only needs the source's time_bin (fq2.w) */
          tma_load(P + my_pool, (LOOP == LOOP_LIMITER ? A.fq2 : A.mv) + first, n16, sFull + s);
          bytes = 2u * n16;
          if (LOOP == LOOP_GRADIENT) {
            tma_load(P + SL + my_pool, A.gq + first, n16, sFull + s);
            bytes += n16;
          }
          if (FORCE) {
            tma_load(P + SL + my_pool, A.fq1 + first, n16, sFull + s);
            tma_load(P + 2 * SL + my_pool, A.fq2 + first, n16, sFull + s);
            bytes += 2u * n16;
            if (SCHEME == SCH_SPHENIX) {
              tma_load(P + 3 * SL + my_pool, A.fq3 + first, n16, sFull + s);
              bytes += n16;
            }
          }
          if (pdbl) {
            const uint32_t n8 = (uint32_t)((my_n + dpar + 1) & ~1) * 8u;
            double *D = (double *)(st + SM::kStageD) + my_dbase;
            tma_load(D, A.xs0 + (first - dpar), n8, sFull + s);
            tma_load(D + SM::kDCol, A.xs1 + (first - dpar), n8, sFull + s);
            tma_load(D + 2 * SM::kDCol, A.xs2 + (first - dpar), n8, sFull + s);
            bytes += 3u * n8;
          }
          const uint32_t nb32 = (uint32_t)(o1 - o0) * 32u;
          tma_load(st + SM::kStageOB + o0 * 32, A.boxes + 2 * ((size_t)aux.w + (size_t)(my_off >> 3)), nb32,
                   sFull + s);
          bytes += nb32;
        }
        /* octets no fragment owns (an item's padding to the minimum size, possibly carried over from the
         * previous stage): boxes no cull accepts */
        for (int o = gap_from >> 3; o < (gap_to >> 3); o++) {
          ((uint8_t *)(st + SM::kStageO2F))[o] = 0;
          ((float4 *)(st + SM::kStageOB))[2 * o] = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
          ((float4 *)(st + SM::kStageOB))[2 * o + 1] = make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
        }
        bytes = __reduce_add_sync(FULL_MASK, bytes);
        __syncwarp();
        if (lane == 0) {
          meta[PM_NFR] = nfr;
          meta[PM_NOCT] = used >> 3;
          meta[PM_FLAG] = first_stage ? 1 : 0;
          meta[PM_TASK] = 0;
          meta[PM_TGT_OFF] = T.tgt_off;
          meta[PM_NTGT] = T.ntgt;
          double *const ml = (double *)(meta + PM_LOC);
          ml[0] = cur_cell.loc[0];
          ml[1] = cur_cell.loc[1];
          ml[2] = cur_cell.loc[2];
          mbar_arrive_tx(sFull + s, bytes);
        }
        first_stage = false;
        it++;
        if (!(A.hold & 1)) prefetch_step(); /* one dependent load of the next task per stage */
      };

      for (int win_base = 0;; win_base += 32) {
        if (win_base > 0) {
          /* a further window of a long item list (multi-level trees): loaded in place */
          WinLoad W;
          win_load_items(T, win_base, W);
          win_load_cells(W);
          km = win_finish(T, cur_cell, win_base, W, cur_buf);
        }
        /* ---- layout of the window's kept items in a virtual source array cut into stages of
         * SL (lane = item): exclusive prefix sum of the padded sizes. An item takes at least
         * 40 slots, so a stage holds at most 2 partial + 6 whole items = PL_FRAGS fragments. ---- */
        const int4 myaux = sWinAux[cur_buf * 32 + lane];
        const bool kept = (km >> lane) & 1u;
        const bool mydbl = kept && sWin[cur_buf * 32 + lane].dbl != 0;
        const int n = kept ? myaux.y : 0;
        const int pn = (n + 7) & ~7;
        /* a double-mode item larger than the stage's double columns: serial fallback below */
        /* (... or, in the main loops whose double columns hold one small item, a second double-mode item) */
        const bool big = (A.hold & 2) || __any_sync(FULL_MASK, mydbl && pn > (DS & ~7)) ||
                         (DS < SL && __popc(__ballot_sync(FULL_MASK, mydbl)) > 1);
        if (!big) {
          const int v = kept ? max(pn, 40) : 0;
          int V = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, V, o);
            if (lane >= o) V += t;
          }
          const int vstart = V - v;
          const int total = __shfl_sync(FULL_MASK, V, 31);
          const int nstages = (total + SL - 1) / SL;
          for (int k = 0; k < nstages; k++) {
            const int lo = k * SL, hi = lo + SL;
            const int a0 = max(vstart, lo), b0 = min(vstart + pn, hi);
            const bool has = kept && a0 < b0;
            const int my_off = a0 - vstart;
            const int my_n = min(n - my_off, b0 - a0);
            const int my_pool = a0 - lo;
            const unsigned fmask = __ballot_sync(FULL_MASK, has);
            const int nfr = __popc(fmask);
            const int myfrag = __popc(fmask & ((1u << lane) - 1u));
            const int used = __reduce_max_sync(FULL_MASK, has ? my_pool + ((my_n + 7) & ~7) : 0);
            /* my padding (to the minimum size) inside this stage, below the last real slot: it must not
             * look like sources */
            const int gap_from = kept ? min(max(vstart + pn, lo), hi) - lo : 0;
            const int gap_to = kept ? min(min(max(vstart + v, lo), hi) - lo, used) : 0;
            publish(has, myfrag, nfr, used, lane, my_off, my_n, my_pool, DS >= SL ? my_pool + 2 * myfrag : 0,
                    gap_from, gap_to);
          }
        } else {
          /* ---- serial assembly (stage by stage, warp-uniform) ---- */
          int j = 0;   /* next position in the window */
          int off = 0; /* source offset inside the current item */
          for (;;) {
            int used = 0, nfr = 0, dused = 0;
            int my_w = 0, my_off = 0, my_n = 0, my_pool = 0, my_dbase = 0;
            while (nfr < PL_FRAGS && used < SL) {
              const unsigned mm = j >= 32 ? 0u : (km >> j) << j;
              if (!mm) break;
              const int jj = __ffs(mm) - 1;
              const int4 aux = sWinAux[cur_buf * 32 + jj];
              const bool dbl = sWin[cur_buf * 32 + jj].dbl != 0;
              const int left = aux.y - off, room = SL - used;
              int take = left <= room ? left : (room & ~7);
              if (dbl) {
                const int droom = (DS - dused) & ~7;
                if (take > droom) take = droom;
              }
              if (take <= 0) break;
              if (lane == nfr) {
                my_w = jj;
                my_off = off;
                my_n = take;
                my_pool = used;
                my_dbase = dused;
              }
              used += (take + 7) & ~7;
              if (dbl) dused += ((take + 7) & ~7) + 2;
              nfr++;
              if (take == left) {
                j = jj + 1;
                off = 0;
              } else {
                off += take;
                break;
              }
            }
            if (nfr == 0) break; /* the window's items are exhausted */
            publish(lane < nfr, lane, nfr, used, my_w, my_off, my_n, my_pool, my_dbase, 0, 0);
          }
        }
        if (win_base + 32 >= T.item_count) break;
      }
      while (pstate != 5) prefetch_step();
      cur_valid = nxt_valid;
      if (nxt_valid) {
        cur = nxt;
        cur_cell = nxt_cell;
        cur_km = nxt_km;
        cur_buf ^= 1;
      }
    }
    /* terminator stage */
    {
      const int s = it % NS;
      const uint32_t ph = (uint32_t)((it / NS) & 1);
      int32_t *const meta = (int32_t *)(smem + s * SM::kStageBytes + SM::kStageMeta);
      mbar_wait(sEmpty + s, ph ^ 1u);
      if (lane == 0) {
        meta[PM_NFR] = 0;
        meta[PM_NOCT] = 0;
        meta[PM_FLAG] = 2;
        meta[PM_TASK] = -1;
        mbar_arrive(sFull + s);
      }
    }
    return;
  }

  /* ===================================================================== */
  /* CONSUMERS                                                              */
  /* ===================================================================== */
  auto slow_args = [&]() {
    SlowArgs SA;
    SA.items = A.items;
    SA.cells = A.cells;
    SA.ext = A.ext;
    SA.dim[0] = A.dim[0];
    SA.dim[1] = A.dim[1];
    SA.dim[2] = A.dim[2];
    return SA;
  };

  /* my target of the current task (4 lanes share one) */
  bool tvalid = false;
  int ti = -1, tdepth = 0, ttb = 0;
  double tx = 0., ty = 0., tz = 0.;
  float th = 1.f, tvx = 0.f, tvy = 0.f, tvz = 0.f, tu = 0.f, tcs = 0.f;
  float thg2 = 0.f, th_inv = 1.f, thg = 0.f, tsure2 = 0.f, r2e = 0.f;
  ForceQ tq;
  float4 *const sTP = (float4 *)(smem + SM::kTP) + warp * (2 * PL_FRAGS * 8); /* [par][frag][t8] */

  DensityAcc dacc;
  GradientAcc gacc;
  ForceAcc facc;
  int nhit = 0;
  int ntests = 0, nhit_all = 0;
  int nsub = 0; /* entries in my sub-list */
  uint16_t *const wlist = sList + warp * (QCAP * 32);
  uint16_t *const mylist = wlist + lane;

  /* ---- INTERACT: merge the 4 sub-lists of each target and drain ---- */
  auto drain = [&]() {
    __syncwarp();
    const int n0 = __shfl_sync(FULL_MASK, nsub, t8);
    const int n1 = __shfl_sync(FULL_MASK, nsub, t8 + 8);
    const int n2 = __shfl_sync(FULL_MASK, nsub, t8 + 16);
    const int n3 = __shfl_sync(FULL_MASK, nsub, t8 + 24);
    const int c1 = n0 + n1, c2 = c1 + n2, total = c2 + n3;
    const int steps = (__reduce_max_sync(FULL_MASK, total) + 3) >> 2;
    for (int jstep = 0; jstep < steps; jstep++) {
      const int m = 4 * jstep + s4;
      const bool act = m < total;
      /* sub-list q and index in it of position m of the concatenated list (selects, no branches) */
      const bool q1 = m >= n0, q2 = m >= c1, q3 = m >= c2;
      const int base = q3 ? c2 : (q2 ? c1 : (q1 ? n0 : 0));
      const int q = (int)q1 + (int)q2 + (int)q3;
      const int kk = act ? m - base : 0;
      /* entry: dbl(15) | parity(14) | fragment(13:11) | ring slot(10:8) | source slot(7:0) */
      const int entry = act ? (int)wlist[kk * 32 + t8 + 8 * q] : 0;
      const int sl = entry & 255;
      const char *const st = smem + ((entry >> 8) & 7) * SM::kStageBytes;
      const int fr = (entry >> 11) & 7;
      const PipeItem &ii = ((const PipeItem *)(st + SM::kStageIT))[fr];
      const float4 src = ((const float4 *)(st + SM::kStageF))[sl];
      const float4 tp = sTP[(((entry >> 14) & 1) * PL_FRAGS + fr) * 8 + t8];
      float dx = __fsub_rn(tp.x, src.x), dy = __fsub_rn(tp.y, src.y), dz = __fsub_rn(tp.z, src.z);
      const bool dbl = (entry >> 15) != 0;
      const int gi = ii.gi_base + sl;
      double Xx = 0., Xy = 0., Xz = 0.;
      if (__any_sync(FULL_MASK, dbl)) {
        if (dbl) {
          const double *const D = (const double *)(st + SM::kStageD) + sl + ii.dofs;
          Xx = D[0];
          Xy = D[SM::kDCol];
          Xz = D[2 * SM::kDCol];
          dx = dsubf(__dsub_rn(tx, ii.sh[0]), Xx);
          dy = dsubf(__dsub_rn(ty, ii.sh[1]), Xy);
          dz = dsubf(__dsub_rn(tz, ii.sh[2]), Xz);
        }
      }
      const float r2 = r2_exact(dx, dy, dz);
      /* the depth-range rule was applied when the candidate was listed (test loop) */
      const bool part = act && gi != ti;
      const float4 *const P = (const float4 *)(st + SM::kStageP);
      if (!FORCE) {
        bool hit = part && (r2 < thg2);
        if (hit && !ii.nokey && !(r2 < tsure2 && thg <= ii.hcap)) {
          if (!dbl) {
            Xx = A.xs0[gi];
            Xy = A.xs1[gi];
            Xz = A.xs2[gi];
          }
          hit = exact_type1(slow_args(), ii.item, tx, ty, tz, thg, Xx, Xy, Xz);
        }
        if (hit) {
          const float4 f0 = P[sl];
          if (LOOP == LOOP_DENSITY) {
            iact_density(dacc, r2, dx, dy, dz, th_inv, tvx, tvy, tvz, f0.x, f0.y, f0.z, f0.w);
          } else if (LOOP == LOOP_LIMITER) {
            /* runner_iact_nonsym_limiter, timestep_limiter_iact.h:106-117: wake up the neighbour? */
            if (__float_as_int(f0.w) > ttb + TIME_BIN_NEIGHBOUR_MAX_DELTA_BIN) atomicMax(&A.wakeup[gi], -ttb);
          } else {
            const float4 f1 = P[SL + sl];
            iact_gradient(gacc, r2, dx, dy, dz, th, tvx, tvy, tvz, tu, tcs, f0.x, f0.y, f0.z, f0.w,
                          f1.x, f1.y, f1.z, f1.w, A.a2_Hubble);
          }
          nhit++;
        }
      } else {
        const float4 q2 = P[2 * SL + sl];
        const float sh = act ? q2.y : 1.f;
        const float shg2 = hg2_exact(sh);
        const bool a1 = r2 < thg2, a2 = r2 < shg2;
        bool ok = part && (a1 || a2);
        if (ok && !ii.nokey) {
          /* DOSELF2 (nokey) takes r2 < hig2 || r2 < hjg2 as is (:2792) */
          const float shg = __fmul_rn(sh, KERNEL_GAMMA);
          const bool sure = (!a1 || (r2 < tsure2 && thg <= ii.hcap)) &&
                            (!a2 || (r2 < sure_r2(shg, A.keyE) && shg <= ii.rsrc));
          if (!sure) {
            if (!dbl) {
              Xx = A.xs0[gi];
              Xy = A.xs1[gi];
              Xz = A.xs2[gi];
            }
            ok = exact_type2(slow_args(), ii.item, tx, ty, tz, thg, thg2, Xx, Xy, Xz, shg, shg2, r2);
          }
        }
        if (ok) {
          ForceQ sq;
          const float4 q0 = P[sl], q1 = P[SL + sl];
          sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
          sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
          sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
          sq.alpha_visc = sq.alpha_diff = 0.f;
          if (SCHEME == SCH_SPHENIX) {
            const float4 q3 = P[3 * SL + sl];
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

  const int hold_max = min(NS - 1, 2);
  int held = 0; /* stages tested but not yet released */
  int it = 0;   /* next stage to wait for (whole CTA life) */
  for (;;) {
    /* ---- next task: its first stage carries the targets ---- */
    float tox = 0.f, toy = 0.f, toz = 0.f; /* own-frame floats of my target (cull box) */
    {
      const int s0 = it % NS;
      mbar_wait(sFull + s0, (uint32_t)((it / NS) & 1));
      const int32_t *const meta0 = (const int32_t *)(smem + s0 * SM::kStageBytes + SM::kStageMeta);
      if (meta0[PM_FLAG] == 2) break;
      /* the task's targets are dealt evenly to the CW warps (a sparse task - ghost re-runs, few
       * active particles - keeps every warp busy with a small target box instead of filling the
       * first warps only); 64 targets: 8 per warp as they lie */
      const int ntgt = meta0[PM_NTGT];
      const int per = (ntgt + CW - 1) / CW;
      const int slot_t = warp * per + t8;
      tvalid = t8 < per && slot_t < ntgt;
      ti = tvalid ? A.tgt_list[meta0[PM_TGT_OFF] + slot_t] : -1;
      const double *const ml = (const double *)(meta0 + PM_LOC);
      const double l0 = ml[0], l1 = ml[1], l2 = ml[2];
      tx = ty = tz = 0.;
      th = 1.f;
      tvx = tvy = tvz = tu = tcs = 0.f;
      tq.m = tq.vx = tq.vy = tq.vz = 0.f;
      tq.rho = 1.f;
      tq.P = tq.f = tq.cs = tq.balsara = 0.f;
      tq.h = 1.f;
      tq.u = tq.alpha_visc = tq.alpha_diff = 0.f;
      tq.time_bin = 0;
      tdepth = 0;
      if (tvalid) {
        tx = A.xs0[ti];
        ty = A.xs1[ti];
        tz = A.xs2[ti];
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
          if (LOOP == LOOP_LIMITER) ttb = A.time_bin[ti];
          if (LOOP == LOOP_GRADIENT) {
            tu = A.fq2[ti].z;
            tcs = A.fq1[ti].w;
          }
        }
      }
      thg2 = hg2_exact(th);
      th_inv = 1.f / th;
      thg = __fmul_rn(th, KERNEL_GAMMA);
      tsure2 = sure_r2(thg, A.keyE);
      {
        const float re = fmaf(thg, PREFILTER_REL, A.margin);
        r2e = re * re;
      }
      tox = dsubf(tx, l0);
      toy = dsubf(ty, l1);
      toz = dsubf(tz, l2);
      __syncwarp(); /* the previous task's drains are done with sBox / sTP */
      /* the warp's target box (own-frame floats of the target cell) and its reach */
      float blo[3], bhi[3], rmax;
      blo[0] = tvalid ? tox : 3.0e30f;
      blo[1] = tvalid ? toy : 3.0e30f;
      blo[2] = tvalid ? toz : 3.0e30f;
      bhi[0] = tvalid ? tox : -3.0e30f;
      bhi[1] = tvalid ? toy : -3.0e30f;
      bhi[2] = tvalid ? toz : -3.0e30f;
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
      __syncwarp();
      dacc.zero();
      gacc.v_sig = 0.f;
      gacc.laplace_u = 0.f;
      gacc.alpha_max = 0.f;
      facc.ax = facc.ay = facc.az = facc.u_dt = facc.h_dt = 0.f;
      facc.v_sig = 0.f;
      facc.min_ngb = NUM_TIME_BINS + 1;
      nhit = 0;
    }
    const int first_it = it;
    const bool wvalid = __any_sync(FULL_MASK, tvalid);
    /* ---- stage loop: one state machine with a single drain() call site ---- */
    int s = 0;      /* ring slot of the stage under test */
    unsigned m = 0; /* accepted octets of that stage still to test */
    bool done = false;
    for (;;) {
      if (m == 0) {
        /* ---- next stage ---- */
        s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        mbar_wait(sFull + s, ph);
        const char *const st = smem + s * SM::kStageBytes;
        const int32_t *const meta = (const int32_t *)(st + SM::kStageMeta);
        if (meta[PM_FLAG] == 2 || (meta[PM_FLAG] == 1 && it != first_it)) {
          done = true; /* terminator, or the first stage of the next task: leave it where it is */
        } else {
          const int noct = meta[PM_NOCT];
          const int par = it & 1;
          it++;
          held++;
          /* ---- cull: lane = octet ---- */
          bool acc = false;
          int myfl = 0;
          if (lane < noct && wvalid) {
            const float *const wb = sBox + warp * 8; /* the warp's target box and reach */
            const float rmax = wb[6];
            const float4 lo = ((const float4 *)(st + SM::kStageOB))[2 * lane];
            const float4 hi = ((const float4 *)(st + SM::kStageOB))[2 * lane + 1];
            myfl = ((const uint8_t *)(st + SM::kStageO2F))[lane];
            const PipeItem &ii = ((const PipeItem *)(st + SM::kStageIT))[myfl];
            const float r = fmaf(FORCE ? fmaxf(rmax, ii.rsrc) : rmax, PREFILTER_REL, A.margin);
            float d2;
            {
              const float a = lo.x - (wb[3] - ii.d[0]), b = (wb[0] - ii.d[0]) - hi.x;
              const float gx = fmaxf(0.f, fmaxf(a, b));
              d2 = gx * gx;
            }
            {
              const float a = lo.y - (wb[4] - ii.d[1]), b = (wb[1] - ii.d[1]) - hi.y;
              const float gy = fmaxf(0.f, fmaxf(a, b));
              d2 = fmaf(gy, gy, d2);
            }
            {
              const float a = lo.z - (wb[5] - ii.d[2]), b = (wb[2] - ii.d[2]) - hi.z;
              const float gz = fmaxf(0.f, fmaxf(a, b));
              d2 = fmaf(gz, gz, d2);
            }
            acc = d2 < r * r;
          }
          m = __ballot_sync(FULL_MASK, acc);
          /* ---- the target's coordinates in the frame of every fragment with an accepted octet ---- */
          unsigned fm = __reduce_or_sync(FULL_MASK, acc ? (1u << myfl) : 0u);
          unsigned keepf = 0u; /* fragments in which at least one target of the warp takes part */
          while (fm) {
            const int fr = __ffs(fm) - 1;
            fm &= fm - 1u;
            const PipeItem &ii = ((const PipeItem *)(st + SM::kStageIT))[fr];
            const bool part = tvalid && tdepth >= ii.min_depth && tdepth <= ii.max_depth;
            float4 tp;
            tp.x = part ? dsubf(tx, ii.ot[0]) : 3.0e30f;
            tp.y = dsubf(ty, ii.ot[1]);
            tp.z = dsubf(tz, ii.ot[2]);
            tp.w = ii.dbl ? r2e : thg2;
            if (s4 == 0) sTP[(par * PL_FRAGS + fr) * 8 + t8] = tp;
            if (__any_sync(FULL_MASK, part)) keepf |= 1u << fr;
          }
          /* drop the octets of fragments no target of this warp takes part in */
          if (!((keepf >> myfl) & 1u)) acc = false;
          m = __ballot_sync(FULL_MASK, acc);
          __syncwarp();
        }
      }
      if (m) {
        /* ---- test the accepted octets (until done or a sub-list may overflow) ---- */
        const char *const st = smem + s * SM::kStageBytes;
        const float4 *const F = (const float4 *)(st + SM::kStageF);
        const uint8_t *const o2f = (const uint8_t *)(st + SM::kStageO2F);
        const PipeItem *const IT = (const PipeItem *)(st + SM::kStageIT);
        const int par = (it - 1) & 1;
        const float4 *const TPp = sTP + par * PL_FRAGS * 8 + t8;
        int nsub_ub = __reduce_max_sync(FULL_MASK, nsub);
        while (m) {
          if (nsub_ub > QCAP - 2) {
            nsub_ub = __reduce_max_sync(FULL_MASK, nsub);
            if (nsub_ub > QCAP - 2) break; /* drain first, then come back to this octet */
          }
          const int o = __ffs(m) - 1;
          m &= m - 1u;
          nsub_ub += 2;
          const int fl = o2f[o];
          const float4 tp = TPp[fl * 8];
          const int sl = o * 8 + 2 * s4;
          const float4 a = F[sl], c = F[sl + 1];
          const int code = (par << 14) | (fl << 11) | (s << 8) | sl;
          ntests++;
          float lima, limc;
          int cdbl = 0;
          if (FORCE) {
            const float *const HQ = (const float *)(st + SM::kStageP) + (PL_HG2_COL(SCHEME) * SL + sl) * 4 + 2;
            const float2 rq = *(const float2 *)&IT[fl].relq;
            lima = fmaf(fmaxf(thg2, HQ[0]), rq.x, rq.y);
            limc = fmaf(fmaxf(thg2, HQ[4]), rq.x, rq.y);
            cdbl = rq.x != 1.f ? 0x8000 : 0;
          } else {
            lima = limc = tp.w;
            cdbl = tp.w != thg2 ? 0x8000 : 0;
          }
          {
            const float r2 = r2_exact(__fsub_rn(tp.x, a.x), __fsub_rn(tp.y, a.y), __fsub_rn(tp.z, a.z));
            if (r2 < lima) {
              mylist[nsub * 32] = (uint16_t)(code | cdbl);
              nsub++;
            }
          }
          {
            const float r2 = r2_exact(__fsub_rn(tp.x, c.x), __fsub_rn(tp.y, c.y), __fsub_rn(tp.z, c.z));
            if (r2 < limc) {
              mylist[nsub * 32] = (uint16_t)((code + 1) | cdbl);
              nsub++;
            }
          }
        }
      }
      if (m != 0 || held >= hold_max || done) {
        drain();
        if (m == 0) { /* every held stage is fully tested and drained: give them back */
          if (lane == 0)
            for (int k = 1; k <= held; k++) mbar_arrive(sEmpty + ((it - k) % NS));
          held = 0;
        }
      }
      if (done) break;
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
    } else if (FORCE) {
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
    if (LOOP != LOOP_LIMITER && tvalid && s4 == 0 && nh) atomicAdd(&A.count[ti], nh);
    nhit_all += nhit;
  }
  int tot = nhit_all, tt = ntests;
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
