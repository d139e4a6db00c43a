/*
 * legacy/loops_tile.cuh - THIRD generation (round 1's default; superseded by loops_pipe.cuh; built only
 * with -DSWIFTGPU_LEGACY_LOOPS): the neighbour loops as a TMA-fed producer/consumer pipeline.
 *
 * One CTA = 8 consumer warps (64 TARGET particles of one target cell, 8 per
 * warp; 4 warps for sparse target sets) + 1 producer warp. No
 * __syncthreads() after the prologue.
 *
 * The CTAs are persistent: each draws tasks (8 targets per consumer warp of
 * one group) from a global counter and the ring runs across task boundaries.
 *
 *   PRODUCER  walks the items of the target cell's group, culls source cells
 *             that are out of reach of the CTA's target box and cuts the rest
 *             into fragments. A ring STAGE holds up to 256 source slots / 8
 *             fragments. The source data are copied as they lie in HBM by
 *             bulk TMA (cp.async.bulk -> mbarrier complete_tx): the
 *             prefilter record xf = (float x, y, z, reach^2), the three
 *             double position columns xs, the payload columns (mv, gq or
 *             fq1, fq2, fq3) and the precomputed OCTET boxes of the source
 *             cell. No per-source arithmetic is done at staging.
 *   CONSUMER  per stage: lane = octet box-box cull against the warp's target
 *             box (one ballot), then per accepted octet every lane
 *             (t = lane & 7 target, s = lane >> 3) tests its target against
 *             sources 2s, 2s+1 with a conservative float prefilter on
 *             absolute-position floats and appends candidates to its
 *             sub-list. Stages are held (not released) until the lists are
 *             drained; a drain merges the 4 sub-lists of a target over its 4
 *             lanes, re-evaluates every candidate with the reference's EXACT
 *             arithmetic from the double positions (frames, un-fused r2) and
 *             applies the interaction; then the held stages are released.
 *
 * Exactness. The reference decides a pair by r2 < h^2 gamma^2 in the float
 * frame of the leaf-level item (functions_hydro.h:1327-1347) AND by the
 * sorted-axis conditions of DOPAIR1/DOPAIR2 (:1296-1332, :1420-1448,
 * :1652-1735, :1806-2238). The sorted-axis conditions are geometrically
 * implied by the distance condition up to the rounding of the float sort keys
 * (|error| < 5e-7 * max(dim)); the drain therefore evaluates them only for
 * candidates whose r lies within keyE = 2e-6 * max(dim) of the cut-off (or
 * whose h exceeds the cell's h_max cap) - exact_type1()/exact_type2(), a
 * rare divergent path that re-derives the item constants from global memory
 * exactly as the reference does. Everything else takes the fast path, whose
 * accept set is provably the reference's.
 */
#ifndef SWIFTGPU_LOOPS_TILE_CUH
#define SWIFTGPU_LOOPS_TILE_CUH

#include "../loops_common.cuh"

namespace swiftgpu {

/* Constants of one fragment's item, written by the producer for the consumers. */
struct __align__(8) TileItem {
  double ot[3]; /* drain: subtracted from the target double */
  double fs[3]; /* drain: frame origin of the source (0 for the double modes) */
  float d[3];   /* prefilter: target absolute float - d is compared with the source absolute float */
  float rsrc;   /* force: h_max * gamma of the source cell (also the source-side cap) */
  float hcap;   /* target-side cap of h * gamma under which the key conditions are implied */
  int32_t gi_base; /* global particle index = gi_base + slot-in-stage */
  int32_t item;    /* global item index (slow path) */
  int8_t mode, sid, min_depth, max_depth;
  int8_t dbl, nokey, dofs, pad1_; /* dofs: slot -> index into the staged double columns */
};
static_assert(sizeof(TileItem) == 88, "TileItem");

template <int NP, int NS, int QCAP, int CW>
struct TileSmem {
  static constexpr int kStageF = 0;
  static constexpr int kStageP = kStageF + TL_SLOTS * 16;
  static constexpr int kStageD = kStageP + NP * TL_SLOTS * 16; /* 3 double columns of TL_DCOL entries */
  static constexpr int kStageOB = kStageD + 3 * TL_DCOL * 8;
  static constexpr int kStageIT = kStageOB + TL_OCT * 32;
  static constexpr int kStageO2F = kStageIT + TL_FRAGS * (int)sizeof(TileItem);
  static constexpr int kStageMeta = kStageO2F + TL_OCT;
  static constexpr int kStageBytes = ((kStageMeta + 16) + 127) & ~127;
  static constexpr int kList = NS * kStageBytes;
  static constexpr int kBar = kList + QCAP * 32 * CW * 2;
  static constexpr int kBox = kBar + 2 * NS * 8;
  static constexpr int kWin = kBox + (CW + 1) * 32; /* producer: constants of 32 items */
  static constexpr int kWinAux = kWin + 32 * (int)sizeof(TileItem);
  static constexpr int kTX = kWinAux + 32 * 8; /* target doubles: 3 columns of TL_TARGETS */
  static constexpr int kBytes = kTX + 3 * TL_TARGETS * 8;
};

/* Timing hooks of the development build (-DTL_TIMING): warp 0 of a few CTAs prints how its life splits
 * into waits, cull, test and drain (fenced clock64 reads). */
#ifdef TL_TIMING
#define TM(...) __VA_ARGS__
#define TCLK(v) { __syncwarp(); asm volatile("" ::: "memory"); v = clock64(); asm volatile("" ::: "memory"); }
#else
#define TM(...)
#endif

/* resident CTAs per SM the kernel is compiled for: 3 (type-1) / 2 (force) standard CTAs, 5 / 3 small ones */
#define TL_MIN_BLOCKS(LOOP, CW)                                     \
  ((CW) == 8 ? ((LOOP) == LOOP_FORCE ? 2 : TL_DENS_BLOCKS) \
             : ((LOOP) == LOOP_FORCE ? 3 : TL_SPARSE_BLOCKS))
template <int LOOP, int SCHEME, int NS, int CW>
__global__ void __launch_bounds__(32 * (CW + 1), TL_MIN_BLOCKS(LOOP, CW)) k_tile(const LoopArgs A) {
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  constexpr int QCAP = FORCE ? TL_SUBCAP2 : TL_SUBCAP1;
  typedef TileSmem<NP, NS, QCAP, CW> SM;
  constexpr int CTA_TGT = 8 * CW;          /* targets of one CTA task */
  constexpr int SUBS = TL_CWARPS / CW;     /* CTA tasks per 64-target chunk of the host task list */
  extern __shared__ __align__(128) char smem_tl[];
  char *const smem = smem_tl;
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
  /* PRODUCER (persistent: CTA tasks are drawn from a global counter; the   */
  /* ring runs across task boundaries, so the next task's sources are in    */
  /* flight while the consumers finish the current one)                     */
  /* ===================================================================== */
  if (!consumer) {
    TileItem *const sWin = (TileItem *)(smem + SM::kWin);
    int2 *const sWinAux = (int2 *)(smem + SM::kWinAux); /* (first particle, first octet box) of the source cell */
    int it = 0; /* stage counter of the whole CTA life */
    const int ntask_cta = A.ntasks * SUBS;
    for (;;) {
      int task = 0;
      if (lane == 0) task = (int)atomicAdd(A.task_counter, 1u);
      task = __shfl_sync(FULL_MASK, task, 0);
      if (task >= ntask_cta) break;
      const int tk = task / SUBS, sub = task % SUBS;
      const int g = A.task_group[tk];
      const int nt = A.tgt_count[g];
      const int tgt0 = A.task_chunk[tk] * TL_TARGETS + sub * CTA_TGT; /* first target slot of this CTA task */
      if (tgt0 >= nt) continue;
      const Group G = A.groups[g];
      /* the task's target box (absolute floats) and reach, from the target list */
      float clo[3], chi[3], crmax = 0.f;
      clo[0] = clo[1] = clo[2] = 3.0e30f;
      chi[0] = chi[1] = chi[2] = -3.0e30f;
      {
        const int tend = min(nt, tgt0 + CTA_TGT);
        for (int slot = tgt0 + lane; slot < tend; slot += 32) {
          const int ti = A.tgt_list[A.tgt_first[g] + slot];
          const float4 f = A.xf[ti];
          const float hh = FORCE ? A.fq2[ti].y : A.h[ti];
          clo[0] = fminf(clo[0], f.x); clo[1] = fminf(clo[1], f.y); clo[2] = fminf(clo[2], f.z);
          chi[0] = fmaxf(chi[0], f.x); chi[1] = fmaxf(chi[1], f.y); chi[2] = fmaxf(chi[2], f.z);
          crmax = fmaxf(crmax, __fmul_rn(hh, KERNEL_GAMMA));
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          clo[k] = warp_min(clo[k]);
          chi[k] = warp_max(chi[k]);
        }
        crmax = warp_max(crmax);
      }
      const DevCell tcell = A.cells[G.tcell];
      bool first_stage = true; /* the first published stage of the task carries the task id */
      int win_base = -32; /* first item of the window held in sWin */
      unsigned km = 0u;   /* kept items of the window */
      int wcount = 0;     /* my window item's source count */
      int j = 32;         /* next position in the window */
      int off = 0;        /* source offset inside the current item */
      bool exhausted = false;
      for (;; it++) {
        const int s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        /* ---- assemble the fragments of this stage (warp-uniform) ---- */
        int used = 0, nfr = 0;
        int my_w = 0, my_off = 0, my_n = 0, my_pool = 0;
        while (!exhausted && nfr < TL_FRAGS && used < TL_SLOTS) {
          const unsigned mm = j >= 32 ? 0u : (km >> j) << j;
          if (!mm) {
            if (nfr > 0) break; /* the window table is still referenced by this stage's fragments */
            win_base += 32;
            if (win_base >= G.item_count) {
              exhausted = true;
              break;
            }
            /* ---- load the constants of the next 32 items (lane = item) ---- */
            bool keep = false;
            wcount = 0;
            if (win_base + lane < G.item_count) {
              const int item = G.item_first + win_base + lane;
              const Item I = A.items[item];
              const DevCell sc = A.cells[I.scell];
              const int bfirst = A.cell_box_first[I.scell];
              wcount = sc.count;
              const int mode = I.mode;
              const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1],
                           shz = I.shift[2] * A.dim[2];
              TileItem ti_;
              ti_.mode = (int8_t)mode;
              ti_.sid = (int8_t)I.sid;
              ti_.min_depth = I.min_depth;
              ti_.max_depth = I.max_depth;
              ti_.dbl = 0;
              ti_.nokey = 0;
              ti_.dofs = ti_.pad1_ = 0;
              ti_.item = item;
              ti_.gi_base = 0;
              ti_.rsrc = FORCE ? __fmul_rn(sc.h_max, KERNEL_GAMMA) : 0.f;
              ti_.hcap = 3.402823466e+38f;
              double otx = 0., oty = 0., otz = 0., fsx = 0., fsy = 0., fsz = 0.;
              if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
                const DevCell &ci = (mode == MODE_PAIR_L) ? tcell : sc;
                const DevCell &cj = (mode == MODE_PAIR_L) ? sc : tcell;
                const double oix = __dadd_rn(cj.loc[0], shx), oiy = __dadd_rn(cj.loc[1], shy),
                             oiz = __dadd_rn(cj.loc[2], shz);
                if (mode == MODE_PAIR_L) {
                  otx = oix; oty = oiy; otz = oiz;
                  fsx = cj.loc[0]; fsy = cj.loc[1]; fsz = cj.loc[2];
                } else {
                  otx = cj.loc[0]; oty = cj.loc[1]; otz = cj.loc[2];
                  fsx = oix; fsy = oiy; fsz = oiz;
                }
                if (FORCE) {
                  ti_.hcap = __fmul_rn(tcell.h_max, KERNEL_GAMMA);
                } else {
                  const float h_max_lim = (I.flags & 1) ? ci.h_max_allowed : 3.402823466e+38f;
                  ti_.hcap = __fmul_rn(fminf(h_max_lim, tcell.h_max_active), KERNEL_GAMMA);
                }
                ti_.d[0] = (float)__dsub_rn(otx, fsx);
                ti_.d[1] = (float)__dsub_rn(oty, fsy);
                ti_.d[2] = (float)__dsub_rn(otz, fsz);
              } else if (mode == MODE_SUB_SELF) {
                otx = fsx = sc.loc[0];
                oty = fsy = sc.loc[1];
                otz = fsz = sc.loc[2];
                ti_.d[0] = ti_.d[1] = ti_.d[2] = 0.f;
                ti_.nokey = 1;
              } else {
                ti_.dbl = 1;
                if (mode != MODE_SELF) {
                  otx = shx; oty = shy; otz = shz;
                } else {
                  ti_.nokey = 1;
                }
                ti_.d[0] = (float)otx;
                ti_.d[1] = (float)oty;
                ti_.d[2] = (float)otz;
              }
              ti_.ot[0] = otx; ti_.ot[1] = oty; ti_.ot[2] = otz;
              ti_.fs[0] = fsx; ti_.fs[1] = fsy; ti_.fs[2] = fsz;
              sWin[lane] = ti_;
              sWinAux[lane] = make_int2(sc.first, bfirst);
              /* item-level cull: source cell box against the CTA's target box */
              const float r = fmaf(fmaxf(crmax, ti_.rsrc), PREFILTER_REL, A.margin) + sc.dx_max_part;
              const float c0[3] = {(float)sc.loc[0], (float)sc.loc[1], (float)sc.loc[2]};
              float q2 = 0.f;
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const float a = c0[k] - (chi[k] - ti_.d[k]), b = (clo[k] - ti_.d[k]) - (c0[k] + sc.width);
                const float gk = fmaxf(0.f, fmaxf(a, b));
                q2 = fmaf(gk, gk, q2);
              }
              keep = (q2 < r * r) && wcount > 0;
            }
            __syncwarp();
            km = __ballot_sync(FULL_MASK, keep);
            j = 0;
            continue;
          }
          const int jj = __ffs(mm) - 1;
          const int cnt = __shfl_sync(FULL_MASK, wcount, jj);
          const int left = cnt - off, room = TL_SLOTS - used;
          const int take = left <= room ? left : (room & ~7);
          if (take == 0) break;
          if (lane == nfr) {
            my_w = jj;
            my_off = off;
            my_n = take;
            my_pool = used;
          }
          used += (take + 7) & ~7;
          nfr++;
          if (take == left) {
            j = jj + 1;
            off = 0;
          } else {
            off += take;
            break;
          }
        }
        char *const st = smem + s * SM::kStageBytes;
        int32_t *const meta = (int32_t *)(st + SM::kStageMeta);
        if (nfr == 0) break; /* the task's items are exhausted */
        mbar_wait(sEmpty + s, ph ^ 1u);
        uint32_t bytes = 0;
        if (lane < nfr) {
          TileItem ti_ = sWin[my_w];
          const int2 aux = sWinAux[my_w];
          const int first = aux.x + my_off;
          const int dpar = first & 1; /* the double columns are copied from an even index */
          ti_.gi_base = first - my_pool;
          ti_.dofs = (int8_t)(2 * lane + dpar);
          ((TileItem *)(st + SM::kStageIT))[lane] = ti_;
          /* octet -> fragment map, sentinel records of the padding slots */
          const int o0 = my_pool >> 3, o1 = (my_pool + my_n + 7) >> 3;
          uint8_t *o2f = (uint8_t *)(st + SM::kStageO2F);
          for (int o = o0; o < o1; o++) o2f[o] = (uint8_t)lane;
          float4 *F = (float4 *)(st + SM::kStageF);
          for (int k = my_pool + my_n; k < o1 * 8; k++)
            F[k] = make_float4(-3.0e30f, -3.0e30f, -3.0e30f, 0.f);
          /* bulk copies */
          const uint32_t n16 = (uint32_t)my_n * 16u;
          tma_load(F + my_pool, A.xf + first, n16, sFull + s);
          float4 *P = (float4 *)(st + SM::kStageP);
          tma_load(P + my_pool, A.mv + first, n16, sFull + s);
          bytes = 2u * n16;
          if (LOOP == LOOP_GRADIENT) {
            tma_load(P + TL_SLOTS + my_pool, A.gq + first, n16, sFull + s);
            bytes += n16;
          }
          if (FORCE) {
            tma_load(P + TL_SLOTS + my_pool, A.fq1 + first, n16, sFull + s);
            tma_load(P + 2 * TL_SLOTS + my_pool, A.fq2 + first, n16, sFull + s);
            bytes += 2u * n16;
            if (SCHEME == SCH_SPHENIX) {
              tma_load(P + 3 * TL_SLOTS + my_pool, A.fq3 + first, n16, sFull + s);
              bytes += n16;
            }
          }
          const uint32_t n8 = (uint32_t)((my_n + dpar + 1) & ~1) * 8u;
          double *D = (double *)(st + SM::kStageD) + my_pool + 2 * lane;
          tma_load(D, A.xs0 + (first - dpar), n8, sFull + s);
          tma_load(D + TL_DCOL, A.xs1 + (first - dpar), n8, sFull + s);
          tma_load(D + 2 * TL_DCOL, A.xs2 + (first - dpar), n8, sFull + s);
          bytes += 3u * n8;
          const uint32_t nb32 = (uint32_t)(o1 - o0) * 32u;
          tma_load(st + SM::kStageOB + o0 * 32, A.boxes + 2 * ((size_t)aux.y + (size_t)(my_off >> 3)), nb32,
                   sFull + s);
          bytes += nb32;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(FULL_MASK, bytes, o);
        __syncwarp();
        if (lane == 0) {
          meta[0] = nfr;
          meta[1] = used >> 3;
          meta[2] = first_stage ? 1 : 0;
          meta[3] = task;
          mbar_arrive_tx(sFull + s, bytes);
        }
        first_stage = false;
      }
    }
    /* terminator stage */
    {
      const int s = it % NS;
      const uint32_t ph = (uint32_t)((it / NS) & 1);
      int32_t *const meta = (int32_t *)(smem + s * SM::kStageBytes + SM::kStageMeta);
      mbar_wait(sEmpty + s, ph ^ 1u);
      if (lane == 0) {
        meta[0] = 0;
        meta[1] = 0;
        meta[2] = 2;
        meta[3] = -1;
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
  int ti = -1, tdepth = 0;
  float th = 1.f, tvx = 0.f, tvy = 0.f, tvz = 0.f, tu = 0.f, tcs = 0.f;
  float thg2 = 0.f, th_inv = 1.f, thg = 0.f, tsure2 = 0.f, tfx = 0.f, tfy = 0.f, tfz = 0.f;
  ForceQ tq;
  double *const sTX = (double *)(smem + SM::kTX) + warp * 8 + t8;

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
      const int entry = act ? (int)wlist[kk * 32 + t8 + 8 * q] : 0;
      const int slot = entry & 2047;
      const char *const st = smem + (slot >> 8) * SM::kStageBytes;
      const int sl = slot & 255;
      const TileItem &ii = ((const TileItem *)(st + SM::kStageIT))[entry >> 11];
      const int gi = ii.gi_base + sl;
      const double *const D = (const double *)(st + SM::kStageD) + sl + ii.dofs;
      const double Xx = D[0], Xy = D[TL_DCOL], Xz = D[2 * TL_DCOL];
      float dx, dy, dz;
      {
        const double tx = sTX[0], ty = sTX[TL_TARGETS], tz = sTX[2 * TL_TARGETS];
        const double ax = __dsub_rn(tx, ii.ot[0]), ay = __dsub_rn(ty, ii.ot[1]), az = __dsub_rn(tz, ii.ot[2]);
        if (ii.dbl) {
          dx = dsubf(ax, Xx);
          dy = dsubf(ay, Xy);
          dz = dsubf(az, Xz);
        } else {
          dx = __fsub_rn(__double2float_rn(ax), dsubf(Xx, ii.fs[0]));
          dy = __fsub_rn(__double2float_rn(ay), dsubf(Xy, ii.fs[1]));
          dz = __fsub_rn(__double2float_rn(az), dsubf(Xz, ii.fs[2]));
        }
      }
      const float r2 = r2_exact(dx, dy, dz);
      /* the depth-range rule was applied when the candidate was listed (test loop) */
      const bool part = act && gi != ti;
      const float4 *const P = (const float4 *)(st + SM::kStageP);
      if (!FORCE) {
        bool hit = part && (r2 < thg2);
        if (hit && !ii.nokey && !(r2 < tsure2 && thg <= ii.hcap))
          hit = exact_type1(slow_args(), ii.item, sTX[0], sTX[TL_TARGETS], sTX[2 * TL_TARGETS], thg, Xx, Xy, Xz);
        if (hit) {
          const float4 f0 = P[sl];
          if (LOOP == LOOP_DENSITY) {
            iact_density(dacc, r2, dx, dy, dz, th_inv, tvx, tvy, tvz, f0.x, f0.y, f0.z, f0.w);
          } else {
            const float4 f1 = P[TL_SLOTS + sl];
            iact_gradient(gacc, r2, dx, dy, dz, th, tvx, tvy, tvz, tu, tcs, f0.x, f0.y, f0.z, f0.w,
                          f1.x, f1.y, f1.z, f1.w, A.a2_Hubble);
          }
          nhit++;
        }
      } else {
        const float4 q2 = P[2 * TL_SLOTS + sl];
        const float sh = act ? q2.y : 1.f;
        const float shg2 = hg2_exact(sh);
        const bool a1 = r2 < thg2, a2 = r2 < shg2;
        bool ok = part && (a1 || a2);
        if (ok && !ii.nokey) {
          /* DOSELF2 (nokey) takes r2 < hig2 || r2 < hjg2 as is (:2792) */
          const float shg = __fmul_rn(sh, KERNEL_GAMMA);
          const bool sure = (!a1 || (r2 < tsure2 && thg <= ii.hcap)) &&
                            (!a2 || (r2 < sure_r2(shg, A.keyE) && shg <= ii.rsrc));
          if (!sure) ok = exact_type2(slow_args(), ii.item, sTX[0], sTX[TL_TARGETS], sTX[2 * TL_TARGETS], thg, thg2, Xx, Xy, Xz, shg, shg2, r2);
        }
        if (ok) {
          ForceQ sq;
          const float4 q0 = P[sl], q1 = P[TL_SLOTS + sl];
          sq.m = q0.x; sq.vx = q0.y; sq.vy = q0.z; sq.vz = q0.w;
          sq.rho = q1.x; sq.P = q1.y; sq.f = q1.z; sq.cs = q1.w;
          sq.balsara = q2.x; sq.h = q2.y; sq.u = q2.z; sq.time_bin = __float_as_int(q2.w);
          sq.alpha_visc = sq.alpha_diff = 0.f;
          if (SCHEME == SCH_SPHENIX) {
            const float4 q3 = P[3 * TL_SLOTS + sl];
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


  const int hold_max = min(NS - 1, max(1, A.hold));
  TM(long long tm_t0, tm_wait = 0, tm_peek = 0, tm_setup = 0, tm_drain = 0, tm_first = 0, tm_cull = 0, tm_test = 0, tm_a, tm_b;
     int tm_nwait = 0, tm_ndrain = 0, tm_ntask = 0; TCLK(tm_t0);)
  int held = 0; /* stages tested but not yet released */
  int it = 0;   /* next stage to wait for (whole CTA life) */
  for (;;) {
    /* ---- next task: its first stage carries the task id ---- */
    {
      const int s0 = it % NS;
      TM(TCLK(tm_a);)
      mbar_wait(sFull + s0, (uint32_t)((it / NS) & 1));
      TM(TCLK(tm_b); tm_peek += tm_b - tm_a; tm_ntask++;)
      const int32_t *const meta0 = (const int32_t *)(smem + s0 * SM::kStageBytes + SM::kStageMeta);
      if (meta0[2] == 2) break;
      const int task = meta0[3];
      const int tk = task / SUBS, sub = task % SUBS;
      const int g = A.task_group[tk];
      const int nt = A.tgt_count[g];
      const int slot_t = A.task_chunk[tk] * TL_TARGETS + sub * CTA_TGT + warp * 8 + t8;
      tvalid = slot_t < min(nt, A.task_chunk[tk] * TL_TARGETS + (sub + 1) * CTA_TGT);
      ti = tvalid ? A.tgt_list[A.tgt_first[g] + slot_t] : -1;
      double tx = 0., ty = 0., tz = 0.;
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
      thg2 = hg2_exact(th);
      th_inv = 1.f / th;
      thg = __fmul_rn(th, KERNEL_GAMMA);
      tsure2 = sure_r2(thg, A.keyE);
      tfx = __double2float_rn(tx);
      tfy = __double2float_rn(ty);
      tfz = __double2float_rn(tz);
      __syncwarp(); /* the previous task's drains are done with sTX / sBox */
      if (s4 == 0) {
        sTX[0] = tx;
        sTX[TL_TARGETS] = ty;
        sTX[2 * TL_TARGETS] = tz;
      }
      /* the warp's target box (absolute floats) and its reach */
      float blo[3], bhi[3], rmax;
      blo[0] = tvalid ? tfx : 3.0e30f;
      blo[1] = tvalid ? tfy : 3.0e30f;
      blo[2] = tvalid ? tfz : 3.0e30f;
      bhi[0] = tvalid ? tfx : -3.0e30f;
      bhi[1] = tvalid ? tfy : -3.0e30f;
      bhi[2] = tvalid ? tfz : -3.0e30f;
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
      TM(TCLK(tm_a); tm_setup += tm_a - tm_b;)
    }
    const int first_it = it;
    /* ---- stage loop: one state machine with a single drain() call site, so
     * that the test loop does not carry the drain's registers ---- */
    const float r2e = [&]() {
      const float re = fmaf(thg, PREFILTER_REL, A.margin);
      return re * re;
    }();
    int s = 0;      /* ring slot of the stage under test */
    unsigned m = 0; /* accepted octets of that stage still to test */
    bool done = false;
    for (;;) {
      if (m == 0) {
        /* ---- next stage ---- */
        s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        TM(TCLK(tm_a);)
        mbar_wait(sFull + s, ph);
        TM(TCLK(tm_b); tm_wait += tm_b - tm_a; tm_nwait++; if (it == 0) tm_first = tm_b - tm_t0;)
        const char *const st = smem + s * SM::kStageBytes;
        const int32_t *const meta = (const int32_t *)(st + SM::kStageMeta);
        if (meta[2] == 2 || (meta[2] == 1 && it != first_it)) {
          done = true; /* terminator, or the first stage of the next task: leave it where it is */
        } else {
          const int noct = meta[1];
          it++;
          held++;
          /* ---- cull: lane = octet ---- */
          bool acc = false;
          if (lane < noct) {
            const float *const wb = sBox + warp * 8; /* the warp's target box and reach */
            const float rmax = wb[6];
            const float4 lo = ((const float4 *)(st + SM::kStageOB))[2 * lane];
            const float4 hi = ((const float4 *)(st + SM::kStageOB))[2 * lane + 1];
            const TileItem &ii = ((const TileItem *)(st + SM::kStageIT))[((const uint8_t *)(st + SM::kStageO2F))[lane]];
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
          TM(TCLK(tm_a); tm_cull += tm_a - tm_b;)
        }
      }
      if (m) {
        /* ---- test the accepted octets (until done or a sub-list may overflow) ---- */
        TM(TCLK(tm_a);)
        const char *const st = smem + s * SM::kStageBytes;
        const float4 *const F = (const float4 *)(st + SM::kStageF);
        const uint8_t *const o2f = (const uint8_t *)(st + SM::kStageO2F);
        const TileItem *const IT = (const TileItem *)(st + SM::kStageIT);
        int cur = -1;
        float tpx = 3.0e30f, tpy = 0.f, tpz = 0.f;
        int nsub_ub = __reduce_max_sync(FULL_MASK, nsub);
        bool skip = true;
        while (m) {
          const int o = __ffs(m) - 1;
          const int fl = o2f[o];
          if (fl != cur) {
            cur = fl;
            const TileItem &ii = IT[fl];
            const bool part = tvalid && tdepth >= ii.min_depth && tdepth <= ii.max_depth;
            tpx = part ? tfx - ii.d[0] : 3.0e30f;
            tpy = tfy - ii.d[1];
            tpz = tfz - ii.d[2];
            skip = !__any_sync(FULL_MASK, part);
          }
          if (skip) {
            m &= m - 1u;
            continue;
          }
          if (nsub_ub > QCAP - 2) {
            nsub_ub = __reduce_max_sync(FULL_MASK, nsub);
            if (nsub_ub > QCAP - 2) break; /* drain first, then come back to this octet */
          }
          m &= m - 1u;
          nsub_ub += 2;
          const int sl = o * 8 + 2 * s4;
          const float4 a = F[sl], c = F[sl + 1];
          const int code = (fl << 11) | (s << 8) | sl;
          ntests++;
          {
            const float dx = tpx - a.x, dy = tpy - a.y, dz = tpz - a.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const bool ok = FORCE ? (r2 < fmaxf(r2e, a.w)) : (r2 < r2e);
            if (ok) {
              mylist[nsub * 32] = (uint16_t)code;
              nsub++;
            }
          }
          {
            const float dx = tpx - c.x, dy = tpy - c.y, dz = tpz - c.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const bool ok = FORCE ? (r2 < fmaxf(r2e, c.w)) : (r2 < r2e);
            if (ok) {
              mylist[nsub * 32] = (uint16_t)(code + 1);
              nsub++;
            }
          }
        }
        TM(TCLK(tm_b); tm_test += tm_b - tm_a;)
      }
      if (m != 0 || held >= hold_max || done) {
        TM(TCLK(tm_a);)
        drain();
        TM(TCLK(tm_b); tm_drain += tm_b - tm_a; tm_ndrain++;)
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
    nhit_all += nhit;
  }
  TM(TCLK(tm_b); if (lane == 0 && warp == 0 && (blockIdx.x % 64) == 5)
       printf("TM cta %d life %lld tasks %d peek %lld setup %lld wait %lld (%d) cull %lld test %lld drain %lld (%d) first %lld\n",
              (int)blockIdx.x, tm_b - tm_t0, tm_ntask, tm_peek, tm_setup, tm_wait, tm_nwait, tm_cull, tm_test, tm_drain,
              tm_ndrain, tm_first);)
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
