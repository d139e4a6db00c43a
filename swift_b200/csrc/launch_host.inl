/*
 * launch_host.inl - launching the neighbour loops (included by swiftgpu.cu): LoopArgs of a list,
 * target lists and task records of a launch, the kernel variants of the frame pipeline (ring depths,
 * small-task variant, device-side choice between them: run_pipe_loop), the direct kernel, and - in
 * the legacy build only - the superseded loop generations.
 */
static LoopArgs loop_args(H *h, const DevList &D, int32_t *count, int counter) {
  LoopArgs A;
  memset(&A, 0, sizeof(A));
  A.cells = h->d_cells;
  A.items = D.items;
  A.groups = D.groups;
  A.task_group = D.task_group;
  A.task_chunk = D.task_chunk;
  A.ntasks = D.ntasks;
  A.tgt_list = D.tgt_list;
  A.tgt_first = D.tgt_first;
  A.tgt_count = D.tgt_count;
  A.sort_idx = h->sort_idx;
  A.ext = h->d_ext;
  A.x = h->x; A.mv = h->mv; A.h = h->hh; A.depth_h = h->depth_h; A.time_bin = h->time_bin;
  A.fq1 = h->fq1; A.fq2 = h->fq2; A.fq3 = h->fq3;
  A.xf = h->xf; A.xs0 = h->xs; A.xs1 = h->xs + (h->n + 4); A.xs2 = h->xs + 2 * (h->n + 4); A.gq = h->gq; A.boxes = h->boxes; A.cell_box_first = h->d_box_first;
  A.keyE = tile_keyE(h);
  A.margin = tile_margin(h);
  A.task_counter = (unsigned int *)(h->d_counters + 14);
  A.frames = h->d_frames;
  A.task_recs = D.task_recs;
  A.ntask_dev = (const unsigned int *)(h->d_counters + 15);
  {
    static int hold = -1;
    if (hold < 0) {
      const char *e = getenv("SWIFTGPU_HOLD");
      hold = e ? atoi(e) : (loop_kind() == 3 ? 0 : 2); /* tile: stages held before a drain; pipe: debug bits */
    }
    A.hold = hold;
  }
  A.dA = h->dA; A.dB = h->dB; A.g_vsig = h->g_vsig; A.g_lap = h->g_lap; A.g_amax = h->g_amax;
  A.fo1 = h->fo1; A.f_hdt = h->f_hdt; A.f_vsig = h->f_vsig; A.f_minngb = h->f_minngb;
  A.count = count;
  A.wakeup = h->d_wakeup;
  A.total = h->d_counters + counter;
  A.tests = h->d_counters + 8 + counter;
  for (int k = 0; k < 3; k++) A.dim[k] = h->cfg.dim[k];
  A.a2_Hubble = h->step.a * h->step.a * h->step.H;
  A.max_active_bin = h->step.max_active_bin;
  return A;
}

/* sparse_out: fewer than SWIFTGPU_SPARSE targets per non-empty task on average */
static int sparse_threshold() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SWIFTGPU_SPARSE");
    v = e ? atoi(e) : 28;
  }
  return v;
}
static int build_targets(H *h, DevList &D, bool *sparse_out = nullptr) {
  if (sparse_out) *sparse_out = false;
  if (D.ngroups == 0) return 0;
  CK(cudaMemsetAsync(h->d_counters + 12, 0, 2 * sizeof(unsigned long long), h->stream));
  k_build_targets<<<(D.ngroups * 32 + 127) / 128, 128, 0, h->stream>>>(
      D.groups, D.ngroups, h->d_cells, h->time_bin, h->step.max_active_bin, D.tgt_first, D.tgt_count,
      D.tgt_list, D.items, h->depth_h, h->d_counters + 12);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  if (sparse_out && loop_kind() == 2) {
    unsigned long long t[2] = {0, 0};
    CK(cudaMemcpyAsync(t, h->d_counters + 12, sizeof(t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->stats.n_host_syncs++;
    *sparse_out = t[1] > 0 && t[0] < (unsigned long long)sparse_threshold() * t[1];
  }
  return 0;
}

/* The compacted TaskRecs of one launch of the frame pipeline, from the target lists as they are NOW
 * (no host round trip: the kernel reads the number of tasks from device memory). */
static int build_task_recs(H *h, const DevList &D, int chunk, const unsigned long long *gate = nullptr,
                           unsigned long long gate_lo = 0, unsigned long long gate_hi = ~0ull,
                           const unsigned long long *gate_den = nullptr) {
  CK(cudaMemsetAsync(h->d_counters + 15, 0, sizeof(unsigned long long), h->stream));
  if (D.ntasks == 0) return 0;
  const int64_t n = h->n;
  k_task_recs<<<(unsigned)(((int64_t)D.ntasks * 32 + 127) / 128), 128, 0, h->stream>>>(
      D.task_group, D.task_chunk, D.ntasks, D.groups, h->d_cells, D.tgt_first, D.tgt_count, D.tgt_list, h->xs,
      h->xs + (n + 4), h->xs + 2 * (n + 4), h->hh, D.task_recs, (unsigned int *)(h->d_counters + 15), gate, gate_lo,
      gate_hi, chunk, gate_den);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  return 0;
}

static int read_counter(H *h, int k, int64_t *out) {
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, h->d_counters + k, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *out = (int64_t)v;
  return 0;
}

/* CTA-cooperative type-1 loops (loops_cta.cuh); SWIFTGPU_WARP_LOOPS=1 selects the
 * warp-private kernels of loops.cuh instead (kept for A/B measurements). */
/* SWIFTGPU_LOOPS=pipe (default: frame pipeline, loops_pipe.cuh) | tile (loops_tile.cuh) | cta
 * (loops_cta.cuh) | warp (loops.cuh); the older kernels are kept for A/B measurements. */
static int loop_kind() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SWIFTGPU_LOOPS");
    const char *w = getenv("SWIFTGPU_WARP_LOOPS");
    v = 3;
#ifdef SWIFTGPU_LEGACY_LOOPS
    if (e && !strcmp(e, "tile")) v = 2;
    if (e && !strcmp(e, "cta")) v = 1;
    if (e && !strcmp(e, "warp")) v = 0;
    if (w && w[0] == '1') v = 0;
#else
    (void)e;
    (void)w;
#endif
  }
  return v;
}
static bool use_cta_loops() { return loop_kind() >= 1; }
#ifdef SWIFTGPU_LEGACY_LOOPS
#ifndef TL_FORCE_NS
#define TL_FORCE_NS 4 /* ring stages of the force kernel (SPHENIX: one less, 4 payload columns) */
#endif
template <int LOOP, int SCHEME, int CW>
static cudaError_t launch_tile_cw(H *h, const LoopArgs &A) {
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  /* ring stages: as many as keep 3 (type-1) / 2 (force) standard CTAs, or 5 small CTAs, on an SM */
  constexpr int NS = CW == 8 ? (FORCE ? (SCHEME == SCH_SPHENIX ? TL_FORCE_NS - 1 : TL_FORCE_NS) : (LOOP == LOOP_GRADIENT ? 3 : TL_DENS_NS)) : (FORCE ? 2 : TL_SPARSE_NS);
  constexpr int bytes = TileSmem<NP, NS, (FORCE ? TL_SUBCAP2 : TL_SUBCAP1), CW>::kBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_tile<LOOP, SCHEME, NS, CW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  /* persistent CTAs: as many as are resident at once, tasks drawn from a counter */
  static int resident = 0;
  if (!resident) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile<LOOP, SCHEME, NS, CW>,
                                                                  32 * (CW + 1), bytes);
    if (e != cudaSuccess) return e;
    resident = std::max(1, per_sm) * std::max(1, sms);
  }
  const long long ncta = (long long)A.ntasks * (TL_CWARPS / CW);
  const int grid = (int)std::min<long long>(ncta, resident);
  cudaError_t e = cudaMemsetAsync(A.task_counter, 0, sizeof(unsigned int), h->stream);
  if (e != cudaSuccess) return e;
  k_tile<LOOP, SCHEME, NS, CW><<<grid, 32 * (CW + 1), bytes, h->stream>>>(A);
  return cudaGetLastError();
}
/* sparse: few targets per group (late ghost iterations): 4-consumer-warp CTAs, 5 per SM */
template <int LOOP, int SCHEME>
static cudaError_t launch_tile(H *h, const LoopArgs &A, bool sparse = false) {
  if (sparse) return launch_tile_cw<LOOP, SCHEME, 4>(h, A);
  return launch_tile_cw<LOOP, SCHEME, 8>(h, A);
}
#endif /* SWIFTGPU_LEGACY_LOOPS */

/* sparse target sets (late ghost iterations): one warp per target, loops_direct.cuh */
static int launch_direct_density(H *h, const DevList &D, const LoopArgs &A, const unsigned long long *gate = nullptr,
                                 unsigned long long gate_hi = ~0ull) {
  if (!h->d_flat_tgt) CK(cudaMalloc((void **)&h->d_flat_tgt, sizeof(int2) * (size_t)std::max<int64_t>(h->n, 1)));
  unsigned int *nflat = (unsigned int *)(h->d_counters + 13);
  CK(cudaMemsetAsync(nflat, 0, sizeof(unsigned long long), h->stream));
  if (D.ngroups == 0) return 0;
  k_flat_targets<<<(D.ngroups * 32 + 127) / 128, 128, 0, h->stream>>>(D.groups, D.ngroups, D.tgt_first, D.tgt_count,
                                                                     D.tgt_list, h->d_flat_tgt, nflat, gate, gate_hi);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  k_direct<LOOP_DENSITY><<<sms * 8, 256, 0, h->stream>>>(A, h->d_flat_tgt, nflat);
  h->stats.n_launches += 2;
  CK(cudaGetLastError());
  return 0;
}

/* frame pipeline (loops_pipe.cuh): DS = double-column slots per stage (64: only the self item of a
 * main loop is evaluated on doubles; 256: the ghost re-runs, where every pair item is) */
template <int LOOP, int SCHEME, int NS, int DS, int CW = 8, int SL = PL_SLOTS>
static cudaError_t launch_pipe_ns(H *h, const LoopArgs &A) {
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  constexpr int bytes = PipeSmem<NP, NS, (FORCE ? TL_SUBCAP2 : TL_SUBCAP1), CW, DS, SL>::kBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_pipe<LOOP, SCHEME, NS, CW, DS, SL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  static int resident = 0;
  if (!resident) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pipe<LOOP, SCHEME, NS, CW, DS, SL>,
                                                                  32 * (CW + 1), bytes);
    if (e != cudaSuccess) return e;
    resident = std::max(1, per_sm) * std::max(1, sms);
    if (getenv("SWIFTGPU_VERBOSE"))
      fprintf(stderr, "k_pipe<%d,%d,NS=%d,CW=%d,DS=%d,SL=%d>: %d B smem, %d CTAs/SM\n", LOOP, SCHEME, NS, CW, DS, SL, bytes, per_sm);
  }
  const int grid = (int)std::min<long long>(A.ntasks, resident);
  if (grid <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(A.task_counter, 0, sizeof(unsigned int), h->stream);
  if (e != cudaSuccess) return e;
  k_pipe<LOOP, SCHEME, NS, CW, DS, SL><<<grid, 32 * (CW + 1), bytes, h->stream>>>(A);
  return cudaGetLastError();
}
#ifndef PL_NS_DENSITY
/* ring depth = what fits 2 CTAs per SM: the consumer warps of a CTA need different stages (each
 * culls against its own 8 targets), so the fastest runs ahead of the slowest by up to the ring depth;
 * a deeper ring is what removes the full-barrier waits (ncu: 17-27 % of the samples at 4 stages) */
#define PL_NS_DENSITY 7
#define PL_NS_SUBSET 5
#define PL_NS_GRADIENT 5
#define PL_NS_FORCE 4
/* SPHENIX force: 5 float4 per source, 256-slot stages fit 3 times only. 192-slot stages x 4 were
 * measured: -7 % before the direction-balanced item order, +-0 after it (sweeps in profiles/r02_sweeps.log) */
#define PL_NS_FORCE_SPHENIX 3
#define PL_SLOTS_FORCE_SPHENIX 256
#endif
#ifndef PL_NS_SPARSE
/* the small-task variant (PL_SPARSE_CW consumer warps): 3-4 CTAs per SM */
#define PL_NS_SPARSE 3
#define PL_NS_SPARSE_FORCE 2
#endif
/* small = the variant with tasks of 8 * PL_SPARSE_CW targets: per task every consumer warp walks all
 * stages of the group's sources whatever the number of its targets, so a sparse target set (ghost
 * re-runs, few active particles) is served by fewer warps per task and more tasks in flight */
template <int LOOP, bool SUBSET, int SCHEME>
static cudaError_t launch_pipe(H *h, const LoopArgs &A, bool small = false) {
  if (LOOP == LOOP_LIMITER)
    return small ? launch_pipe_ns<LOOP_LIMITER, 0, PL_NS_SPARSE, 64, PL_SPARSE_CW>(h, A)
                 : launch_pipe_ns<LOOP_LIMITER, 0, PL_NS_DENSITY, 64>(h, A);
  if (small) {
    if (LOOP == LOOP_FORCE) return launch_pipe_ns<LOOP_FORCE, SCHEME, PL_NS_SPARSE_FORCE, 64, PL_SPARSE_CW>(h, A);
    if (LOOP == LOOP_GRADIENT) return launch_pipe_ns<LOOP_GRADIENT, 0, PL_NS_SPARSE, 64, PL_SPARSE_CW>(h, A);
    if (SUBSET) return launch_pipe_ns<LOOP_DENSITY, 0, PL_NS_SPARSE, 256, PL_SPARSE_CW>(h, A);
    return launch_pipe_ns<LOOP_DENSITY, 0, PL_NS_SPARSE, 64, PL_SPARSE_CW>(h, A);
  }
  if (LOOP == LOOP_FORCE && SCHEME == SCH_SPHENIX)
    return launch_pipe_ns<LOOP_FORCE, SCHEME, PL_NS_FORCE_SPHENIX, 64, 8, PL_SLOTS_FORCE_SPHENIX>(h, A);
  if (LOOP == LOOP_FORCE) return launch_pipe_ns<LOOP_FORCE, SCHEME, PL_NS_FORCE, 64>(h, A);
  if (LOOP == LOOP_GRADIENT) return launch_pipe_ns<LOOP_GRADIENT, 0, PL_NS_GRADIENT, 64>(h, A);
  if (SUBSET) return launch_pipe_ns<LOOP_DENSITY, 0, PL_NS_SUBSET, 256>(h, A);
  return launch_pipe_ns<LOOP_DENSITY, 0, PL_NS_DENSITY, 64>(h, A);
}
/* Fraction of a list's potential targets below which the small-task variant takes the launch
 * (SWIFTGPU_SPARSE_FRAC, A/B knob; 0 = never, > 1 = always). The choice is made ON THE DEVICE: both
 * variants are enqueued, k_task_recs of the one whose gate is closed emits no task. */
static double sparse_frac() {
  static double v = -1.;
  if (v < 0.) {
    const char *e = getenv("SWIFTGPU_SPARSE_FRAC");
    v = e ? atof(e) : 0.4;
  }
  return v;
}

/* One neighbour loop of the frame pipeline over the targets list D holds NOW. Both task sizes are
 * enqueued; which one finds tasks is decided on the device: `gate` (targets of the list, or
 * unconverged particles) in [lo, split) -> small tasks, [split, inf) -> 64-target tasks; below lo the
 * caller's direct kernel. With `den` the bounds are per 64-target chunk (the mean fill of the tasks). */
template <int LOOP, bool SUBSET, int SCHEME>
static int run_pipe_loop(H *h, DevList &D, int32_t *counts, int counter_slot, const unsigned long long *gate,
                         unsigned long long lo, unsigned long long split, const unsigned long long *den) {
  if (ensure_frames(h)) return 1;
  if (split != ~0ull) {
    if (build_task_recs(h, D, PL_TARGETS, gate, std::max(lo, split), ~0ull, den)) return 1;
    LoopArgs A = loop_args(h, D, counts, counter_slot);
    CK((launch_pipe<LOOP, SUBSET, SCHEME>(h, A, false)));
    h->stats.n_launches++;
  }
  if (split > lo) {
    if (build_task_recs(h, D, 8 * PL_SPARSE_CW, gate, lo, split, den)) return 1;
    LoopArgs A = loop_args(h, D, counts, counter_slot);
    CK((launch_pipe<LOOP, SUBSET, SCHEME>(h, A, true)));
    h->stats.n_launches++;
  }
  return 0;
}
/* main loops: split on the mean number of targets per 64-target chunk of the list (k_build_targets' totals) */
static unsigned long long main_split() {
  const double f = sparse_frac();
  if (f <= 0.) return 0ull;
  if (f > 1.) return ~0ull;
  return (unsigned long long)(f * TASK_TARGETS + 0.5);
}

#ifdef SWIFTGPU_LEGACY_LOOPS
template <int LOOP, bool SUBSET, int SCHEME>
static cudaError_t launch_cta(H *h, const LoopArgs &A) {
  constexpr bool FORCE = (LOOP == LOOP_FORCE);
  constexpr int NP = FORCE ? (SCHEME == SCH_SPHENIX ? 4 : 3) : (LOOP == LOOP_GRADIENT ? 2 : 1);
  constexpr int bytes = CtaSmem<NP, SUBSET, FORCE>::kBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_cta<LOOP, SUBSET, SCHEME>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_cta<LOOP, SUBSET, SCHEME><<<A.ntasks, CTA_THREADS, bytes, h->stream>>>(A);
  return cudaGetLastError();
}
#endif
template <int LOOP, bool SUBSET>
static cudaError_t launch_loop1(H *h, const LoopArgs &A, bool sparse = false) {
  if (loop_kind() == 3) return launch_pipe<LOOP, SUBSET, 0>(h, A);
#ifdef SWIFTGPU_LEGACY_LOOPS
  if (loop_kind() == 2) return launch_tile<LOOP, 0>(h, A, sparse);
  if (use_cta_loops()) return launch_cta<LOOP, SUBSET, 0>(h, A);
  k_loop1<LOOP, SUBSET><<<A.ntasks, 32, Tile1<LOOP, SUBSET>::kBytes, h->stream>>>(A);
  return cudaGetLastError();
#else
  (void)sparse;
  return cudaErrorNotSupported;
#endif
}
template <int SCHEME>
static cudaError_t launch_loop2(H *h, const LoopArgs &A, bool sparse = false) {
  if (loop_kind() == 3) return launch_pipe<LOOP_FORCE, false, SCHEME>(h, A);
#ifdef SWIFTGPU_LEGACY_LOOPS
  if (loop_kind() == 2) return launch_tile<LOOP_FORCE, SCHEME>(h, A, sparse);
  if (use_cta_loops()) return launch_cta<LOOP_FORCE, false, SCHEME>(h, A);
  k_loop2<SCHEME><<<A.ntasks, 32, Tile2<SCHEME>::kBytes, h->stream>>>(A);
  return cudaGetLastError();
#else
  (void)sparse;
  return cudaErrorNotSupported;
#endif
}

