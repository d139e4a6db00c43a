/*
 * abi_time_integration.inl - the entry points of the time-integration rows of SURVEY 8f (included by
 * swiftgpu.cu, where the handle and the list / launch helpers live): struct xpart upload / download,
 * swiftgpu_run_drift (cell_drift_part), swiftgpu_run_kick (runner_do_kick1/2), swiftgpu_run_limiter
 * (runner_dosub_{self,pair}1_limiter). Kernels: kernels_drift.cuh, loops_pipe.cuh (LOOP_LIMITER).
 */
/* ---- drift on the device (kernels_drift.cuh) ---- */
extern "C" int swiftgpu_upload_xparts(swiftgpu_t *h, const swiftgpu_xpart_layout *layout, const void *xparts_aos,
                                      int64_t nparts) {
  if (!h || !layout || !xparts_aos || nparts <= 0) return 1;
  cudaSetDevice(h->cfg.device);
  const int64_t nh = h->n_host > 0 ? h->n_host : h->n;
  if (nparts != nh) return h->fail("upload_xparts: one xpart per uploaded part (upload the parts first)");
  if (layout->size <= 0 || layout->x_diff < 0 || layout->x_diff_sort < 0 || layout->v_full < 0 ||
      layout->x_diff + 12 > layout->size || layout->x_diff_sort + 12 > layout->size ||
      layout->v_full + 12 > layout->size)
    return h->fail("upload_xparts: bad struct xpart layout");
  if (layout->u_full >= 0 && layout->u_full + 4 > layout->size) return h->fail("upload_xparts: bad u_full offset");
  if (h->n_x != nparts || h->xlayout.size != layout->size) {
    cudaFree(h->d_xaos);
    h->d_xaos = nullptr;
    CK(cudaMalloc((void **)&h->d_xaos, (size_t)layout->size * (size_t)nparts));
    h->n_x = nparts;
  }
  h->xlayout = *layout;
  CK(cudaMemcpyAsync(h->d_xaos, xparts_aos, (size_t)layout->size * (size_t)nparts, cudaMemcpyHostToDevice,
                     h->stream));
  return 0;
}

extern "C" int swiftgpu_download_xparts(swiftgpu_t *h, void *xparts_aos, int64_t nparts) {
  if (!h || !xparts_aos || nparts != h->n_x || !h->d_xaos) return 1;
  cudaSetDevice(h->cfg.device);
  CK(cudaMemcpyAsync(xparts_aos, h->d_xaos, (size_t)h->xlayout.size * (size_t)nparts, cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.n_host_syncs++;
  return 0;
}

extern "C" int swiftgpu_run_drift(swiftgpu_t *h, const swiftgpu_drift_args *args) {
  if (!h || !args) return 1;
  cudaSetDevice(h->cfg.device);
  if (!h->d_xaos) return h->fail("run_drift before upload_xparts");
  if (!(args->dt_drift >= 0.)) return h->fail("run_drift: attempt to drift to the past");
  if (ensure_lists(h)) return 1; /* the device cell table */
  if (transpose_out(h)) return 1; /* a_hydro, h_dt, u_dt ... of the last step into the AoS copy */
  DriftArgs A;
  A.aos = h->d_aos;
  A.xaos = h->d_xaos;
  A.D.L = h->cfg.layout;
  A.D.scheme = h->cfg.scheme;
  A.X = h->xlayout;
  A.cells = h->d_cells;
  A.ncells = h->ncells;
  A.d2h = h->d_d2h;
  A.dt_drift = args->dt_drift;
  A.dt_kick_hydro = args->dt_kick_hydro;
  A.dt_therm = args->dt_therm;
  A.min_u = args->minimal_internal_energy; /* / cosmo->a_factor_internal_energy = 1 */
  A.h_max = h->cfg.h_max;
  A.h_min = h->cfg.h_min;
  A.init_particles = args->init_particles;
  A.max_active_bin = h->step.max_active_bin;
  A.n_host = h->n_x;
  const int nc = h->ncells;
  k_drift_begin<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_cells, nc);
  const unsigned grid = (unsigned)(((int64_t)nc * 32 + 127) / 128);
  if (h->cfg.scheme == SCH_MINIMAL) k_drift<SCH_MINIMAL><<<grid, 128, 0, h->stream>>>(A);
  else if (h->cfg.scheme == SCH_GADGET2) k_drift<SCH_GADGET2><<<grid, 128, 0, h->stream>>>(A);
  else k_drift<SCH_SPHENIX><<<grid, 128, 0, h->stream>>>(A);
  float *d_tmp = nullptr;
  CK(cudaMalloc((void **)&d_tmp, 4 * sizeof(float) * (size_t)nc));
  k_get_cell_drift<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_cells, nc, d_tmp, h->d_dxp);
  /* the drifted cell table is the state every following step starts from (run_density restores it) */
  CK(cudaMemcpyAsync(h->d_cells_init, h->d_cells, sizeof(DevCell) * (size_t)nc, cudaMemcpyDeviceToDevice, h->stream));
  /* do the worklists survive? the density / subset lists were flattened with the predicates
   * cell.h:951,992 on h_max_active (the force list is revalidated by run_force in every step) */
  CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int32_t), h->stream));
  k_pred_bits<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_cells, h->d_dmin, h->d_dxp_old, h->d_loop1_bits, nc, 1,
                                                     h->d_flag);
  h->stats.n_launches += 4;
  std::vector<float> v(4 * (size_t)nc);
  int32_t flag = 0;
  cudaError_t e = cudaMemcpyAsync(v.data(), d_tmp, sizeof(float) * v.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, h->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_tmp);
  if (e != cudaSuccess) return h->fail(cudaGetErrorString(e));
  h->stats.n_host_syncs++;
  /* the host mirror of the cells follows: a later list rebuild flattens the recursion with it */
  for (int c = 0; c < nc; c++) {
    swiftgpu_cell &C = h->cells[c];
    if (C.nodeID != h->cfg.rank && h->cfg.nranks > 1) continue;
    if (C.count == 0) continue;
    C.h_max = v[c];
    C.h_max_active = v[(size_t)nc + c];
    C.dx_max_part = v[2 * (size_t)nc + c];
    C.dx_max_sort = v[3 * (size_t)nc + c];
    h->up_hmax[c] = C.h_max;
    h->up_hmax_active[c] = C.h_max_active;
  }
  if (flag) h->lists_built = false;
  /* device order, SoA columns, frames: from the drifted AoS copy */
  return transpose_in(h);
}

extern "C" int swiftgpu_run_kick(swiftgpu_t *h, int which, float minimal_internal_energy) {
  if (!h || (which != 1 && which != 2)) return 1;
  cudaSetDevice(h->cfg.device);
  if (!h->d_xaos) return h->fail("run_kick before upload_xparts");
  if (h->xlayout.u_full < 0) return h->fail("run_kick: the xpart layout has no u_full / entropy_full offset");
  if (!h->has_step) return h->fail("swiftgpu_set_step must be called before run_kick");
  if (transpose_out(h)) return 1; /* a_hydro, u_dt | entropy_dt of the last step into the AoS copy */
  KickArgs A;
  A.aos = h->d_aos;
  A.xaos = h->d_xaos;
  A.D.L = h->cfg.layout;
  A.D.scheme = h->cfg.scheme;
  A.X = h->xlayout;
  A.n = h->n_x;
  A.which = which;
  A.max_active_bin = h->step.max_active_bin;
  A.time_base = h->step.time_base;
  A.min_u = minimal_internal_energy;
  const unsigned grid = (unsigned)((A.n + 255) / 256);
  if (h->cfg.scheme == SCH_MINIMAL) k_kick<SCH_MINIMAL><<<grid, 256, 0, h->stream>>>(A);
  else if (h->cfg.scheme == SCH_GADGET2) k_kick<SCH_GADGET2><<<grid, 256, 0, h->stream>>>(A);
  else k_kick<SCH_SPHENIX><<<grid, 256, 0, h->stream>>>(A);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  /* the SoA columns follow the AoS copy: v, u and the force members changed (kick2); positions did
   * not move, so the device order, frames and lists stay */
  if (which == 2) return transpose_in(h);
  /* kick1 leaves struct part untouched except a zeroed rate at the energy floor; the results of the
   * last step are already in the AoS copy, which is now the current state */
  h->phases_done = 0;
  return 0;
}

/* ---- the time-step limiter loop (runner_doiact_limiter.h) ---- */
__global__ void k_limiter_io(char *aos, int part_size, int off, const int32_t *d2h, int64_t n, int64_t n_host,
                             int32_t *wakeup, int store) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t row = d2h[p];
  if (row >= n_host) {
    if (!store) wakeup[p] = -56; /* time_bin_not_awake, timeline.h:48 */
    return;
  }
  char *b = aos + (size_t)part_size * (size_t)row;
  if (store)
    *(int8_t *)(b + off) = (int8_t)wakeup[p];
  else
    wakeup[p] = *(const int8_t *)(b + off);
}

/* runner_dosub_{self,pair}1_limiter (runner_main.c:233,292 -> runner_doiact_functions_limiter.h): the
 * density decomposition (cell.h:951,992 on the h_max_active the ghost left), targets = the particles
 * starting their step, r2 < h_i^2 gamma^2, and runner_iact_nonsym_limiter: a neighbour more than
 * time_bin_neighbour_max_delta_bin bins above the target gets limiter_data.wakeup =
 * max(wakeup, -time_bin_i) - a scatter, done with atomicMax. wakeup_offset =
 * offsetof(struct part, limiter_data.wakeup); the time bins are those of the last upload. */
extern "C" int swiftgpu_run_limiter(swiftgpu_t *h, int32_t wakeup_offset) {
  if (!h) return 1;
  cudaSetDevice(h->cfg.device);
  if (wakeup_offset < 0 || wakeup_offset >= h->cfg.layout.size) return h->fail("run_limiter: bad wakeup offset");
  if (h->cfg.nranks > 1) return h->fail("run_limiter: single rank only (woken-up proxies are not sent back)");
  if (loop_kind() != 3) return h->fail("run_limiter needs the frame pipeline");
  if (!(h->phases_done & SWIFTGPU_PHASE_GHOST)) return h->fail("run_limiter before run_ghost");
  if (phase_begin(h)) return 1;
  if (revalidate_list(h, LISTS_GRADIENT)) return 1; /* the loop-1 predicates on the post-ghost h_max_active */
  DevList &L = h->gradient_own ? h->L_gradient : h->L_density;
  if (build_targets(h, L)) return 1;
  const int64_t n = h->n, nh = h->n_host > 0 ? h->n_host : h->n;
  if (!h->d_wakeup) CK(cudaMalloc((void **)&h->d_wakeup, sizeof(int32_t) * (size_t)n));
  const unsigned grid = (unsigned)((n + 255) / 256);
  k_limiter_io<<<grid, 256, 0, h->stream>>>(h->d_aos, h->cfg.layout.size, wakeup_offset, h->d_d2h, n, nh,
                                          h->d_wakeup, 0);
  h->stats.n_launches++;
  CK(cudaMemsetAsync(h->d_counters + 11, 0, sizeof(unsigned long long), h->stream));
  if (L.ntasks > 0) {
    /* counter slot 3: the loop credits no interaction (total untouched), its distance tests go to slot 11 */
    if (run_pipe_loop<LOOP_LIMITER, false, 0>(h, L, nullptr, 3, h->d_counters + 12, 0, main_split(), h->d_counters + 13))
      return 1;
  }
  k_limiter_io<<<grid, 256, 0, h->stream>>>(h->d_aos, h->cfg.layout.size, wakeup_offset, h->d_d2h, n, nh,
                                          h->d_wakeup, 1);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  return 0;
}

