/*
 * libswiftgpu - B200-native SPH neighbour-interaction path for SWIFT.
 *
 * Public C ABI. Plain C types only (no CUDA, no torch types): this is the file
 * a SWIFT maintainer includes from runner_main.c / engine.c, or binds through
 * ctypes/cffi. Every entry point names the reference interface it stands in
 * for (paths relative to the SWIFT 2026.04 source tree).
 *
 * The reference has no run-time plugin interface; hydro schemes are chosen by
 * #define (src/hydro.h:31-93) and the loops are called from the task switch in
 * runner_main (src/runner_main.c:214-375). The seam is therefore "one call per
 * task TYPE over the whole active set" instead of one call per task:
 *
 *   task_type_sort                      -> swiftgpu_run_sort
 *   task_type_{self,pair}/density       -> swiftgpu_run_density
 *   task_type_ghost                     -> swiftgpu_run_ghost
 *   task_type_{self,pair}/gradient      -> swiftgpu_run_gradient      (SPHENIX)
 *   task_type_extra_ghost               -> swiftgpu_run_extra_ghost   (SPHENIX)
 *   task_type_{self,pair}/force         -> swiftgpu_run_force
 *   task_type_end_hydro_force           -> swiftgpu_run_end_force
 *
 * All functions return 0 on success and non-zero on failure; the library never
 * calls abort()/exit() (the reference's error() macro, src/error.h:60-80, is the
 * caller's job). swiftgpu_last_error() returns the message of the last failure.
 * One handle drives one GPU; calls on one handle must be serialised.
 */
#ifndef SWIFTGPU_H
#define SWIFTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWIFTGPU_ABI_VERSION 2

/* --with-hydro=minimal|gadget2|sphenix (configure.ac:2233-2300). */
enum swiftgpu_scheme {
  SWIFTGPU_SCHEME_MINIMAL = 0,
  SWIFTGPU_SCHEME_GADGET2 = 1,
  SWIFTGPU_SCHEME_SPHENIX = 2
};

/* Task-type bits for swiftgpu_run_step() / the oracle driver. */
enum swiftgpu_phase {
  SWIFTGPU_PHASE_SORT = 1 << 0,        /* runner_do_hydro_sort, runner_sort.c:203 */
  SWIFTGPU_PHASE_DENSITY = 1 << 1,     /* runner_dosub_{self,pair}1_density */
  SWIFTGPU_PHASE_GHOST = 1 << 2,       /* runner_do_ghost, runner_ghost.c:1113 */
  SWIFTGPU_PHASE_GRADIENT = 1 << 3,    /* runner_dosub_{self,pair}1_gradient */
  SWIFTGPU_PHASE_EXTRA_GHOST = 1 << 4, /* runner_do_extra_ghost, runner_ghost.c:1016 */
  SWIFTGPU_PHASE_FORCE = 1 << 5,       /* runner_dosub_{self,pair}2_force */
  SWIFTGPU_PHASE_END_FORCE = 1 << 6,   /* runner_do_end_hydro_force, runner_others.c:815 */
  SWIFTGPU_PHASE_ALL = 0x7f
};

/*
 * Byte offsets of the fields of the host's `struct part` (AoS) that the path
 * reads or writes. The three schemes lay the record out differently
 * (hydro/Minimal/hydro_part.h:103-232, Gadget2/hydro_part.h:91-236,
 * SPHENIX/hydro_part.h:105-324) and debugging options change it again, so the
 * caller passes offsetof() values; -1 marks a field the scheme does not have.
 * swiftgpu_default_layout() returns the layout of the default build.
 */
typedef struct swiftgpu_part_layout {
  int32_t size; /* sizeof(struct part) */
  int32_t id, x, v, a_hydro, mass, h;
  int32_t u, u_dt;             /* Minimal, SPHENIX */
  int32_t entropy, entropy_dt; /* Gadget2 */
  int32_t rho;
  /* union { density; force } */
  int32_t wcount, wcount_dh, rho_dh, rot_v;
  int32_t div_v; /* density.div_v (Minimal, Gadget2) or viscosity.div_v (SPHENIX) */
  int32_t f, pressure /* Minimal, SPHENIX */, P_over_rho2 /* Gadget2 */, soundspeed;
  int32_t v_sig; /* force.v_sig (Minimal, Gadget2) or viscosity.v_sig (SPHENIX) */
  int32_t h_dt, balsara;
  /* SPHENIX only */
  int32_t div_v_dt, div_v_previous_step, visc_alpha, laplace_u, diff_alpha,
      alpha_visc_max_ngb;
  int32_t time_bin, depth_h;
  int32_t min_ngb_time_bin; /* limiter_data.min_ngb_time_bin */
} swiftgpu_part_layout;

/*
 * The scalars the loops read from struct engine / struct space /
 * struct hydro_props / struct cosmology (SURVEY section 5 "Config" row).
 */
typedef struct swiftgpu_config {
  int32_t abi_version; /* SWIFTGPU_ABI_VERSION */
  int32_t scheme;      /* enum swiftgpu_scheme */
  int32_t device;      /* CUDA device ordinal */
  int32_t periodic;    /* space->periodic */
  double dim[3];       /* space->dim */

  /* struct hydro_props (hydro_properties.c:60-215) */
  float eta_neighbours;
  float h_tolerance;
  float h_max;
  float h_min;
  int32_t max_smoothing_iterations;
  int32_t use_mass_weighted_num_ngb;
  float CFL_condition;
  /* SPHENIX viscosity / diffusion (hydro/SPHENIX/hydro_parameters.h:53-176) */
  float viscosity_alpha, viscosity_alpha_max, viscosity_alpha_min,
      viscosity_length;
  float diffusion_alpha, diffusion_beta, diffusion_alpha_max,
      diffusion_alpha_min;

  /* domain decomposition: this rank and the number of ranks (engine->nodeID);
   * cells with nodeID != rank are foreign: read, never updated. */
  int32_t rank, nranks;

  swiftgpu_part_layout layout;
} swiftgpu_config;

/* Per-step scalars (struct engine / struct cosmology). */
typedef struct swiftgpu_step {
  int64_t ti_current;     /* engine->ti_current */
  int32_t max_active_bin; /* engine->max_active_bin (active.h:349) */
  int32_t with_cosmology; /* must be 0 in this version (and a = 1, H = 0) */
  double time_base;       /* engine->time_base (dt = 2^(bin+1) * time_base... timeline.h:91) */
  float a, H;             /* cosmology->a, cosmology->H (1, 0 without cosmology) */
} swiftgpu_step;

/*
 * One node of the flattened cell tree: the fields of struct cell
 * (cell.h:372-529) and struct cell_hydro (cell_hydro.h:34-177) the path reads.
 * Cells are referred to by index into the array passed to
 * swiftgpu_upload_cells(); -1 = NULL. Particles of a cell are the contiguous
 * range [first_part, first_part+count) of the particle array and a split
 * cell's progeny partition that range in progeny order, exactly like
 * c->hydro.parts in the reference.
 */
typedef struct swiftgpu_cell {
  double loc[3];
  double width[3];
  float dmin;
  float h_min_allowed;
  float h_max_allowed;
  float h_max;        /* hydro.h_max */
  float h_max_active; /* hydro.h_max_active */
  float h_max_old;    /* hydro.h_max_old */
  float dx_max_part;
  float dx_max_part_old;
  float dx_max_sort;
  float dx_max_sort_old;
  int32_t depth;
  int32_t split;
  int32_t parent;
  int32_t progeny[8];
  int32_t nodeID;
  int32_t top;   /* index of the top-level ancestor (itself for depth 0) */
  int32_t count; /* hydro.count */
  int64_t first_part; /* hydro.parts - space->parts */
  int64_t ti_end_min; /* hydro.ti_end_min: cell active iff == ti_current */
} swiftgpu_cell;

typedef struct swiftgpu_handle swiftgpu_t;

/* Timing / counters of the last swiftgpu_run_* calls (device time, CUDA events). */
typedef struct swiftgpu_stats {
  double ms_sort, ms_density, ms_ghost, ms_gradient, ms_extra_ghost, ms_force,
      ms_end_force;
  int64_t n_density;  /* directed density interactions (runner_iact_nonsym_density calls) */
  int64_t n_gradient; /* directed gradient interactions */
  int64_t n_force;    /* directed force interactions */
  int64_t n_launches; /* kernel launches by the library */
  /* distance tests (candidate pairs examined) of the same loops: the
   * pruning efficiency is n_x / t_x */
  int64_t t_density, t_gradient, t_force;
  int32_t ghost_iterations;
  int32_t ghost_unconverged;
  /* worklists rebuilt because the ghost changed a recursion predicate
   * (cell.h:951,992 for the gradient loop, :966,1007 for the force loop) */
  int32_t force_list_rebuilds, gradient_list_rebuilds;
  int64_t n_host_syncs; /* cudaStreamSynchronize / event waits issued by the library */
} swiftgpu_stats;

/* Fills `out` with the struct part layout of the reference's default build of
 * `scheme` (no debugging options, *_NONE sub-grid models). */
int swiftgpu_default_layout(int scheme, swiftgpu_part_layout *out);

/* Fills cfg with the defaults of hydro_props_init (hydro_properties.c:41-48,
 * hydro/SPHENIX/hydro_parameters.h:53-101) and the default layout. */
int swiftgpu_default_config(int scheme, swiftgpu_config *cfg);

/* Creates a handle bound to cfg->device. Stands in for the scheme selection
 * done at configure time + hydro_props_init + the engine fields of
 * tests/test125cells.c:570-610. */
int swiftgpu_init(swiftgpu_t **h, const swiftgpu_config *cfg);
void swiftgpu_destroy(swiftgpu_t *h);
const char *swiftgpu_last_error(const swiftgpu_t *h);

/* Upload the flattened tree (space->cells_top + progeny, cell.h:372). `top`
 * lists the indices of the top-level cells in space->cells_top order. */
int swiftgpu_upload_cells(swiftgpu_t *h, const swiftgpu_cell *cells,
                          int32_t ncells, const int32_t *top, int32_t ntop);

/* Upload space->parts: raw AoS `struct part[nparts]` in the layout given at
 * init. A device kernel transposes it to the SoA columns the loops read. */
int swiftgpu_upload_parts(swiftgpu_t *h, const void *parts_aos, int64_t nparts);

/* Same with the AoS array already resident in device memory (bench "value"). */
int swiftgpu_upload_parts_device(swiftgpu_t *h, const void *d_parts_aos,
                                 int64_t nparts);

/* Multi-rank: only the rank's OWN particles cross the host boundary. The host's array holds them in
 * [0, nlocal) and the proxies of the foreign cells behind them (the layout of space->parts /
 * space->nr_parts in the reference); the proxies' slots [nlocal, ntotal) are filled on the device by
 * the xv halo exchange - the reference fills them with recv tasks (scheduler.c:1088-1112) - never by
 * the host. The matching download returns the local particles only. */
int swiftgpu_upload_parts_local(swiftgpu_t *h, const void *parts_aos, int64_t nlocal, int64_t ntotal);
int swiftgpu_download_parts_local(swiftgpu_t *h, void *parts_aos, int64_t nlocal);

int swiftgpu_set_step(swiftgpu_t *h, const swiftgpu_step *step);

/* Run all further work of this handle on the caller's CUDA stream
 * (`cuda_stream` is a cudaStream_t; NULL = the handle's own stream). The
 * reference has no equivalent: it is what lets an engine that already owns
 * streams (or a benchmark timing with events on its stream) order the GPU
 * phases against its own copies. */
int swiftgpu_set_stream(swiftgpu_t *h, void *cuda_stream);

/* runner_do_hydro_sort (runner.h:107) for every cell and every sid the
 * density/force work lists need. */
int swiftgpu_run_sort(swiftgpu_t *h);
/* hydro_init_part (as cell_drift_part does for active parts, cell_drift.c:361)
 * + runner_dosub_self1_density / runner_dosub_pair1_density
 * (runner_doiact_hydro.h:187,192) over all top-level selfs and pairs. */
int swiftgpu_run_density(swiftgpu_t *h);
/* runner_do_ghost (runner.h:96) for every top-level cell, including the
 * subset re-runs of the density loop for unconverged particles. Returns
 * non-zero if particles remain unconverged after max_smoothing_iterations
 * (runner_ghost.c:1583-1593). */
int swiftgpu_run_ghost(swiftgpu_t *h);
/* runner_dosub_{self,pair}1_gradient; no-op for Minimal and Gadget2. */
int swiftgpu_run_gradient(swiftgpu_t *h);
/* runner_do_extra_ghost (runner.h:98); no-op for Minimal and Gadget2. */
int swiftgpu_run_extra_ghost(swiftgpu_t *h);
/* runner_dosub_self2_force / runner_dosub_pair2_force (runner_doiact_hydro.h:189,194). */
int swiftgpu_run_force(swiftgpu_t *h);
/* runner_do_end_hydro_force (runner_others.c:815). */
int swiftgpu_run_end_force(swiftgpu_t *h);
/* Convenience: the phases of `mask` in dependency order
 * (engine_maketasks.c:2541-2583). */
int swiftgpu_run_step(swiftgpu_t *h, uint32_t phase_mask);

/* Scatter the device columns back into the host AoS array (the fields valid
 * after the last phase run: struct part's density/force union). */
int swiftgpu_download_parts(swiftgpu_t *h, void *parts_aos, int64_t nparts);
int swiftgpu_download_parts_device(swiftgpu_t *h, void *d_parts_aos,
                                   int64_t nparts);
/* Per-cell h_max / h_max_active after the ghost (runner_ghost.c:1621-1632);
 * after swiftgpu_run_drift also dx_max_part / dx_max_sort. */
int swiftgpu_download_cells(swiftgpu_t *h, swiftgpu_cell *cells, int32_t ncells);
/* hydro_compute_timestep (hydro/Minimal/hydro.h:440, Gadget2/hydro.h:444,
 * SPHENIX/hydro.h:475): the CFL time-step 2 kernel_gamma CFL a h /
 * (a_factor_sound_speed v_sig) of every ACTIVE particle, evaluated in the
 * epilogue of swiftgpu_run_end_force from the h and signal velocity the step
 * just produced (the host no longer needs v_sig back to get
 * dt). Inactive particles get -1. dt_cfl holds nparts floats in the host's
 * particle order. */
int swiftgpu_download_timestep(swiftgpu_t *h, float *dt_cfl, int64_t nparts);
/*
 * Drift on the device: cell_drift_part (src/cell_drift.c:159-400)
 * with force = 1 over every LOCAL top-level cell, i.e. engine_drift_all's mapper
 * (src/engine_drift.c:83) / the drift_part tasks of a step in which every cell
 * is drifted from the same ti_old_part. Per particle: drift_part
 * (src/drift.h:141-215: x += v_full dt_drift, v += a_hydro dt_kick_hydro,
 * hydro_predict_extra of the scheme - Minimal hydro.h:815, Gadget2 :798,
 * SPHENIX :1029 -, x_diff / x_diff_sort -= v_full dt_drift), the h_min/h_max
 * clamp, cell_set_part_h_depth, optionally part_init -> hydro_init_part of the
 * ACTIVE particles (runner_drift.c:45 passes init_particles = 1); per cell: the
 * reductions h_max, h_max_active, dx_max_part, dx_max_sort up the tree
 * (cell_drift.c:219-232,380-390).
 *
 * It works on the device copy of the caller's struct part[] (what
 * swiftgpu_upload_parts brought, with the results of the last step written
 * back into it first) and of struct xpart[] (only x_diff, x_diff_sort, v_full
 * are touched; gravity, cosmological factors, entropy / pressure floors,
 * forcing and particle removal at non-periodic borders are the host's). After
 * the call swiftgpu_download_parts / _xparts / _cells return the drifted state
 * and the next swiftgpu_run_step starts from it, without the particles
 * crossing the host boundary in between. With several ranks only the rank's
 * own cells are drifted (the reference drifts local cells only); the metadata
 * of the proxy cells (h_max, dx_max_*) stays as uploaded until the caller
 * refreshes it with swiftgpu_upload_cells, as it does after its own exchange
 * of cell data (the parity tests of the drift are single-rank).
 */
typedef struct swiftgpu_xpart_layout {
  int32_t size;        /* sizeof(struct xpart) */
  int32_t x_diff;      /* offsetof(struct xpart, x_diff), float[3] */
  int32_t x_diff_sort; /* float[3] */
  int32_t v_full;      /* float[3] */
  int32_t u_full;      /* float: u_full (Minimal, SPHENIX) | entropy_full (Gadget2); -1 if the kick is not used */
} swiftgpu_xpart_layout;

typedef struct swiftgpu_drift_args {
  /* the factors cell_drift_part derives (cell_drift.c:236-252): without
   * cosmology all three are (ti_current - ti_old_part) * time_base */
  double dt_drift, dt_kick_hydro, dt_therm;
  float minimal_internal_energy; /* hydro_props->minimal_internal_energy */
  int32_t init_particles;        /* 1: hydro_init_part of the active particles */
} swiftgpu_drift_args;

int swiftgpu_upload_xparts(swiftgpu_t *h, const swiftgpu_xpart_layout *layout,
                           const void *xparts_aos, int64_t nparts);
int swiftgpu_download_xparts(swiftgpu_t *h, void *xparts_aos, int64_t nparts);
int swiftgpu_run_drift(swiftgpu_t *h, const swiftgpu_drift_args *args);

/*
 * Kick on the device (the kick half): runner_do_kick1 (which = 1,
 * src/runner_time_integration.c:87) for the particles STARTING their step and
 * runner_do_kick2 (which = 2, :360) for the ACTIVE ones, hydro particles
 * without gravity: kick_part (src/kick.h:113: v_full += a_hydro dt_kick_hydro,
 * hydro_kick_extra of the scheme - u_full | entropy_full += du/dt dt_therm, at
 * most halved, floored at minimal_internal_energy) with the half time-step of
 * the particle's own time_bin, (ti_step / 2) * time_base; kick2 then calls
 * hydro_reset_predicted_values (p->v, p->u | entropy, pressure | P_over_rho2,
 * sound speed, v_sig from the full-step values). Works on the device copies of
 * struct part[] / struct xpart[] like the drift; together with it and
 * swiftgpu_run_step a fixed-time-step leapfrog runs without the particles
 * crossing the host boundary. Not covered: the assignment of new time bins
 * (runner_do_timestep), cosmological kick factors.
 */
int swiftgpu_run_kick(swiftgpu_t *h, int which, float minimal_internal_energy);

/*
 * The time-step limiter loop (the loop half):
 * runner_dosub_self1_limiter / runner_dosub_pair1_limiter (runner_main.c:233,292;
 * runner_doiact_functions_limiter.h) with runner_iact_nonsym_limiter
 * (timestep_limiter_iact.h:106-117) - every particle within the kernel of a
 * particle STARTING its step whose time bin lies more than
 * time_bin_neighbour_max_delta_bin (2) above it gets limiter_data.wakeup =
 * max(wakeup, -time_bin of the starter). Runs after the ghost (any time after
 * the step) on the decomposition of the density loop; wakeup_offset =
 * offsetof(struct part, limiter_data.wakeup). swiftgpu_download_parts returns
 * the field with the particles. The follow-up runner_do_limiter (re-binning
 * the woken particles) is the host's. Single rank only.
 */
int swiftgpu_run_limiter(swiftgpu_t *h, int32_t wakeup_offset);

/* Per-particle directed interaction counts of the last density / gradient /
 * force loops (the reference's N_density/N_gradient/N_force debugging counters,
 * hydro/SPHENIX/hydro_iact.h:121-126, minus the self term). Any pointer may be
 * NULL. */
int swiftgpu_download_counts(swiftgpu_t *h, int32_t *n_density,
                             int32_t *n_gradient, int32_t *n_force,
                             int64_t nparts);
int swiftgpu_get_stats(swiftgpu_t *h, swiftgpu_stats *out);
/* The sorted index array of (cell, sid) - c->hydro.sort of the reference
 * (sort_part.h:32, runner_sort.c:203; indices relative to the cell's first
 * particle, ascending key) - and its first / last key. The step itself only
 * consumes the extrema; the full arrays are produced on demand by this call. */
int swiftgpu_download_sort(swiftgpu_t *h, int32_t cell, int32_t sid, int32_t *idx_out,
                           float *key_min, float *key_max);

/* Host-only (no CUDA call): flattens the reference's recursive task functions
 * (DOSUB_SELF1/PAIR1 for loop 0 = density/gradient, DOSUB_SELF2/PAIR2 for loop
 * 2 = force, DOSUB_*_SUBSET for loop 3 = ghost re-runs) over the given tree and
 * reports out[0] = directed leaf-level items, out[1] = target-cell groups,
 * out[2] = sum over items of count(target cell) * count(source cell),
 * out[3] = (cell, sid) sorted arrays requested, out[4] = self items,
 * out[5] = items restricted by depth_h (limit_min_h / limit_max_h). Used by
 * the host-logic tests and to size device buffers. */
int swiftgpu_worklist_stats(const swiftgpu_config *cfg, const swiftgpu_step *step,
                            const swiftgpu_cell *cells, int32_t ncells,
                            const int32_t *top, int32_t ntop, int loop,
                            int64_t out[6]);

/* Host-only: order-sensitive 64-bit digest of the flattened list of `loop`
 * (every item and group in list order). */
int swiftgpu_worklist_digest(const swiftgpu_config *cfg, const swiftgpu_step *step,
                             const swiftgpu_cell *cells, int32_t ncells,
                             const int32_t *top, int32_t ntop, int loop,
                             uint64_t *digest);

/*
 * Multi-GPU: one rank (process) per GPU, top-level cells assigned to ranks by
 * the reference's partition (cell.nodeID, src/partition.c:104-121). A rank
 * uploads its own cells plus a read-only copy of every foreign top-level cell
 * adjacent to one of them (the reference's proxies, src/engine_proxy.c). The
 * halo exchange replaces the send/recv tasks of the reference
 * (scheduler.c:977-988,1088-1112; dependencies engine_maketasks.c:208-330):
 *   phase 0 "xv"        before density:  x, v, m, h, u, time_bin, depth_h and
 *                                        the force-union members of inactive parts
 *   phase 1 "rho"       after the ghost: h, rho, depth_h, P|P/rho^2, c_s, f,
 *                                        balsara, (alpha); then h_max/h_max_active
 *                                        of the foreign cells are recomputed like
 *                                        runner_do_recv_part (runner_recv.c:95-137)
 *   phase 2 "gradient"  after the extra ghost (SPHENIX): alpha_visc, alpha_diff
 * Transport is NCCL point-to-point (grouped ncclSend/ncclRecv of device-packed
 * SoA slabs) over NVLink; libnccl.so.2 is resolved at run time.
 *
 * swiftgpu_nccl_unique_id: rank 0 obtains the 128-byte NCCL id and distributes
 * it by any means (MPI_Bcast in SWIFT, torch.distributed in the benchmark);
 * swiftgpu_halo_setup creates the communicator (cfg.rank / cfg.nranks) and the
 * send/receive lists from the uploaded cells. swiftgpu_run_step() performs the
 * three exchanges itself once the halo is set up.
 */
int swiftgpu_nccl_unique_id(void *id128);
int swiftgpu_halo_setup(swiftgpu_t *h, const void *id128);
int swiftgpu_halo_exchange(swiftgpu_t *h, int phase);

/* Host-only: the halo plan of this rank for `peer`, from the cells alone.
 * Fills send_cells / recv_cells (capacity ntop each) with indices into
 * `cells` of the top-level cells sent to / received from that peer, both
 * ordered by cell location so that the two sides agree without negotiation,
 * and the particle totals. */
int swiftgpu_halo_plan(const swiftgpu_config *cfg, const swiftgpu_cell *cells, int32_t ncells,
                       const int32_t *top, int32_t ntop, int32_t peer, int32_t *send_cells,
                       int32_t *nsend_cells, int32_t *recv_cells, int32_t *nrecv_cells,
                       int64_t *nsend_parts, int64_t *nrecv_parts);

#ifdef __cplusplus
}
#endif
#endif /* SWIFTGPU_H */
