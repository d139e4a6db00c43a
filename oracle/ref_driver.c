/*
 * oracle/ref_driver.c - TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Driver that is compiled together with the UNMODIFIED reference objects
 * (oracle/Makefile, target `ref`) into oracle/_ref/libswiftref_<scheme>.so.
 * It plays the role of tests/test125cells.c in the reference: it hand-builds a
 * struct space / engine / runner on the stack (test125cells.c:583-625),
 * reconstructs the reference's `struct cell` tree from the flattened
 * swiftgpu_cell array of include/swiftgpu.h, and then calls the reference's own
 * task functions in the order of the task graph
 * (engine_maketasks.c:2541-2583):
 *
 *   runner_do_hydro_sort -> runner_dosub_self1/pair1_density -> runner_do_ghost
 *   -> [runner_dosub_self1/pair1_gradient -> runner_do_extra_ghost]
 *   -> runner_dosub_self2/pair2_force -> runner_do_end_hydro_force
 *
 * Nothing of the algorithm is restated here: every number comes out of the
 * reference's code. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 */
#include <config.h>

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "swift.h"

#include "../include/swiftgpu.h"

#ifndef SWIFTREF_SCHEME_NAME
#define SWIFTREF_SCHEME_NAME "unknown"
#endif

#if defined(MINIMAL_SPH)
#define SWIFTREF_SCHEME SWIFTGPU_SCHEME_MINIMAL
#elif defined(GADGET2_SPH)
#define SWIFTREF_SCHEME SWIFTGPU_SCHEME_GADGET2
#elif defined(SPHENIX_SPH)
#define SWIFTREF_SCHEME SWIFTGPU_SCHEME_SPHENIX
#else
#error "unsupported scheme"
#endif

/* engine_rank (the global read by the reference's error()/message() macros)
 * is defined by the reference's engine.c. */

struct ref_task {
  struct task t;        /* only type, subtype, ci, cj are read (runner_ghost.c:1554-1566) */
  volatile int state;   /* 0 = todo, 1 = taken/done (per phase) */
};

typedef struct swiftref {
  swiftgpu_config cfg;
  swiftgpu_step step;
  struct space space;
  struct engine engine;
  struct hydro_props hp;
  struct cosmology cosmo;
  struct phys_const phys_const;
  struct sink_props sink_props;
  struct lightcone_array_props lightcone_props;
  struct pressure_floor_props pressure_floor;
  struct cell *cells;
  int ncells;
  int *top;
  int ntop;
  struct part *parts;
  struct xpart *xparts;
  long long nparts;
  struct ref_task *tasks;
  int ntasks;
  struct link *links;
  pthread_mutex_t *top_locks; /* one per cell index (only top-level used) */
  double seconds[8];
} swiftref_t;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

const char *swiftref_scheme_name(void) { return SWIFTREF_SCHEME_NAME; }
int swiftref_scheme(void) { return SWIFTREF_SCHEME; }
int swiftref_part_size(void) { return (int)sizeof(struct part); }
int swiftref_has_counts(void) {
#if defined(SWIFT_HYDRO_DENSITY_CHECKS) && defined(SPHENIX_SPH)
  return 1;
#elif defined(DEBUG_INTERACTIONS_SPH)
  return 2;
#else
  return 0;
#endif
}

/* offsetof() table of this build's struct part in swiftgpu_part_layout order. */
void swiftref_layout(swiftgpu_part_layout *L) {
  memset(L, 0xff, sizeof(*L)); /* all -1 */
  L->size = sizeof(struct part);
  L->id = offsetof(struct part, id);
  L->x = offsetof(struct part, x);
  L->v = offsetof(struct part, v);
  L->a_hydro = offsetof(struct part, a_hydro);
  L->mass = offsetof(struct part, mass);
  L->h = offsetof(struct part, h);
  L->rho = offsetof(struct part, rho);
  L->wcount = offsetof(struct part, density.wcount);
  L->wcount_dh = offsetof(struct part, density.wcount_dh);
  L->rho_dh = offsetof(struct part, density.rho_dh);
  L->rot_v = offsetof(struct part, density.rot_v);
  L->f = offsetof(struct part, force.f);
  L->soundspeed = offsetof(struct part, force.soundspeed);
  L->h_dt = offsetof(struct part, force.h_dt);
  L->balsara = offsetof(struct part, force.balsara);
  L->time_bin = offsetof(struct part, time_bin);
  L->depth_h = offsetof(struct part, depth_h);
  L->min_ngb_time_bin = offsetof(struct part, limiter_data.min_ngb_time_bin);
#if defined(MINIMAL_SPH)
  L->u = offsetof(struct part, u);
  L->u_dt = offsetof(struct part, u_dt);
  L->div_v = offsetof(struct part, density.div_v);
  L->pressure = offsetof(struct part, force.pressure);
  L->v_sig = offsetof(struct part, force.v_sig);
#elif defined(GADGET2_SPH)
  L->entropy = offsetof(struct part, entropy);
  L->entropy_dt = offsetof(struct part, entropy_dt);
  L->div_v = offsetof(struct part, density.div_v);
  L->P_over_rho2 = offsetof(struct part, force.P_over_rho2);
  L->v_sig = offsetof(struct part, force.v_sig);
#elif defined(SPHENIX_SPH)
  L->u = offsetof(struct part, u);
  L->u_dt = offsetof(struct part, u_dt);
  L->div_v = offsetof(struct part, viscosity.div_v);
  L->pressure = offsetof(struct part, force.pressure);
  L->v_sig = offsetof(struct part, viscosity.v_sig);
  L->div_v_dt = offsetof(struct part, viscosity.div_v_dt);
  L->div_v_previous_step = offsetof(struct part, viscosity.div_v_previous_step);
  L->visc_alpha = offsetof(struct part, viscosity.alpha);
  L->laplace_u = offsetof(struct part, diffusion.laplace_u);
  L->diff_alpha = offsetof(struct part, diffusion.alpha);
  L->alpha_visc_max_ngb = offsetof(struct part, force.alpha_visc_max_ngb);
#endif
}

/* Constants of the build, for pinning the port and the CUDA constants. */
void swiftref_constants(double *out) {
  out[0] = kernel_gamma;
  out[1] = kernel_gamma2;
  out[2] = kernel_root;
  out[3] = kernel_norm;
  out[4] = kernel_constant;
  out[5] = kernel_gamma_inv_dim;
  out[6] = kernel_gamma_inv_dim_plus_one;
  out[7] = hydro_gamma;
  out[8] = space_splitsize;
  out[9] = space_recurse_size_self_hydro;
  out[10] = space_recurse_size_pair_hydro;
  out[11] = space_maxreldx;
  out[12] = const_viscosity_beta;
  out[13] = sizeof(struct part);
  out[14] = sizeof(struct xpart);
  out[15] = sizeof(struct cell);
}

/* kernel_deval (kernel_hydro.h:257) exposed for known-answer tests. */
void swiftref_kernel_deval(float u, float *w, float *dw) {
  kernel_deval(u, w, dw);
}

static void link_task(swiftref_t *s, int *nlinks, struct cell *c,
                      struct task *t) {
  struct link *l = &s->links[(*nlinks)++];
  l->t = t;
  l->next = c->hydro.density;
  c->hydro.density = l;
}

static int is_neighbour(const swiftref_t *s, const struct cell *a,
                        const struct cell *b) {
  /* Top-level cells are neighbours if they touch (periodic-aware), the rule
   * of engine_make_hydroloop_tasks_mapper (engine_maketasks.c:3501). */
  for (int k = 0; k < 3; k++) {
    double dx = fabs(b->loc[k] - a->loc[k]);
    if (s->space.periodic && dx > 0.5 * s->space.dim[k])
      dx = s->space.dim[k] - dx;
    if (dx > 1.0001 * a->width[k]) return 0;
  }
  return 1;
}

swiftref_t *swiftref_create(const swiftgpu_config *cfg,
                            const swiftgpu_step *step,
                            const swiftgpu_cell *cells, int ncells,
                            const int *top, int ntop, const void *parts_aos,
                            long long nparts) {
  if (cfg->scheme != SWIFTREF_SCHEME) return NULL;
  if (cfg->layout.size != (int)sizeof(struct part)) return NULL;

  swiftref_t *s = (swiftref_t *)calloc(1, sizeof(swiftref_t));
  s->cfg = *cfg;
  s->step = *step;
  s->ncells = ncells;
  s->ntop = ntop;
  s->nparts = nparts;

  /* Infrastructure, as tests/test125cells.c:583-625. */
  s->space.periodic = cfg->periodic;
  for (int k = 0; k < 3; k++) s->space.dim[k] = cfg->dim[k];
  hydro_space_init(&s->space.hs, &s->space);
  s->phys_const.const_newton_G = 1.f;
  s->phys_const.const_vacuum_permeability = 1.0;

  hydro_props_init_no_hydro(&s->hp);
  s->hp.eta_neighbours = cfg->eta_neighbours;
  s->hp.h_tolerance = cfg->h_tolerance;
  s->hp.h_max = cfg->h_max;
  s->hp.h_min = cfg->h_min;
  s->hp.h_min_ratio = 0.f;
  s->hp.max_smoothing_iterations = cfg->max_smoothing_iterations;
  s->hp.use_mass_weighted_num_ngb = cfg->use_mass_weighted_num_ngb;
  s->hp.CFL_condition = cfg->CFL_condition;
  s->hp.target_neighbours = pow_dimension(s->hp.eta_neighbours) * kernel_norm;
#if defined(SPHENIX_SPH)
  s->hp.viscosity.alpha = cfg->viscosity_alpha;
  s->hp.viscosity.alpha_max = cfg->viscosity_alpha_max;
  s->hp.viscosity.alpha_min = cfg->viscosity_alpha_min;
  s->hp.viscosity.length = cfg->viscosity_length;
  s->hp.diffusion.alpha = cfg->diffusion_alpha;
  s->hp.diffusion.beta = cfg->diffusion_beta;
  s->hp.diffusion.alpha_max = cfg->diffusion_alpha_max;
  s->hp.diffusion.alpha_min = cfg->diffusion_alpha_min;
#endif

  bzero(&s->engine, sizeof(struct engine));
  s->engine.hydro_properties = &s->hp;
  s->engine.physical_constants = &s->phys_const;
  s->engine.s = &s->space;
  s->engine.time = 0.1f;
  s->engine.ti_current = step->ti_current;
  s->engine.max_active_bin = step->max_active_bin;
  s->engine.time_base = step->time_base;
  s->engine.nodeID = cfg->rank;
  s->engine.policy = engine_policy_hydro;
  cosmology_init_no_cosmo(&s->cosmo);
  s->engine.cosmology = &s->cosmo;
  bzero(&s->sink_props, sizeof(struct sink_props));
  s->engine.sink_properties = &s->sink_props;
  s->lightcone_props.nr_lightcones = 0;
  s->engine.lightcone_array_properties = &s->lightcone_props;
  s->engine.pressure_floor_props = &s->pressure_floor;
  s->space.e = &s->engine;
  engine_rank = cfg->rank;

  /* Particles. */
  if (posix_memalign((void **)&s->parts, part_align,
                     (nparts + 1) * sizeof(struct part)) != 0)
    return NULL;
  if (posix_memalign((void **)&s->xparts, xpart_align,
                     (nparts + 1) * sizeof(struct xpart)) != 0)
    return NULL;
  memcpy(s->parts, parts_aos, nparts * sizeof(struct part));
  bzero(s->xparts, (nparts + 1) * sizeof(struct xpart));
  for (long long k = 0; k < nparts; k++) s->parts[k].gpart = NULL;
  s->space.parts = s->parts;
  s->space.xparts = s->xparts;
  s->space.nr_parts = nparts;

  /* Cells. */
  if (posix_memalign((void **)&s->cells, cell_align,
                     ncells * sizeof(struct cell)) != 0)
    return NULL;
  bzero(s->cells, ncells * sizeof(struct cell));
  s->top = (int *)malloc(sizeof(int) * ntop);
  memcpy(s->top, top, sizeof(int) * ntop);
  for (int i = 0; i < ncells; i++) {
    const swiftgpu_cell *g = &cells[i];
    struct cell *c = &s->cells[i];
    for (int k = 0; k < 3; k++) {
      c->loc[k] = g->loc[k];
      c->width[k] = g->width[k];
    }
    c->dmin = g->dmin;
    c->h_min_allowed = g->h_min_allowed;
    c->h_max_allowed = g->h_max_allowed;
    c->depth = g->depth;
    c->split = g->split;
    c->parent = g->parent >= 0 ? &s->cells[g->parent] : NULL;
    for (int k = 0; k < 8; k++)
      c->progeny[k] = g->progeny[k] >= 0 ? &s->cells[g->progeny[k]] : NULL;
    c->nodeID = g->nodeID;
    c->top = &s->cells[g->top];
    c->super = c->top;
    c->hydro.super = c->top;
    c->hydro.parts = s->parts + g->first_part;
    c->hydro.xparts = s->xparts + g->first_part;
    c->hydro.count = g->count;
    c->hydro.count_total = g->count;
    c->hydro.h_max = g->h_max;
    c->hydro.h_max_active = g->h_max_active;
    c->hydro.h_max_old = g->h_max_old;
    c->hydro.dx_max_part = g->dx_max_part;
    c->hydro.dx_max_part_old = g->dx_max_part_old;
    c->hydro.dx_max_sort = g->dx_max_sort;
    c->hydro.dx_max_sort_old = g->dx_max_sort_old;
    c->hydro.ti_old_part = step->ti_current; /* drifted */
    c->hydro.ti_end_min = g->ti_end_min;
    c->hydro.sorted = 0;
    c->hydro.sort = NULL;
    lock_init(&c->hydro.lock);
    lock_init(&c->hydro.extra_sort_lock);
  }

  /* Tasks: one self per top-level cell, one pair per unordered couple of
   * touching top-level cells, all linked into c->hydro.density so that the
   * ghost's redo loop finds them (runner_ghost.c:1548-1572). */
  int max_tasks = ntop * 14 + 16;
  s->tasks = (struct ref_task *)calloc(max_tasks, sizeof(struct ref_task));
  s->links = (struct link *)calloc(2 * max_tasks, sizeof(struct link));
  int nt = 0, nl = 0;
  for (int a = 0; a < ntop; a++) {
    struct cell *ca = &s->cells[top[a]];
    struct ref_task *t = &s->tasks[nt++];
    t->t.type = task_type_self;
    t->t.subtype = task_subtype_density;
    t->t.ci = ca;
    t->t.cj = NULL;
    link_task(s, &nl, ca, &t->t);
  }
  /* Touching pairs. Use the top-level grid if it is regular: find neighbours
   * through a hash on integer coordinates to stay O(ntop). */
  {
    const struct cell *c0 = &s->cells[top[0]];
    int cdim[3];
    for (int k = 0; k < 3; k++)
      cdim[k] = (int)floor(s->space.dim[k] / c0->width[k] + 0.5);
    long long ngrid = (long long)cdim[0] * cdim[1] * cdim[2];
    int *grid = (int *)malloc(sizeof(int) * ngrid);
    for (long long i = 0; i < ngrid; i++) grid[i] = -1;
    for (int a = 0; a < ntop; a++) {
      const struct cell *ca = &s->cells[top[a]];
      int ix = (int)floor(ca->loc[0] / ca->width[0] + 0.5);
      int iy = (int)floor(ca->loc[1] / ca->width[1] + 0.5);
      int iz = (int)floor(ca->loc[2] / ca->width[2] + 0.5);
      grid[((long long)ix * cdim[1] + iy) * cdim[2] + iz] = a;
    }
    for (int a = 0; a < ntop; a++) {
      struct cell *ca = &s->cells[top[a]];
      int ix = (int)floor(ca->loc[0] / ca->width[0] + 0.5);
      int iy = (int)floor(ca->loc[1] / ca->width[1] + 0.5);
      int iz = (int)floor(ca->loc[2] / ca->width[2] + 0.5);
      for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dz = -1; dz <= 1; dz++) {
            if (dx == 0 && dy == 0 && dz == 0) continue;
            int jx = ix + dx, jy = iy + dy, jz = iz + dz;
            if (s->space.periodic) {
              jx = (jx + cdim[0]) % cdim[0];
              jy = (jy + cdim[1]) % cdim[1];
              jz = (jz + cdim[2]) % cdim[2];
            } else if (jx < 0 || jy < 0 || jz < 0 || jx >= cdim[0] ||
                       jy >= cdim[1] || jz >= cdim[2])
              continue;
            int b = grid[((long long)jx * cdim[1] + jy) * cdim[2] + jz];
            if (b < 0 || b <= a) continue; /* each unordered pair once */
            struct cell *cb = &s->cells[top[b]];
            if (!is_neighbour(s, ca, cb)) continue;
            /* With fewer than 3 cells along a periodic axis a couple would
             * appear more than once: keep the first occurrence only. */
            int dup = 0;
            for (struct link *l = ca->hydro.density; l != NULL; l = l->next)
              if (l->t->type == task_type_pair &&
                  ((l->t->ci == ca && l->t->cj == cb) ||
                   (l->t->ci == cb && l->t->cj == ca)))
                dup = 1;
            if (dup) continue;
            /* A pair task is only created if one side is local
             * (engine_maketasks.c:3562-3569). */
            if (ca->nodeID != cfg->rank && cb->nodeID != cfg->rank) continue;
            struct ref_task *t = &s->tasks[nt++];
            t->t.type = task_type_pair;
            t->t.subtype = task_subtype_density;
            t->t.ci = ca;
            t->t.cj = cb;
            link_task(s, &nl, ca, &t->t);
            link_task(s, &nl, cb, &t->t);
          }
    }
    free(grid);
  }
  s->ntasks = nt;
  s->top_locks = (pthread_mutex_t *)malloc(sizeof(pthread_mutex_t) * ncells);
  for (int i = 0; i < ncells; i++) pthread_mutex_init(&s->top_locks[i], NULL);
  return s;
}

void swiftref_destroy(swiftref_t *s) {
  if (!s) return;
  for (int i = 0; i < s->ncells; i++) {
    if (s->cells[i].hydro.sort) cell_free_hydro_sorts(&s->cells[i]);
    pthread_mutex_destroy(&s->top_locks[i]);
  }
  free(s->top_locks);
  free(s->cells);
  free(s->parts);
  free(s->xparts);
  free(s->tasks);
  free(s->links);
  free(s->top);
  free(s);
}

/* ---- threaded execution of one phase ---- */

enum ref_kind {
  K_INIT,
  K_SORT,
  K_DENSITY,
  K_GHOST,
  K_GRADIENT,
  K_EXTRA_GHOST,
  K_FORCE,
  K_END_FORCE,
  K_LIMITER
};

struct worker {
  swiftref_t *s;
  enum ref_kind kind;
  volatile int *next; /* shared counter */
  int n;
  struct runner runner;
  pthread_t th;
};

static void init_cell_parts(swiftref_t *s, struct cell *c) {
  /* cell_drift_part zeroes the density sums of the active particles
   * (cell_drift.c:361); test125cells.c:668-674 does it for all. */
  for (int k = 0; k < c->hydro.count; k++) {
    struct part *p = &c->hydro.parts[k];
    if (part_is_active(p, &s->engine)) {
      hydro_init_part(p, &s->space.hs);
      adaptive_softening_init_part(p);
      mhd_init_part(p);
    }
  }
}

static void run_per_cell(struct worker *w, struct cell *c) {
  swiftref_t *s = w->s;
  struct runner *r = &w->runner;
  switch (w->kind) {
    case K_INIT:
      init_cell_parts(s, c);
      break;
    case K_SORT:
      if (c->hydro.count > 0)
        runner_do_hydro_sort(r, c, 0x1FFF, /*cleanup=*/0, /*lock=*/0,
                             /*rt=*/0, /*clock=*/0);
      break;
    case K_GHOST:
      if (c->nodeID == s->engine.nodeID)
        runner_do_ghost(r, c, /*offset=*/0, /*ntasks=*/1, /*timer=*/0);
      break;
    case K_EXTRA_GHOST:
#ifdef EXTRA_HYDRO_LOOP
      if (c->nodeID == s->engine.nodeID) runner_do_extra_ghost(r, c, 0);
#endif
      break;
    case K_END_FORCE:
      if (c->nodeID == s->engine.nodeID) runner_do_end_hydro_force(r, c, 0);
      break;
    default:
      break;
  }
}

/* defined by src/runner_doiact_limiter.c (runner_doiact_functions_limiter.h:852,957) */
void runner_dosub_self1_limiter(struct runner *r, struct cell *c, int recurse_below_h_max,
                                const int gettimer);
void runner_dosub_pair1_limiter(struct runner *r, struct cell *ci, struct cell *cj,
                                int recurse_below_h_max, const int gettimer);

static void run_task(struct worker *w, struct ref_task *t) {
  struct runner *r = &w->runner;
  struct cell *ci = t->t.ci, *cj = t->t.cj;
  switch (w->kind) {
    case K_DENSITY:
      if (cj == NULL)
        runner_dosub_self1_density(r, ci, /*below_h_max=*/0, 0);
      else
        runner_dosub_pair1_density(r, ci, cj, /*below_h_max=*/0, 0);
      break;
    case K_GRADIENT:
#ifdef EXTRA_HYDRO_LOOP
      if (cj == NULL)
        runner_dosub_self1_gradient(r, ci, 0, 0);
      else
        runner_dosub_pair1_gradient(r, ci, cj, 0, 0);
#endif
      break;
    case K_FORCE:
      if (cj == NULL)
        runner_dosub_self2_force(r, ci, 0, 0);
      else
        runner_dosub_pair2_force(r, ci, cj, 0, 0);
      break;
    case K_LIMITER: /* runner_main.c:233-234,292-293 */
      if (cj == NULL)
        runner_dosub_self1_limiter(r, ci, /*below_h_max=*/0, 0);
      else
        runner_dosub_pair1_limiter(r, ci, cj, /*below_h_max=*/0, 0);
      break;
    default:
      break;
  }
}

static void *worker_main(void *arg) {
  struct worker *w = (struct worker *)arg;
  swiftref_t *s = w->s;
  const int per_cell = (w->kind == K_INIT || w->kind == K_SORT ||
                        w->kind == K_GHOST || w->kind == K_EXTRA_GHOST ||
                        w->kind == K_END_FORCE);
  if (per_cell) {
    /* The ghost of one top-level cell re-runs density loops that only WRITE
     * particles of that cell (non-symmetric subset loops), so top-level cells
     * are independent units, like the per-super-cell ghost tasks. */
    for (;;) {
      int i = __atomic_fetch_add((int *)w->next, 1, __ATOMIC_RELAXED);
      if (i >= s->ntop) break;
      run_per_cell(w, &s->cells[s->top[i]]);
    }
  } else {
    /* Self/pair tasks write both cells: take the two top-level locks like
     * task_lock (task.c:759-830); skip to another task when busy. */
    int remaining = 1;
    while (remaining) {
      remaining = 0;
      for (int i = 0; i < s->ntasks; i++) {
        struct ref_task *t = &s->tasks[i];
        if (t->state) continue;
        int a = (int)(t->t.ci - s->cells);
        int b = t->t.cj ? (int)(t->t.cj - s->cells) : -1;
        if (pthread_mutex_trylock(&s->top_locks[a]) != 0) {
          remaining = 1;
          continue;
        }
        if (b >= 0 && pthread_mutex_trylock(&s->top_locks[b]) != 0) {
          pthread_mutex_unlock(&s->top_locks[a]);
          remaining = 1;
          continue;
        }
        if (!t->state) {
          t->state = 1;
          run_task(w, t);
        }
        if (b >= 0) pthread_mutex_unlock(&s->top_locks[b]);
        pthread_mutex_unlock(&s->top_locks[a]);
      }
    }
  }
  return NULL;
}

static double run_kind(swiftref_t *s, enum ref_kind kind, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  volatile int next = 0;
  for (int i = 0; i < s->ntasks; i++) s->tasks[i].state = 0;
  struct worker *ws = (struct worker *)calloc(nthreads, sizeof(struct worker));
  const double t0 = now_s();
  for (int t = 0; t < nthreads; t++) {
    ws[t].s = s;
    ws[t].kind = kind;
    ws[t].next = &next;
    ws[t].runner.e = &s->engine;
    ws[t].runner.id = t;
#ifdef WITH_VECTORIZATION
    /* engine_config.c:1038-1043: the particle caches of the hand-vectorised loops */
    ws[t].runner.ci_cache.count = 0;
    ws[t].runner.cj_cache.count = 0;
    cache_init(&ws[t].runner.ci_cache, 512);
    cache_init(&ws[t].runner.cj_cache, 512);
#endif
    if (nthreads > 1) pthread_create(&ws[t].th, NULL, worker_main, &ws[t]);
  }
  if (nthreads == 1)
    worker_main(&ws[0]);
  else
    for (int t = 0; t < nthreads; t++) pthread_join(ws[t].th, NULL);
  const double dt = now_s() - t0;
#ifdef WITH_VECTORIZATION
  for (int t = 0; t < nthreads; t++) {
    cache_clean(&ws[t].runner.ci_cache);
    cache_clean(&ws[t].runner.cj_cache);
  }
#endif
  free(ws);
  return dt;
}

/* Runs the phases of `mask` in task-graph order. seconds[7] (may be NULL)
 * receives wall-clock per phase: sort, density(+init), ghost, gradient,
 * extra_ghost, force, end_force. */
int swiftref_run(swiftref_t *s, unsigned mask, int nthreads, double *seconds) {
  double sec[7] = {0, 0, 0, 0, 0, 0, 0};
  if (mask & SWIFTGPU_PHASE_SORT) sec[0] = run_kind(s, K_SORT, nthreads);
  if (mask & SWIFTGPU_PHASE_DENSITY) {
    sec[1] = run_kind(s, K_INIT, nthreads);
    sec[1] += run_kind(s, K_DENSITY, nthreads);
  }
  if (mask & SWIFTGPU_PHASE_GHOST) sec[2] = run_kind(s, K_GHOST, nthreads);
#ifdef EXTRA_HYDRO_LOOP
  if (mask & SWIFTGPU_PHASE_GRADIENT) sec[3] = run_kind(s, K_GRADIENT, nthreads);
  if (mask & SWIFTGPU_PHASE_EXTRA_GHOST)
    sec[4] = run_kind(s, K_EXTRA_GHOST, nthreads);
#endif
  if (mask & SWIFTGPU_PHASE_FORCE) sec[5] = run_kind(s, K_FORCE, nthreads);
  if (mask & SWIFTGPU_PHASE_END_FORCE)
    sec[6] = run_kind(s, K_END_FORCE, nthreads);
  if (seconds) memcpy(seconds, sec, sizeof(sec));
  return 0;
}

/* Replace the particle state (e.g. to re-run a step from the same ICs). */
int swiftref_set_parts(swiftref_t *s, const void *parts_aos) {
  memcpy(s->parts, parts_aos, s->nparts * sizeof(struct part));
  for (long long k = 0; k < s->nparts; k++) s->parts[k].gpart = NULL;
  return 0;
}

int swiftref_get_parts(swiftref_t *s, void *parts_aos) {
  memcpy(parts_aos, s->parts, s->nparts * sizeof(struct part));
  return 0;
}

int swiftref_get_cells(swiftref_t *s, swiftgpu_cell *cells) {
  for (int i = 0; i < s->ncells; i++) {
    cells[i].h_max = s->cells[i].hydro.h_max;
    cells[i].h_max_active = s->cells[i].hydro.h_max_active;
    cells[i].dx_max_part = s->cells[i].hydro.dx_max_part;
    cells[i].dx_max_sort = s->cells[i].hydro.dx_max_sort;
    cells[i].dx_max_sort_old = s->cells[i].hydro.dx_max_sort_old;
  }
  return 0;
}

/* ---- the reference's own tree builder (checks swift_b200/csrc/host_tree.cpp) ---- */
static int emit_subtree(const struct cell *c, const struct part *parts0, int parent, int top,
                        swiftgpu_cell *out, int max_cells, int *n) {
  if (*n >= max_cells) return -1;
  const int me = (*n)++;
  swiftgpu_cell *g = &out[me];
  memset(g, 0, sizeof(*g));
  for (int k = 0; k < 3; k++) {
    g->loc[k] = c->loc[k];
    g->width[k] = c->width[k];
  }
  g->dmin = c->dmin;
  g->h_min_allowed = c->h_min_allowed;
  g->h_max_allowed = c->h_max_allowed;
  g->h_max = c->hydro.h_max;
  g->h_max_active = c->hydro.h_max_active;
  g->depth = c->depth;
  g->split = c->split;
  g->parent = parent;
  g->top = top;
  g->nodeID = c->nodeID;
  g->count = c->hydro.count;
  g->first_part = (long long)(c->hydro.parts - parts0);
  g->ti_end_min = c->hydro.ti_end_min;
  for (int k = 0; k < 8; k++) {
    g->progeny[k] = -1;
    if (c->split && c->progeny[k] != NULL) {
      g->progeny[k] = *n;
      if (emit_subtree(c->progeny[k], parts0, me, top, out, max_cells, n) < 0) return -1;
    }
  }
  return me;
}

/* space_split_recursive (src/space_split.c:53; cell_split src/cell_split.c) on ONE top-level cell
 * [loc, loc + width) holding parts_aos[0..n), set up as space_regrid does (space_regrid.c:300-330).
 * The particles are re-ordered in place by the reference's cell_split; the subtree comes back in
 * depth-first pre-order (progeny 0..7), indices relative to this subtree, first_part relative to
 * parts_aos[0]. Returns the number of cells, or -1 if max_cells is too small. */
int swiftref_space_split(const swiftgpu_config *cfg, const swiftgpu_step *step, void *parts_aos,
                         long long n, const double loc[3], const double width[3],
                         swiftgpu_cell *out, int max_cells) {
  if (cfg->scheme != SWIFTREF_SCHEME || cfg->layout.size != (int)sizeof(struct part)) return -2;
  struct space *s = (struct space *)calloc(1, sizeof(struct space));
  struct engine *e = (struct engine *)calloc(1, sizeof(struct engine));
  struct part *parts = NULL;
  struct xpart *xparts = NULL;
  if (posix_memalign((void **)&parts, part_align, (n + 1) * sizeof(struct part)) != 0) return -2;
  if (posix_memalign((void **)&xparts, xpart_align, (n + 1) * sizeof(struct xpart)) != 0) return -2;
  memcpy(parts, parts_aos, n * sizeof(struct part));
  bzero(xparts, (n + 1) * sizeof(struct xpart));
  for (long long k = 0; k < n; k++) parts[k].gpart = NULL;
  e->ti_current = step->ti_current;
  e->max_active_bin = step->max_active_bin;
  e->time_base = step->time_base;
  e->policy = engine_policy_hydro;
  e->s = s;
  s->e = e;
  s->periodic = cfg->periodic;
  for (int k = 0; k < 3; k++) s->dim[k] = cfg->dim[k];
  s->parts = parts;
  s->xparts = xparts;
  s->nr_parts = n;
  s->with_self_gravity = 0;
  s->cells_sub = (struct cell **)calloc(4, sizeof(struct cell *));
  s->multipoles_sub = (struct gravity_tensors **)calloc(4, sizeof(struct gravity_tensors *));
  lock_init(&s->lock);
  struct cell *c = NULL;
  if (posix_memalign((void **)&c, cell_align, sizeof(struct cell)) != 0) return -2;
  bzero(c, sizeof(struct cell));
  for (int k = 0; k < 3; k++) {
    c->loc[k] = loc[k];
    c->width[k] = width[k];
  }
  c->dmin = (float)fmin(width[0], fmin(width[1], width[2]));
  c->h_min_allowed = c->dmin * 0.5 * (1. / kernel_gamma);
  c->h_max_allowed = c->dmin * (1. / kernel_gamma);
  c->depth = 0;
  c->split = 0;
  c->hydro.count = (int)n;
  c->hydro.count_total = (int)n;
  c->hydro.parts = parts;
  c->hydro.xparts = xparts;
  c->hydro.ti_old_part = step->ti_current;
  c->nodeID = cfg->rank;
  c->parent = NULL;
  c->top = c;
  c->super = c;
  c->hydro.super = c;
  lock_init(&c->hydro.lock);
  space_split_recursive(s, c, NULL, NULL, NULL, NULL, NULL, /*tpid=*/0);
  int nc = 0;
  const int r = emit_subtree(c, parts, -1, 0, out, max_cells, &nc);
  memcpy(parts_aos, parts, n * sizeof(struct part));
  free(parts);
  free(xparts);
  free(c); /* the progeny live in the space's cell chunks (swift_ignore_leak'ed by the reference) */
  free(s->multipoles_sub);
  free(s->cells_sub);
  free(s);
  free(e);
  return r < 0 ? -1 : nc;
}

/* ---- the time-step limiter loop: runner_dosub_{self,pair}1_limiter over the
 * density tasks, after a step (sorts and h as the step left them) ---- */
int swiftref_wakeup_offset(void) { return (int)offsetof(struct part, limiter_data.wakeup); }
int swiftref_run_limiter(swiftref_t *s, int nthreads) {
  for (int i = 0; i < s->ncells; i++) /* cell_is_starting_hydro: ti_beg_max == ti_current */
    if (s->cells[i].hydro.ti_end_min == s->engine.ti_current)
      s->cells[i].hydro.ti_beg_max = s->engine.ti_current;
  run_kind(s, K_LIMITER, nthreads);
  return 0;
}

/* ---- drift: the reference's own cell_drift_part ---- */

/* offsetof() table of this build's struct xpart in swiftgpu_xpart_layout order. */
void swiftref_xpart_layout(swiftgpu_xpart_layout *X) {
  X->size = (int)sizeof(struct xpart);
  X->x_diff = (int)offsetof(struct xpart, x_diff);
  X->x_diff_sort = (int)offsetof(struct xpart, x_diff_sort);
  X->v_full = (int)offsetof(struct xpart, v_full);
#if defined(GADGET2_SPH)
  X->u_full = (int)offsetof(struct xpart, entropy_full);
#else
  X->u_full = (int)offsetof(struct xpart, u_full);
#endif
}
int swiftref_set_xparts(swiftref_t *s, const void *xparts_aos) {
  memcpy(s->xparts, xparts_aos, s->nparts * sizeof(struct xpart));
  return 0;
}
int swiftref_get_xparts(swiftref_t *s, void *xparts_aos) {
  memcpy(xparts_aos, s->xparts, s->nparts * sizeof(struct xpart));
  return 0;
}
/* runner_do_kick1 / runner_do_kick2 (src/runner_time_integration.c:87,360) on every local top-level
 * cell: the reference's own kick of the hydro particles (no gravity, no mesh). */
int swiftref_run_kick(swiftref_t *s, int which, float minimal_internal_energy) {
  static struct pm_mesh mesh;
  bzero(&mesh, sizeof(mesh));
  mesh.ti_beg_mesh_next = -1;
  mesh.ti_end_mesh_next = -1;
  mesh.ti_beg_mesh_last = -1;
  mesh.ti_end_mesh_last = -1;
  s->engine.mesh = &mesh;
  s->hp.minimal_internal_energy = minimal_internal_energy;
  struct runner r;
  bzero(&r, sizeof(r));
  r.e = &s->engine;
  for (int i = 0; i < s->ncells; i++) /* cell_is_starting_hydro: ti_beg_max == ti_current */
    if (s->cells[i].hydro.ti_end_min == s->engine.ti_current)
      s->cells[i].hydro.ti_beg_max = s->engine.ti_current;
  for (int a = 0; a < s->ntop; a++) {
    struct cell *c = &s->cells[s->top[a]];
    if (c->nodeID != s->cfg.rank) continue;
    if (which == 1)
      runner_do_kick1(&r, c, /*timer=*/0);
    else
      runner_do_kick2(&r, c, /*timer=*/0);
  }
  return 0;
}

/* cell_drift_part(c, e, force = 1, init_particles, NULL) on every local top-level cell
 * (engine_drift.c:83), all cells drifted from ti_old to the step's ti_current. */
int swiftref_run_drift(swiftref_t *s, long long ti_old, float minimal_internal_energy,
                       int init_particles) {
  s->hp.minimal_internal_energy = minimal_internal_energy;
  for (int i = 0; i < s->ncells; i++) s->cells[i].hydro.ti_old_part = ti_old;
  for (int a = 0; a < s->ntop; a++) {
    struct cell *c = &s->cells[s->top[a]];
    if (c->nodeID != s->cfg.rank) continue;
    cell_drift_part(c, &s->engine, /*force=*/1, init_particles, /*replication_list=*/NULL);
  }
  return 0;
}

/* Sort keys of one cell along one sid (sort_part.h:32), in sorted order. */
int swiftref_get_sort(swiftref_t *s, int cell, int sid, float *d, int *ind) {
  struct cell *c = &s->cells[cell];
  if (!(c->hydro.sorted & (1 << sid))) return -1;
  const struct sort_entry *e = cell_get_hydro_sorts(c, sid);
  for (int k = 0; k < c->hydro.count; k++) {
    d[k] = e[k].d;
    ind[k] = e[k].i;
  }
  return c->hydro.count;
}

/* Integer neighbour counters, available in the debugging-option builds only:
 * SPHENIX + SWIFT_HYDRO_DENSITY_CHECKS -> N_density/N_gradient/N_force
 * (hydro/SPHENIX/hydro_iact.h:121-126; the counter includes the self term
 * added by hydro_init_part? no: see hydro/SPHENIX/hydro.h:580), Gadget2 +
 * DEBUG_INTERACTIONS_SPH -> num_ngb_density/num_ngb_force. */
int swiftref_get_counts(swiftref_t *s, int *n_density, int *n_gradient,
                        int *n_force) {
#if defined(SWIFT_HYDRO_DENSITY_CHECKS) && defined(SPHENIX_SPH)
  for (long long k = 0; k < s->nparts; k++) {
    if (n_density) n_density[k] = s->parts[k].N_density;
    if (n_gradient) n_gradient[k] = s->parts[k].N_gradient;
    if (n_force) n_force[k] = s->parts[k].N_force;
  }
  return 0;
#elif defined(DEBUG_INTERACTIONS_SPH)
  for (long long k = 0; k < s->nparts; k++) {
    if (n_density) n_density[k] = s->parts[k].num_ngb_density;
    if (n_gradient) n_gradient[k] = 0;
    if (n_force) n_force[k] = s->parts[k].num_ngb_force;
  }
  return 0;
#else
  (void)s;
  (void)n_density;
  (void)n_gradient;
  (void)n_force;
  return -1;
#endif
}

/* The reference's own hydro_compute_timestep (hydro/<scheme>/hydro.h) for
 * every particle, with the hydro_props / cosmology of this engine. */
int swiftref_get_timesteps(swiftref_t *s, float *dt) {
  for (long long k = 0; k < s->nparts; k++)
    dt[k] = hydro_compute_timestep(&s->parts[k], NULL, &s->hp, &s->cosmo);
  return 0;
}

/* Neighbour ID lists (Gadget2 + DEBUG_INTERACTIONS_SPH only). out has room
 * for max_ngb ids per particle. which: 0 density, 1 force. */
int swiftref_get_ngb_ids(swiftref_t *s, int which, long long *out, int max_ngb) {
#if defined(DEBUG_INTERACTIONS_SPH)
  for (long long k = 0; k < s->nparts; k++) {
    const struct part *p = &s->parts[k];
    const int n = which ? p->num_ngb_force : p->num_ngb_density;
    const long long *ids = which ? p->ids_ngbs_force : p->ids_ngbs_density;
    for (int j = 0; j < max_ngb; j++)
      out[k * max_ngb + j] =
          (j < n && j < MAX_NUM_OF_NEIGHBOURS) ? ids[j] : -1;
  }
  return 0;
#else
  (void)s;
  (void)which;
  (void)out;
  (void)max_ngb;
  return -1;
#endif
}
