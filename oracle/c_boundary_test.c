/*
 * oracle/c_boundary_test.c - TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The drop-in boundary exercised from C, in ONE process, against the real
 * reference types: this file is compiled with the reference's own headers
 * (`struct part` of the scheme selected in oracle/_ref/<scheme>/config.h), fills
 * swiftgpu_part_layout with offsetof() on that struct - what INTEGRATION.md
 * section 2 asks a maintainer to do -, writes the initial conditions through
 * the struct's MEMBERS, hands the very same `struct part[]` to
 *   (a) libswiftgpu (upload cells/parts -> run_step -> download), and
 *   (b) the reference's runner_do_hydro_sort / runner_dosub_* / runner_do_ghost
 *       / runner_do_end_hydro_force (through oracle/ref_driver.c, linked in),
 * and compares the members the path writes (h, rho, a_hydro, u_dt|entropy_dt,
 * limiter_data.min_ngb_time_bin, ...), again through the struct.
 *
 * Needs a CUDA device (tests/test_gpu_parity.py::test_c_boundary runs the
 * binaries that oracle/Makefile `ctest` built where the reference tree was
 * present). Exit code 0 = PASS.
 */
#include <config.h>

#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swift.h"

#include "../include/swiftgpu.h"

/* oracle/ref_driver.c (compiled into the same binary) */
typedef struct swiftref swiftref_t;
swiftref_t *swiftref_create(const swiftgpu_config *cfg, const swiftgpu_step *step, const swiftgpu_cell *cells,
                            int ncells, const int *top, int ntop, const void *parts_aos, long long nparts);
int swiftref_run(swiftref_t *s, unsigned mask, int nthreads, double *seconds);
int swiftref_get_parts(swiftref_t *s, void *parts_aos);
int swiftref_set_xparts(swiftref_t *s, const void *xparts_aos);
int swiftref_get_xparts(swiftref_t *s, void *xparts_aos);
int swiftref_run_kick(swiftref_t *s, int which, float minimal_internal_energy);
int swiftref_run_drift(swiftref_t *s, long long ti_old, float minimal_internal_energy, int init_particles);
void swiftref_destroy(swiftref_t *s);

/* swift_b200/csrc/host_tree.cpp (libswiftgpu_host.so): stands in for space_regrid / space_split */
typedef struct swifthost_tree swifthost_tree;
swifthost_tree *swifthost_build_tree(const double *x, const float *h, const int8_t *time_bin, int64_t n,
                                     const double dim[3], const int cdim[3], int splitsize, int max_active_bin,
                                     int64_t ti_current, const int rank_grid[3], int64_t *perm, int8_t *depth_h);
int32_t swifthost_tree_ncells(const swifthost_tree *t);
int32_t swifthost_tree_ntop(const swifthost_tree *t);
void swifthost_tree_copy(const swifthost_tree *t, swiftgpu_cell *cells, int32_t *top);
void swifthost_tree_free(swifthost_tree *t);

#define CHECK(call)                                                                       \
  do {                                                                                    \
    if ((call) != 0) {                                                                    \
      fprintf(stderr, "FAIL %s: %s\n", #call, swiftgpu_last_error(g) ? swiftgpu_last_error(g) : "?"); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

static double frand(unsigned long long *s) { /* splitmix64 -> [0, 1) */
  unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 20;
  const long long n = (long long)L * L * L;
  swiftgpu_t *g = NULL;

  /* ---- the layout, from the real struct ---- */
  swiftgpu_config cfg;
#if defined(MINIMAL_SPH)
  const int scheme = SWIFTGPU_SCHEME_MINIMAL;
#elif defined(GADGET2_SPH)
  const int scheme = SWIFTGPU_SCHEME_GADGET2;
#else
  const int scheme = SWIFTGPU_SCHEME_SPHENIX;
#endif
  if (swiftgpu_default_config(scheme, &cfg) != 0) return 2;
  swiftgpu_part_layout *Ly = &cfg.layout;
  memset(Ly, 0xff, sizeof(*Ly));
  Ly->size = sizeof(struct part);
  Ly->id = offsetof(struct part, id);
  Ly->x = offsetof(struct part, x);
  Ly->v = offsetof(struct part, v);
  Ly->a_hydro = offsetof(struct part, a_hydro);
  Ly->mass = offsetof(struct part, mass);
  Ly->h = offsetof(struct part, h);
  Ly->rho = offsetof(struct part, rho);
  Ly->wcount = offsetof(struct part, density.wcount);
  Ly->wcount_dh = offsetof(struct part, density.wcount_dh);
  Ly->rho_dh = offsetof(struct part, density.rho_dh);
  Ly->rot_v = offsetof(struct part, density.rot_v);
  Ly->f = offsetof(struct part, force.f);
  Ly->soundspeed = offsetof(struct part, force.soundspeed);
  Ly->h_dt = offsetof(struct part, force.h_dt);
  Ly->balsara = offsetof(struct part, force.balsara);
  Ly->time_bin = offsetof(struct part, time_bin);
  Ly->depth_h = offsetof(struct part, depth_h);
  Ly->min_ngb_time_bin = offsetof(struct part, limiter_data.min_ngb_time_bin);
#if defined(MINIMAL_SPH)
  Ly->u = offsetof(struct part, u);
  Ly->u_dt = offsetof(struct part, u_dt);
  Ly->div_v = offsetof(struct part, density.div_v);
  Ly->pressure = offsetof(struct part, force.pressure);
  Ly->v_sig = offsetof(struct part, force.v_sig);
#elif defined(GADGET2_SPH)
  Ly->entropy = offsetof(struct part, entropy);
  Ly->entropy_dt = offsetof(struct part, entropy_dt);
  Ly->div_v = offsetof(struct part, density.div_v);
  Ly->P_over_rho2 = offsetof(struct part, force.P_over_rho2);
  Ly->v_sig = offsetof(struct part, force.v_sig);
#else
  Ly->u = offsetof(struct part, u);
  Ly->u_dt = offsetof(struct part, u_dt);
  Ly->div_v = offsetof(struct part, viscosity.div_v);
  Ly->pressure = offsetof(struct part, force.pressure);
  Ly->v_sig = offsetof(struct part, viscosity.v_sig);
  Ly->div_v_dt = offsetof(struct part, viscosity.div_v_dt);
  Ly->div_v_previous_step = offsetof(struct part, viscosity.div_v_previous_step);
  Ly->visc_alpha = offsetof(struct part, viscosity.alpha);
  Ly->laplace_u = offsetof(struct part, diffusion.laplace_u);
  Ly->diff_alpha = offsetof(struct part, diffusion.alpha);
  Ly->alpha_visc_max_ngb = offsetof(struct part, force.alpha_visc_max_ngb);
#endif
  cfg.h_max = 1e10f;
  swiftgpu_step step;
  memset(&step, 0, sizeof(step));
  step.ti_current = 8;
  step.max_active_bin = 56;
  step.time_base = 1e-6;
  step.a = 1.f;
  step.H = 0.f;

  /* ---- initial conditions through the members of struct part ---- */
  double *x = (double *)malloc(sizeof(double) * 3 * n);
  float *h0 = (float *)malloc(sizeof(float) * n);
  int8_t *tb = (int8_t *)malloc(n);
  unsigned long long seed = 12345;
  for (long long k = 0; k < n; k++) {
    const int i = (int)(k / ((long long)L * L)), j = (int)((k / L) % L), l = (int)(k % L);
    x[3 * k + 0] = fmod((i + 0.5 + 0.4 * (frand(&seed) - 0.5)) / L + 1.0, 1.0);
    x[3 * k + 1] = fmod((j + 0.5 + 0.4 * (frand(&seed) - 0.5)) / L + 1.0, 1.0);
    x[3 * k + 2] = fmod((l + 0.5 + 0.4 * (frand(&seed) - 0.5)) / L + 1.0, 1.0);
    h0[k] = (float)(1.2348 / L * (0.95 + 0.1 * frand(&seed)));
    tb[k] = 1;
  }
  const double dim[3] = {1., 1., 1.};
  const int cdim[3] = {3, 3, 3}, rank_grid[3] = {1, 1, 1};
  int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * n);
  int8_t *depth_h = (int8_t *)malloc(n);
  swifthost_tree *T = swifthost_build_tree(x, h0, tb, n, dim, cdim, 400, step.max_active_bin, step.ti_current,
                                           rank_grid, perm, depth_h);
  const int ncells = swifthost_tree_ncells(T), ntop = swifthost_tree_ntop(T);
  swiftgpu_cell *cells = (swiftgpu_cell *)malloc(sizeof(swiftgpu_cell) * ncells);
  int32_t *top = (int32_t *)malloc(sizeof(int32_t) * ntop);
  swifthost_tree_copy(T, cells, top);
  swifthost_tree_free(T);

  struct part *parts = NULL;
  if (posix_memalign((void **)&parts, part_align, sizeof(struct part) * n) != 0) return 2;
  memset(parts, 0, sizeof(struct part) * n);
  const double two_pi = 6.283185307179586;
  for (long long k = 0; k < n; k++) {
    struct part *p = &parts[k];
    const long long o = perm[k];
    p->id = o + 1;
    for (int d = 0; d < 3; d++) p->x[d] = x[3 * o + d];
    p->v[0] = 0.05f * (float)sin(two_pi * p->x[1]);
    p->v[1] = 0.05f * (float)sin(two_pi * p->x[2]);
    p->v[2] = 0.05f * (float)sin(two_pi * p->x[0]);
    p->mass = (float)(1.0 / (double)n);
    p->h = h0[o];
    p->time_bin = 1;
    p->depth_h = depth_h[k];
    const float u = 1.f + 0.1f * (float)sin(two_pi * p->x[0]);
#if defined(GADGET2_SPH)
    p->entropy = (float)((hydro_gamma - 1.) * u); /* rho0 = 1 */
#else
    p->u = u;
#endif
#if defined(SPHENIX_SPH)
    p->viscosity.alpha = 0.1f;
#endif
  }

  /* ---- (b) the reference, on a copy ---- */
  struct part *ref_parts = NULL;
  if (posix_memalign((void **)&ref_parts, part_align, sizeof(struct part) * n) != 0) return 2;
  swiftref_t *R = swiftref_create(&cfg, &step, cells, ncells, top, ntop, parts, n);
  if (!R) {
    fprintf(stderr, "FAIL swiftref_create\n");
    return 1;
  }
  swiftref_run(R, SWIFTGPU_PHASE_ALL, 2, NULL);
  swiftref_get_parts(R, ref_parts);

  /* ---- (a) libswiftgpu through its C ABI, on the same array ---- */
  const int rc = swiftgpu_init(&g, &cfg);
  if (rc != 0) {
    fprintf(stderr, "FAIL swiftgpu_init: %s\n", swiftgpu_last_error(NULL));
    return rc == 2 ? 77 : 1; /* 77: no CUDA device */
  }
  CHECK(swiftgpu_upload_cells(g, cells, ncells, top, ntop));
  CHECK(swiftgpu_upload_parts(g, parts, n));
  CHECK(swiftgpu_set_step(g, &step));
  CHECK(swiftgpu_run_step(g, SWIFTGPU_PHASE_ALL));
  CHECK(swiftgpu_download_parts(g, parts, n));
  swiftgpu_stats st;
  CHECK(swiftgpu_get_stats(g, &st));

  /* ---- (c) time integration through the same boundary: kick2, kick1 and the
   * drift of the reference (runner_do_kick2/1, cell_drift_part) against swiftgpu_run_kick / _drift,
   * both starting from the reference's post-step particles and the same struct xpart[] - laid out by
   * offsetof() on the reference's own type. Integer-exact members are compared with memcmp. ---- */
  struct xpart *xp = NULL, *ref_xp = NULL;
  struct part *kd_parts = NULL, *ref_kd = NULL;
  if (posix_memalign((void **)&xp, xpart_align, sizeof(struct xpart) * n) != 0) return 2;
  if (posix_memalign((void **)&ref_xp, xpart_align, sizeof(struct xpart) * n) != 0) return 2;
  if (posix_memalign((void **)&kd_parts, part_align, sizeof(struct part) * n) != 0) return 2;
  if (posix_memalign((void **)&ref_kd, part_align, sizeof(struct part) * n) != 0) return 2;
  memset(xp, 0, sizeof(struct xpart) * n);
  for (long long k = 0; k < n; k++) {
    for (int d = 0; d < 3; d++) {
      xp[k].v_full[d] = ref_parts[k].v[d];
      xp[k].x_diff[d] = 1e-4f * (float)(k % 7 - 3);
      xp[k].x_diff_sort[d] = -1e-4f * (float)(k % 5 - 2);
    }
#if defined(GADGET2_SPH)
    xp[k].entropy_full = ref_parts[k].entropy;
#else
    xp[k].u_full = ref_parts[k].u;
#endif
  }
  const long long ti_span = 4; /* time bin 1: the whole step of every particle */
  swiftgpu_step kstep = step;
  kstep.time_base = 2.5e-4; /* dt = 1e-3 */
  swiftref_destroy(R);
  R = swiftref_create(&cfg, &kstep, cells, ncells, top, ntop, ref_parts, n);
  if (!R) return 1;
  swiftref_set_xparts(R, xp);
  swiftref_run_kick(R, 2, 0.f);
  swiftref_run_kick(R, 1, 0.f);
  swiftref_run_drift(R, kstep.ti_current - ti_span, 0.f, 1);
  swiftref_get_parts(R, ref_kd);
  swiftref_get_xparts(R, ref_xp);
  swiftref_destroy(R);

  swiftgpu_xpart_layout X;
  X.size = (int32_t)sizeof(struct xpart);
  X.x_diff = (int32_t)offsetof(struct xpart, x_diff);
  X.x_diff_sort = (int32_t)offsetof(struct xpart, x_diff_sort);
  X.v_full = (int32_t)offsetof(struct xpart, v_full);
#if defined(GADGET2_SPH)
  X.u_full = (int32_t)offsetof(struct xpart, entropy_full);
#else
  X.u_full = (int32_t)offsetof(struct xpart, u_full);
#endif
  CHECK(swiftgpu_upload_parts(g, ref_parts, n));
  CHECK(swiftgpu_set_step(g, &kstep));
  CHECK(swiftgpu_upload_xparts(g, &X, xp, n));
  CHECK(swiftgpu_run_kick(g, 2, 0.f));
  CHECK(swiftgpu_run_kick(g, 1, 0.f));
  swiftgpu_drift_args da_;
  da_.dt_drift = da_.dt_kick_hydro = da_.dt_therm = (double)ti_span * kstep.time_base;
  da_.minimal_internal_energy = 0.f;
  da_.init_particles = 1;
  CHECK(swiftgpu_run_drift(g, &da_));
  CHECK(swiftgpu_download_parts(g, kd_parts, n));
  CHECK(swiftgpu_download_xparts(g, xp, n));
  swiftgpu_destroy(g);
  long long bad_x = 0, bad_v = 0, bad_xp = 0, bad_h = 0;
  double moved = 0;
  for (long long k = 0; k < n; k++) {
    if (memcmp(kd_parts[k].x, ref_kd[k].x, sizeof(kd_parts[k].x)) != 0) bad_x++;
    if (memcmp(kd_parts[k].v, ref_kd[k].v, sizeof(kd_parts[k].v)) != 0) bad_v++;
    if (memcmp(&kd_parts[k].h, &ref_kd[k].h, sizeof(float)) != 0) bad_h++;
    if (memcmp(xp[k].v_full, ref_xp[k].v_full, sizeof(xp[k].v_full)) != 0 ||
        memcmp(xp[k].x_diff, ref_xp[k].x_diff, sizeof(xp[k].x_diff)) != 0 ||
        memcmp(xp[k].x_diff_sort, ref_xp[k].x_diff_sort, sizeof(xp[k].x_diff_sort)) != 0)
      bad_xp++;
    moved = fmax(moved, fabs(ref_kd[k].x[0] - ref_parts[k].x[0]));
  }
  printf("c_boundary_test kick2 + kick1 + drift: moved %.3e, mismatching x %lld, v %lld, h %lld, xpart %lld of %lld\n",
         moved, bad_x, bad_v, bad_h, bad_xp, n);
  const int ok_ti = moved > 0 && bad_x == 0 && bad_v == 0 && bad_h == 0 && bad_xp == 0;

  /* ---- compare through the struct members ---- */
  double e_h = 0, e_rho = 0, e_a = 0, e_u = 0;
  long long flips = 0, bad_bin = 0;
  double a_scale = 0;
  for (long long k = 0; k < n; k++) {
    const double na = sqrt((double)ref_parts[k].a_hydro[0] * ref_parts[k].a_hydro[0] +
                           (double)ref_parts[k].a_hydro[1] * ref_parts[k].a_hydro[1] +
                           (double)ref_parts[k].a_hydro[2] * ref_parts[k].a_hydro[2]);
    a_scale += na / (double)n;
  }
  for (long long k = 0; k < n; k++) {
    const struct part *a = &parts[k], *b = &ref_parts[k];
    const double dh = fabs((double)a->h - b->h) / b->h;
    if (dh > 2e-6) { /* the ghost accepted h one Newton step apart (tests/util.py): counted, not compared */
      flips++;
      continue;
    }
    e_h = fmax(e_h, dh);
    e_rho = fmax(e_rho, fabs((double)a->rho - b->rho) / b->rho);
    double da = 0;
    for (int d = 0; d < 3; d++) da += ((double)a->a_hydro[d] - b->a_hydro[d]) * ((double)a->a_hydro[d] - b->a_hydro[d]);
    e_a = fmax(e_a, sqrt(da) / a_scale);
#if defined(GADGET2_SPH)
    e_u = fmax(e_u, fabs((double)a->entropy_dt - b->entropy_dt));
#else
    e_u = fmax(e_u, fabs((double)a->u_dt - b->u_dt));
#endif
    if (a->limiter_data.min_ngb_time_bin != b->limiter_data.min_ngb_time_bin) bad_bin++;
  }
  printf("c_boundary_test scheme=%d sizeof(struct part)=%zu n=%lld cells=%d: interactions density=%lld force=%lld "
         "ghost_iterations=%d | max rel err h %.2e rho %.2e, |da| / <|a|> %.2e, |d u_dt| %.2e, flips %lld, "
         "min_ngb_time_bin mismatches %lld\n",
         scheme, sizeof(struct part), n, ncells, (long long)st.n_density, (long long)st.n_force, st.ghost_iterations,
         e_h, e_rho, e_a, e_u, flips, bad_bin);
  /* flipped particles' neighbours see a 1e-4 different h: bars a decade above the clean-particle bar */
  const int ok = e_h < 1e-5 && e_rho < 1e-4 && e_a < 1e-3 && bad_bin == 0 && flips <= 3 + n / 2000 &&
                 st.n_density > 0 && st.n_force > 0 && ok_ti;
  printf(ok ? "C_BOUNDARY PASS\n" : "C_BOUNDARY FAIL\n");
  return ok ? 0 : 1;
}
