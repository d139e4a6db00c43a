/*
 * oracle/swift_port.c - TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement ("port") of the reference's SPH neighbour-interaction
 * path, written from the reference's behaviour (SWIFT 2026.04). It is pinned
 * against the real reference (oracle/_ref, built from the unmodified sources)
 * by tests/test_oracle.py: identical integer neighbour counts, float fields
 * within summation-order tolerance.
 *
 * Differences from the reference in FORM (not in result): every loop is
 * restated as a GATHER - each updatable particle collects the contributions of
 * its neighbours with the non-symmetric interaction, which the reference's own
 * tests/testSymmetry.c shows to be bit-identical per interaction to the
 * symmetric update. The sorted-axis pruning of DOPAIR1/DOPAIR2 is restated as
 * a per-(i,j) predicate on the same float sort keys, so the neighbour SET is
 * the reference's, including its 1-ulp edge behaviour.
 *
 * Build: make -C oracle port  (one .so per scheme, -DPORT_SCHEME_<NAME>=1).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/swiftgpu.h"

/* Numeric (the enum constants of swiftgpu.h are invisible to #if). */
#define SCH_MINIMAL 0
#define SCH_GADGET2 1
#define SCH_SPHENIX 2
#if defined(PORT_SCHEME_MINIMAL)
#define PORT_SCHEME SCH_MINIMAL
#elif defined(PORT_SCHEME_GADGET2)
#define PORT_SCHEME SCH_GADGET2
#elif defined(PORT_SCHEME_SPHENIX)
#define PORT_SCHEME SCH_SPHENIX
#else
#error "define PORT_SCHEME_MINIMAL|GADGET2|SPHENIX"
#endif

/* ---- constants: kernel_hydro.h:41-55,205-237 (cubic spline, 3D) ---- */
#define kernel_gamma ((float)(1.825742))
#define kernel_constant ((float)(16. * M_1_PI))
#define kernel_gamma_inv ((float)(1. / kernel_gamma))
#define kernel_gamma2 ((float)(kernel_gamma * kernel_gamma))
#define kernel_gamma_inv_dim \
  ((float)(1. / (kernel_gamma * kernel_gamma * kernel_gamma)))
#define kernel_gamma_inv_dim_plus_one \
  ((float)(1. / (kernel_gamma * kernel_gamma * kernel_gamma * kernel_gamma)))
#define kernel_degree 3
#define kernel_ivals 2
#define kernel_ivals_f ((float)(kernel_ivals))
/* kernel_hydro.h:61-66: coefficients of the two branches + the zero branch */
static const float kernel_coeffs[(kernel_degree + 1) * (kernel_ivals + 1)] = {
    3.f,  -3.f, 0.f,  0.5f, /* 0 < u < 0.5 */
    -1.f, 3.f,  -3.f, 1.f,  /* 0.5 < u < 1 */
    0.f,  0.f,  0.f,  0.f}; /* 1 < u */
#define kernel_root \
  ((float)(kernel_coeffs[kernel_degree]) * kernel_constant * kernel_gamma_inv_dim)
#define hydro_dimension 3.f
#define hydro_dimension_inv 0.3333333333f
/* adiabatic_index.h:43-46 (gamma = 5/3) */
#define hydro_gamma 1.66666666666666667f
#define hydro_gamma_minus_one 0.66666666666666667f
#define const_viscosity_beta 3.0f
/* space.h:64-65,75 */
#define space_recurse_size_self_hydro 100
#define space_recurse_size_pair_hydro 100
#define time_bin_inhibited (56 + 2)
#define num_time_bins 56

#define pmin(a, b) ((a) < (b) ? (a) : (b))
#define pmax(a, b) ((a) > (b) ? (a) : (b))

/* sort_part.h:42-58 */
static const double runner_shift[13][3] = {
    {5.773502691896258e-01, 5.773502691896258e-01, 5.773502691896258e-01},
    {7.071067811865475e-01, 7.071067811865475e-01, 0.0},
    {5.773502691896258e-01, 5.773502691896258e-01, -5.773502691896258e-01},
    {7.071067811865475e-01, 0.0, 7.071067811865475e-01},
    {1.0, 0.0, 0.0},
    {7.071067811865475e-01, 0.0, -7.071067811865475e-01},
    {5.773502691896258e-01, -5.773502691896258e-01, 5.773502691896258e-01},
    {7.071067811865475e-01, -7.071067811865475e-01, 0.0},
    {5.773502691896258e-01, -5.773502691896258e-01, -5.773502691896258e-01},
    {0.0, 7.071067811865475e-01, 7.071067811865475e-01},
    {0.0, 1.0, 0.0},
    {0.0, 7.071067811865475e-01, -7.071067811865475e-01},
    {0.0, 0.0, 1.0},
};

enum { LOOP_DENSITY = 0, LOOP_GRADIENT = 1, LOOP_FORCE = 2 };

typedef struct port {
  swiftgpu_config cfg;
  swiftgpu_step step;
  const swiftgpu_cell *cells_in;
  swiftgpu_cell *cells;
  int ncells;
  const int *top;
  int ntop;
  long long n;
  /* SoA state */
  double *x;
  float *v, *a, *rot_v;
  float *m, *h, *u, *u_dt, *rho, *wcount, *wcount_dh, *rho_dh, *div_v;
  float *f, *P, *cs, *balsara, *v_sig, *h_dt;
  float *alpha, *alpha_diff, *div_v_prev, *div_v_dt, *laplace_u, *alpha_max_ngb;
  signed char *time_bin, *depth_h, *min_ngb;
  int *nd, *ng, *nf;
  /* un-cancelled sums (sum of |pair terms|) behind the cancelling outputs: the floors of the parity
   * metric (tests/util.py), carried through the same finalising factors as the sums themselves */
  float *g_a, *g_u, *g_hdt, *g_div, *g_rho_dh, *g_lap;
  float *g2_a, *g2_u, *g2_lap; /* squares: kernel-evaluation noise, see edge_weight() */
  /* absolute noise of the inputs that are themselves ratios of cancelling sums: the Balsara switch and
   * (SPHENIX) the diffusion alpha, at the size the metric allows them (1e-5 of their floors) */
  float *nz_B, *nz_ad, *nz_al;
  int *leaf_of; /* leaf cell index of each particle */
  int ghost_iterations;
  long long ghost_redo[32]; /* particles sent to re-run k of the ghost (diagnostic) */
  int ghost_failed;
} port_t;

/* ---------------- activity predicates: active.h:176,349,518 ---------------- */
static int part_active(const port_t *s, long long p) {
  return s->time_bin[p] <= s->step.max_active_bin;
}
static int part_inhibited(const port_t *s, long long p) {
  return s->time_bin[p] == time_bin_inhibited;
}
static int cell_active(const port_t *s, const swiftgpu_cell *c) {
  return c->ti_end_min == s->step.ti_current;
}

/* ---------------- kernel_deval: kernel_hydro.h:257-285 ---------------- */
static inline void kernel_deval(float u, float *W, float *dW_dx) {
  const float x = u * kernel_gamma_inv;
  const int temp = (int)(x * kernel_ivals_f);
  const int ind = temp > kernel_ivals ? kernel_ivals : temp;
  const float *const coeffs = &kernel_coeffs[ind * (kernel_degree + 1)];
  float w = coeffs[0] * x + coeffs[1];
  float dw_dx = coeffs[0];
  for (int k = 2; k <= kernel_degree; k++) {
    dw_dx = dw_dx * x + w;
    w = x * w + coeffs[k];
  }
  w = pmax(w, 0.f);
  dw_dx = pmin(dw_dx, 0.f);
  *W = w * kernel_constant * kernel_gamma_inv_dim;
  *dW_dx = dw_dx * kernel_constant * kernel_gamma_inv_dim_plus_one;
}
void port_kernel_deval(float u, float *w, float *dw) { kernel_deval(u, w, dw); }

/* ---------------- pair interactions (non-symmetric) ---------------- */

/* Kernel-evaluation noise of a pair term, for the sums in which ONE neighbour can dominate (a_hydro,
 * u_dt, laplace_u: weighted by the neighbour's pressure / energy). kernel_deval (kernel_hydro.h:257-285)
 * evaluates the outer branch of the cubic spline, w' = -3 (1-x)^2, by Horner's rule on the EXPANDED
 * coefficients {-1, 3, -3, 1} in FP32: the absolute error is ~4e-7 of the coefficient scale whatever x
 * is, i.e. 1.3e-7 / (1-x)^2 RELATIVE to w' - 2e-4 at q = 0.97 - and it jumps with the last bit of h or r.
 * These errors are independent between pairs, so they are accumulated in quadrature: g2 = sum (|term| /
 * (1-q)^2)^2. For ordinary neighbourhoods sqrt(g2) stays below the plain un-cancelled sum; a neighbour
 * at the kernel edge that is orders of magnitude hotter (the shell around a Sedov blast) dominates both
 * the sum and g2 and carries its 1e-4 into them - in the reference exactly as here. */
static inline float edge_weight(float r, float h) {
  const float q = r / (h * kernel_gamma);
  const float e = 1.f - q;
  return e > 1.e-3f ? 1.f / (e * e) : 1.e6f;
}

/* runner_iact_nonsym_density: Minimal hydro_iact.h:137, Gadget2 :158,
 * SPHENIX :141 (identical arithmetic; div_v lives in viscosity.div_v there) */
static inline void iact_density(port_t *s, float r2, const float dx[3], float hi,
                                long long i, long long j) {
  float wi, wi_dx;
  const float mj = s->m[j];
  const float r = sqrtf(r2);
  const float r_inv = r ? 1.0f / r : 0.0f;
  const float h_inv = 1.f / hi;
  const float ui = r * h_inv;
  kernel_deval(ui, &wi, &wi_dx);
  s->rho[i] += mj * wi;
  s->rho_dh[i] -= mj * (hydro_dimension * wi + ui * wi_dx);
  s->wcount[i] += wi;
  s->wcount_dh[i] -= (hydro_dimension * wi + ui * wi_dx);
  const float faci = mj * wi_dx * r_inv;
  float dv[3], curlvr[3];
  dv[0] = s->v[3 * i + 0] - s->v[3 * j + 0];
  dv[1] = s->v[3 * i + 1] - s->v[3 * j + 1];
  dv[2] = s->v[3 * i + 2] - s->v[3 * j + 2];
  const float dvdr = dv[0] * dx[0] + dv[1] * dx[1] + dv[2] * dx[2];
  s->div_v[i] -= faci * dvdr;
  s->g_div[i] += fabsf(faci * dvdr);
  s->g_rho_dh[i] += fabsf(mj * (hydro_dimension * wi + ui * wi_dx));
  curlvr[0] = dv[1] * dx[2] - dv[2] * dx[1];
  curlvr[1] = dv[2] * dx[0] - dv[0] * dx[2];
  curlvr[2] = dv[0] * dx[1] - dv[1] * dx[0];
  s->rot_v[3 * i + 0] += faci * curlvr[0];
  s->rot_v[3 * i + 1] += faci * curlvr[1];
  s->rot_v[3 * i + 2] += faci * curlvr[2];
  s->nd[i]++;
}

#if PORT_SCHEME == SCH_SPHENIX
/* runner_iact_nonsym_gradient: SPHENIX hydro_iact.h:291 (a=1, H=0) */
static inline void iact_gradient(port_t *s, float r2, const float dx[3], float hi,
                                 long long i, long long j) {
  const float r = sqrtf(r2);
  const float r_inv = r ? 1.0f / r : 0.0f;
  const float fac_mu = 1.f;
  const float a2_Hubble = s->step.a * s->step.a * s->step.H;
  const float dvdr = (s->v[3 * i + 0] - s->v[3 * j + 0]) * dx[0] +
                     (s->v[3 * i + 1] - s->v[3 * j + 1]) * dx[1] +
                     (s->v[3 * i + 2] - s->v[3 * j + 2]) * dx[2];
  const float dvdr_Hubble = dvdr + a2_Hubble * r2;
  const float omega_ij = pmin(dvdr_Hubble, 0.f);
  const float mu_ij = fac_mu * r_inv * omega_ij;
  const float new_v_sig = s->cs[i] + s->cs[j] - const_viscosity_beta * mu_ij;
  s->v_sig[i] = pmax(s->v_sig[i], new_v_sig);
  float wi, wi_dx;
  const float ui = r / hi;
  kernel_deval(ui, &wi, &wi_dx);
  const float delta_u_factor = (s->u[i] - s->u[j]) * r_inv;
  s->laplace_u[i] += s->m[j] * delta_u_factor * wi_dx / s->rho[j];
  {
    const float t = fabsf(s->m[j] * delta_u_factor * wi_dx / s->rho[j]);
    s->g_lap[i] += t;
    s->g2_lap[i] += (t * edge_weight(r, hi)) * (t * edge_weight(r, hi));
  }
  const float alpha_j = s->alpha[j];
  s->alpha_max_ngb[i] = pmax(s->alpha_max_ngb[i], alpha_j);
  s->ng[i]++;
}
#endif

/* runner_iact_nonsym_force: Minimal hydro_iact.h:378, Gadget2 :632,
 * SPHENIX :507; + runner_iact_nonsym_timebin (timestep_limiter_iact.h:41-55) */
static inline void iact_force(port_t *s, float r2, const float dx[3], float hi,
                              float hj, long long i, long long j) {
  const float fac_mu = 1.f; /* pow_three_gamma_minus_five_over_two(a), gamma=5/3 */
  const float a2_Hubble = s->step.a * s->step.a * s->step.H;
  const float r = sqrtf(r2);
  const float r_inv = r ? 1.0f / r : 0.0f;
  const float mj = s->m[j];
  const float rhoi = s->rho[i];
  const float rhoj = s->rho[j];
  const float hi_inv = 1.0f / hi;
  const float hid_inv = hi_inv * hi_inv * hi_inv * hi_inv; /* pow_dimension_plus_one */
  const float xi = r * hi_inv;
  float wi, wi_dx;
  kernel_deval(xi, &wi, &wi_dx);
  const float wi_dr = hid_inv * wi_dx;
  const float hj_inv = 1.0f / hj;
  const float hjd_inv = hj_inv * hj_inv * hj_inv * hj_inv;
  const float xj = r * hj_inv;
  float wj, wj_dx;
  kernel_deval(xj, &wj, &wj_dx);
  const float wj_dr = hjd_inv * wj_dx;
  const float dvdr = (s->v[3 * i + 0] - s->v[3 * j + 0]) * dx[0] +
                     (s->v[3 * i + 1] - s->v[3 * j + 1]) * dx[1] +
                     (s->v[3 * i + 2] - s->v[3 * j + 2]) * dx[2];
  const float dvdr_Hubble = dvdr + a2_Hubble * r2;
  const float omega_ij = pmin(dvdr_Hubble, 0.f);
  const float mu_ij = fac_mu * r_inv * omega_ij;
  const float v_sig = s->cs[i] + s->cs[j] - const_viscosity_beta * mu_ij;
  const float ew = fmaxf(wi_dx != 0.f ? edge_weight(r, hi) : 1.f, wj_dx != 0.f ? edge_weight(r, hj) : 1.f);
  const float balsara_i = s->balsara[i];
  const float balsara_j = s->balsara[j];
#if PORT_SCHEME == SCH_MINIMAL
  const float mi = s->m[i];
  const float pressurei = s->P[i];
  const float pressurej = s->P[j];
  const float f_ij = 1.f - s->f[i] / mj;
  const float f_ji = 1.f - s->f[j] / mi;
  const float P_over_rho2_i = pressurei / (rhoi * rhoi) * f_ij;
  const float P_over_rho2_j = pressurej / (rhoj * rhoj) * f_ji;
  const float rho_ij = 0.5f * (rhoi + rhoj);
  const float visc = -0.25f * v_sig * (balsara_i + balsara_j) * mu_ij / rho_ij;
  const float visc_acc_term = 0.5f * visc * (wi_dr * f_ij + wj_dr * f_ji) * r_inv;
  const float sph_acc_term = (P_over_rho2_i * wi_dr + P_over_rho2_j * wj_dr) * r_inv;
  const float acc = sph_acc_term + visc_acc_term + 0.f;
  s->a[3 * i + 0] -= mj * acc * dx[0];
  s->a[3 * i + 1] -= mj * acc * dx[1];
  s->a[3 * i + 2] -= mj * acc * dx[2];
  const float sph_du_term_i = P_over_rho2_i * dvdr * r_inv * wi_dr;
  const float visc_du_term = 0.5f * visc_acc_term * dvdr_Hubble;
  const float du_dt_i = sph_du_term_i + visc_du_term;
  s->u_dt[i] += du_dt_i * mj;
  s->h_dt[i] -= mj * dvdr * r_inv / rhoj * wi_dr * f_ij;
  s->v_sig[i] = pmax(s->v_sig[i], v_sig);
  s->g_a[i] += fabsf(mj * acc) * r;
  s->g2_a[i] += (fabsf(mj * acc) * r * ew) * (fabsf(mj * acc) * r * ew);
  {
    const float t = (fabsf(sph_du_term_i) + fabsf(visc_du_term)) * mj;
    s->g_u[i] += t;
    s->g2_u[i] += (t * ew) * (t * ew);
  }
  {
    /* viscous terms carry the noise of the two Balsara switches */
    const float wb = 2.e6f * (s->nz_B[i] + s->nz_B[j]) / fmaxf(balsara_i + balsara_j, 1.e-30f);
    s->g2_a[i] += (fabsf(mj * visc_acc_term) * r * wb) * (fabsf(mj * visc_acc_term) * r * wb);
    s->g2_u[i] += (fabsf(mj * visc_du_term) * wb) * (fabsf(mj * visc_du_term) * wb);
  }
  s->g_hdt[i] += fabsf(mj * dvdr * r_inv / rhoj * wi_dr * f_ij) ;
#elif PORT_SCHEME == SCH_GADGET2
  const float f_i = s->f[i];
  const float f_j = s->f[j];
  const float P_over_rho2_i = s->P[i];
  const float P_over_rho2_j = s->P[j];
  const float rho_ij = 0.5f * (rhoi + rhoj);
  const float visc = -0.25f * v_sig * mu_ij * (balsara_i + balsara_j) / rho_ij;
  const float visc_term = 0.5f * visc * (wi_dr + wj_dr) * r_inv;
  const float sph_term =
      (f_i * P_over_rho2_i * wi_dr + f_j * P_over_rho2_j * wj_dr) * r_inv;
  const float acc = visc_term + sph_term + 0.f;
  s->a[3 * i + 0] -= mj * acc * dx[0];
  s->a[3 * i + 1] -= mj * acc * dx[1];
  s->a[3 * i + 2] -= mj * acc * dx[2];
  s->h_dt[i] -= mj * dvdr * r_inv / rhoj * wi_dr;
  s->v_sig[i] = pmax(s->v_sig[i], v_sig);
  s->u_dt[i] += mj * visc_term * dvdr_Hubble; /* entropy_dt */
  s->g_a[i] += fabsf(mj * acc) * r;
  s->g2_a[i] += (fabsf(mj * acc) * r * ew) * (fabsf(mj * acc) * r * ew);
  /* entropy_dt only collects the viscous heating; the scale of the thermal energy equation it belongs
   * to also holds the adiabatic term P/rho^2 dv.dx W'/r (x2: hydro_end_force halves the sum) */
  {
    const float t = fabsf(mj * visc_term * dvdr_Hubble) + 2.f * fabsf(mj * f_i * P_over_rho2_i * dvdr * r_inv * wi_dr);
    s->g_u[i] += t;
    s->g2_u[i] += (t * ew) * (t * ew);
  }
  {
    const float wb = 2.e6f * (s->nz_B[i] + s->nz_B[j]) / fmaxf(balsara_i + balsara_j, 1.e-30f);
    s->g2_a[i] += (fabsf(mj * visc_term) * r * wb) * (fabsf(mj * visc_term) * r * wb);
    s->g2_u[i] += (fabsf(mj * visc_term * dvdr_Hubble) * wb) * (fabsf(mj * visc_term * dvdr_Hubble) * wb);
  }
  s->g_hdt[i] += fabsf(mj * dvdr * r_inv / rhoj * wi_dr) ;
#else /* SPHENIX */
  const float mi = s->m[i];
  const float pressurei = s->P[i];
  const float pressurej = s->P[j];
  const float f_ij = 1.f - s->f[i] / mj;
  const float f_ji = 1.f - s->f[j] / mi;
  const float rho_ij = rhoi + rhoj;
  const float alpha = s->alpha[i] + s->alpha[j];
  const float visc = -0.25f * alpha * v_sig * mu_ij * (balsara_i + balsara_j) / rho_ij;
  const float visc_acc_term = 0.5f * visc * (wi_dr * f_ij + wj_dr * f_ji) * r_inv;
  const float P_over_rho2_i = pressurei / (rhoi * rhoi) * f_ij;
  const float P_over_rho2_j = pressurej / (rhoj * rhoj) * f_ji;
  const float sph_acc_term = (P_over_rho2_i * wi_dr + P_over_rho2_j * wj_dr) * r_inv;
  const float acc = sph_acc_term + visc_acc_term + 0.f;
  s->a[3 * i + 0] -= mj * acc * dx[0];
  s->a[3 * i + 1] -= mj * acc * dx[1];
  s->a[3 * i + 2] -= mj * acc * dx[2];
  const float sph_du_term_i = P_over_rho2_i * dvdr * r_inv * wi_dr;
  const float visc_du_term = 0.5f * visc_acc_term * dvdr_Hubble;
  const float alpha_diff =
      (pressurei * s->alpha_diff[i] + pressurej * s->alpha_diff[j]) /
      (pressurei + pressurej);
  const float v_diff = alpha_diff * 0.5f *
                       (sqrtf(2.f * fabsf(pressurei - pressurej) / rho_ij) +
                        fabsf(fac_mu * r_inv * dvdr_Hubble));
  const float diff_du_term =
      v_diff * (s->u[i] - s->u[j]) * (f_ij * wi_dr / rhoi + f_ji * wj_dr / rhoj);
  const float du_dt_i = sph_du_term_i + visc_du_term + diff_du_term;
  s->u_dt[i] += du_dt_i * mj;
  s->h_dt[i] -= mj * dvdr * r_inv / rhoj * wi_dr;
  s->g_a[i] += fabsf(mj * acc) * r;
  s->g2_a[i] += (fabsf(mj * acc) * r * ew) * (fabsf(mj * acc) * r * ew);
  {
    const float t = (fabsf(sph_du_term_i) + fabsf(visc_du_term) + fabsf(diff_du_term)) * mj;
    s->g_u[i] += t;
    s->g2_u[i] += (t * ew) * (t * ew);
  }
  {
    const float wb = 2.e6f * (s->nz_B[i] + s->nz_B[j]) / fmaxf(balsara_i + balsara_j, 1.e-30f);
    s->g2_a[i] += (fabsf(mj * visc_acc_term) * r * wb) * (fabsf(mj * visc_acc_term) * r * wb);
    s->g2_u[i] += (fabsf(mj * visc_du_term) * wb) * (fabsf(mj * visc_du_term) * wb);
    /* ... of the two viscosity alphas (alpha_loc = alpha_max S / (c^2 + S) is all noise in cold gas) */
    const float wa = 2.e6f * (s->nz_al[i] + s->nz_al[j]) / fmaxf(alpha, 1.e-30f);
    s->g2_a[i] += (fabsf(mj * visc_acc_term) * r * wa) * (fabsf(mj * visc_acc_term) * r * wa);
    s->g2_u[i] += (fabsf(mj * visc_du_term) * wa) * (fabsf(mj * visc_du_term) * wa);
    /* ... and the diffusion term that of the diffusion alphas */
    const float wd = 2.e6f * fmaxf(s->nz_ad[i], s->nz_ad[j]);
    s->g2_u[i] += (fabsf(mj * diff_du_term) * wd) * (fabsf(mj * diff_du_term) * wd);
  }
  s->g_hdt[i] += fabsf(mj * dvdr * r_inv / rhoj * wi_dr) ;
#endif
  /* runner_iact_nonsym_timebin */
  if (s->time_bin[j] > 0) s->min_ngb[i] = pmin(s->min_ngb[i], s->time_bin[j]);
  s->nf[i]++;
}

static inline void interact(port_t *s, int loop, float r2, const float dx[3],
                            long long i, long long j) {
  if (loop == LOOP_DENSITY)
    iact_density(s, r2, dx, s->h[i], i, j);
#if PORT_SCHEME == SCH_SPHENIX
  else if (loop == LOOP_GRADIENT)
    iact_gradient(s, r2, dx, s->h[i], i, j);
#endif
  else if (loop == LOOP_FORCE)
    iact_force(s, r2, dx, s->h[i], s->h[j], i, j);
}

/* ---------------- space_getsid_and_swap_cells: space_getsid.h:47-80 -------- */
static const int sortlistID[27] = {0, 1, 2, 3,  4,  5,  6, 7, 8, 9, 10, 11, 12, 0,
                                   12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0};
static int getsid(const port_t *s, int *ci, int *cj, double shift[3]) {
  const swiftgpu_cell *a = &s->cells[*ci], *b = &s->cells[*cj];
  double dx[3];
  for (int k = 0; k < 3; k++) {
    dx[k] = b->loc[k] - a->loc[k];
    if (s->cfg.periodic && dx[k] < -s->cfg.dim[k] / 2)
      shift[k] = s->cfg.dim[k];
    else if (s->cfg.periodic && dx[k] > s->cfg.dim[k] / 2)
      shift[k] = -s->cfg.dim[k];
    else
      shift[k] = 0.0;
    dx[k] += shift[k];
  }
  int sid = 0;
  for (int k = 0; k < 3; k++)
    sid = 3 * sid + ((dx[k] < 0.0) ? 0 : ((dx[k] > 0.0) ? 2 : 1));
  if (sid < 13) { /* runner_flip[sid] */
    int t = *ci;
    *ci = *cj;
    *cj = t;
    for (int k = 0; k < 3; k++) shift[k] = -shift[k];
  }
  return sortlistID[sid];
}

/* sort key of runner_do_hydro_sort: runner_sort.c:411-413 */
static inline float sort_key(const port_t *s, long long p, int sid) {
  const double *px = &s->x[3 * p];
  return (float)(px[0] * runner_shift[sid][0] + px[1] * runner_shift[sid][1] +
                 px[2] * runner_shift[sid][2]);
}

/* ---------------- DOSELF1 / DOSELF2 in gather form ----------------
 * functions_hydro.h:2299-2569 (self1) and :2624-2875 (self2). */
static void doself(port_t *s, int loop, int c_, int limit_min_h, int limit_max_h) {
  const swiftgpu_cell *c = &s->cells[c_];
  if (!cell_active(s, c)) return;
  if (loop == LOOP_FORCE && c->nodeID != s->cfg.rank) return;
  const int min_depth = limit_max_h ? c->depth : 0;
  const int max_depth = limit_min_h ? c->depth : CHAR_MAX;
  const long long p0 = c->first_part;
  for (int it = 0; it < c->count; it++) {
    const long long i = p0 + it;
    if (part_inhibited(s, i)) continue;
    if (!(part_active(s, i) && s->depth_h[i] >= min_depth && s->depth_h[i] <= max_depth))
      continue;
    const float hi = s->h[i];
    const float hig2 = hi * hi * kernel_gamma2;
    for (int jt = 0; jt < c->count; jt++) {
      const long long j = p0 + jt;
      if (j == i || part_inhibited(s, j)) continue;
      const float hj = s->h[j];
      const float hjg2 = hj * hj * kernel_gamma2;
      /* (float)(pix - pjx) on doubles, :2459; sign-symmetric */
      const float dx[3] = {(float)(s->x[3 * i] - s->x[3 * j]),
                           (float)(s->x[3 * i + 1] - s->x[3 * j + 1]),
                           (float)(s->x[3 * i + 2] - s->x[3 * j + 2])};
      const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
      const int hit = (loop == LOOP_FORCE) ? (r2 < hig2 || r2 < hjg2) : (r2 < hig2);
      if (hit) interact(s, loop, r2, dx, i, j);
    }
  }
}

/* ---------------- DOPAIR1 in gather form: functions_hydro.h:1234-1536 -------- */
static void dopair1(port_t *s, int loop, int ci_, int cj_, int sid,
                    const double shift[3], int limit_min_h, int limit_max_h) {
  const swiftgpu_cell *ci = &s->cells[ci_], *cj = &s->cells[cj_];
  double rshift = 0.0;
  for (int k = 0; k < 3; k++) rshift += shift[k] * runner_shift[sid][k];
  const int min_depth = limit_max_h ? ci->depth : 0;
  const int max_depth = limit_min_h ? ci->depth : CHAR_MAX;
  const float h_max = limit_max_h ? ci->h_max_allowed : FLT_MAX;
  const double hi_max = pmin(h_max, ci->h_max_active) * kernel_gamma - rshift;
  const double hj_max = pmin(h_max, cj->h_max_active) * kernel_gamma;
  const int count_i = ci->count, count_j = cj->count;
  const long long pi0 = ci->first_part, pj0 = cj->first_part;
  float *di_key = (float *)malloc(sizeof(float) * count_i);
  float *dj_key = (float *)malloc(sizeof(float) * count_j);
  float kmax_i = -FLT_MAX, kmin_j = FLT_MAX;
  for (int k = 0; k < count_i; k++) {
    di_key[k] = sort_key(s, pi0 + k, sid);
    kmax_i = pmax(kmax_i, di_key[k]);
  }
  for (int k = 0; k < count_j; k++) {
    dj_key[k] = sort_key(s, pj0 + k, sid);
    kmin_j = pmin(kmin_j, dj_key[k]);
  }
  const double di_max = kmax_i - rshift;
  const double dj_min = kmin_j;
  const float dx_max = (ci->dx_max_sort + cj->dx_max_sort);

  if (cell_active(s, ci)) {
    for (int a = 0; a < count_i; a++) {
      const long long i = pi0 + a;
      if (!(di_key[a] + hi_max + dx_max > dj_min)) continue; /* loop bound :1301 */
      if (!part_active(s, i)) continue;
      if (s->depth_h[i] < min_depth || s->depth_h[i] > max_depth) continue;
      const float hi = s->h[i];
      const double di = di_key[a] + hi * kernel_gamma + dx_max - rshift;
      if (di < dj_min) continue;
      const float hig2 = hi * hi * kernel_gamma2;
      const float pix = s->x[3 * i + 0] - (cj->loc[0] + shift[0]);
      const float piy = s->x[3 * i + 1] - (cj->loc[1] + shift[1]);
      const float piz = s->x[3 * i + 2] - (cj->loc[2] + shift[2]);
      for (int b = 0; b < count_j; b++) {
        if (!(dj_key[b] < di)) continue;
        const long long j = pj0 + b;
        if (part_inhibited(s, j)) continue;
        const float pjx = s->x[3 * j + 0] - cj->loc[0];
        const float pjy = s->x[3 * j + 1] - cj->loc[1];
        const float pjz = s->x[3 * j + 2] - cj->loc[2];
        const float dx[3] = {pix - pjx, piy - pjy, piz - pjz};
        const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        if (r2 < hig2) interact(s, loop, r2, dx, i, j);
      }
    }
  }
  if (cell_active(s, cj)) {
    for (int b = 0; b < count_j; b++) {
      const long long j = pj0 + b;
      if (!(dj_key[b] - hj_max - dx_max < di_max)) continue; /* :1421 */
      if (!part_active(s, j)) continue;
      if (s->depth_h[j] < min_depth || s->depth_h[j] > max_depth) continue;
      const float hj = s->h[j];
      const double dj = dj_key[b] - hj * kernel_gamma - dx_max + rshift;
      if (dj - rshift > di_max) continue;
      const float hjg2 = hj * hj * kernel_gamma2;
      const float pjx = s->x[3 * j + 0] - cj->loc[0];
      const float pjy = s->x[3 * j + 1] - cj->loc[1];
      const float pjz = s->x[3 * j + 2] - cj->loc[2];
      for (int a = 0; a < count_i; a++) {
        if (!(di_key[a] > dj)) continue;
        const long long i = pi0 + a;
        if (part_inhibited(s, i)) continue;
        const float pix = s->x[3 * i + 0] - (cj->loc[0] + shift[0]);
        const float piy = s->x[3 * i + 1] - (cj->loc[1] + shift[1]);
        const float piz = s->x[3 * i + 2] - (cj->loc[2] + shift[2]);
        const float dx[3] = {pjx - pix, pjy - piy, pjz - piz};
        const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        if (r2 < hjg2) interact(s, loop, r2, dx, j, i);
      }
    }
  }
  free(di_key);
  free(dj_key);
}

/* ---------------- DOPAIR2 in gather form: functions_hydro.h:1601-2238 -------- */
static void dopair2(port_t *s, int ci_, int cj_, int sid, const double shift[3],
                    int limit_min_h, int limit_max_h) {
  const swiftgpu_cell *ci = &s->cells[ci_], *cj = &s->cells[cj_];
  double rshift = 0.0;
  for (int k = 0; k < 3; k++) rshift += shift[k] * runner_shift[sid][k];
  const int min_depth = limit_max_h ? ci->depth : 0;
  const int max_depth = limit_min_h ? ci->depth : CHAR_MAX;
  const int local_i = ci->nodeID == s->cfg.rank;
  const int local_j = cj->nodeID == s->cfg.rank;
  const int act_i = cell_active(s, ci), act_j = cell_active(s, cj);
  const double hi_max = ci->h_max;
  const double hj_max = cj->h_max;
  const int count_i = ci->count, count_j = cj->count;
  const long long pi0 = ci->first_part, pj0 = cj->first_part;
  const double dx_max = (ci->dx_max_sort + cj->dx_max_sort);
  float *di_key = (float *)malloc(sizeof(float) * count_i);
  float *dj_key = (float *)malloc(sizeof(float) * count_j);
  float kmax_i = -FLT_MAX, kmin_j = FLT_MAX;
  for (int k = 0; k < count_i; k++) {
    di_key[k] = sort_key(s, pi0 + k, sid);
    kmax_i = pmax(kmax_i, di_key[k]);
  }
  for (int k = 0; k < count_j; k++) {
    dj_key[k] = sort_key(s, pj0 + k, sid);
    kmin_j = pmin(kmin_j, dj_key[k]);
  }
  const double di_max = kmax_i;
  const double dj_min = kmin_j;
  const double shift_i[3] = {cj->loc[0] + shift[0], cj->loc[1] + shift[1],
                             cj->loc[2] + shift[2]};
  const double shift_j[3] = {cj->loc[0], cj->loc[1], cj->loc[2]};

  for (int a = 0; a < count_i; a++) {
    const long long i = pi0 + a;
    const float hi = s->h[i];
    const float hig2 = hi * hi * kernel_gamma2;
    const float pix = s->x[3 * i + 0] - shift_i[0];
    const float piy = s->x[3 * i + 1] - shift_i[1];
    const float piz = s->x[3 * i + 2] - shift_i[2];
    const int update_i = act_i && part_active(s, i) && local_i &&
                         s->depth_h[i] >= min_depth && s->depth_h[i] <= max_depth;
    /* pass A bookkeeping for i (:1740-1752) */
    const int inA_i = (di_key[a] + hi_max * kernel_gamma + dx_max - rshift > dj_min) &&
                      !part_inhibited(s, i);
    const double di = di_key[a] + hi * kernel_gamma + dx_max - rshift;
    const int okA_i = inA_i && !(di < dj_min);
    for (int b = 0; b < count_j; b++) {
      const long long j = pj0 + b;
      const float hj = s->h[j];
      const float hjg2 = hj * hj * kernel_gamma2;
      const int update_j = act_j && part_active(s, j) && local_j &&
                           s->depth_h[j] >= min_depth && s->depth_h[j] <= max_depth;
      if (!update_i && !update_j) continue;
      const float pjx = s->x[3 * j + 0] - shift_j[0];
      const float pjy = s->x[3 * j + 1] - shift_j[1];
      const float pjz = s->x[3 * j + 2] - shift_j[2];
      const float dx[3] = {pix - pjx, piy - pjy, piz - pjz};
      const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
      int hit = 0;
      /* pass A (:1737-1975): found from pi's side with r2 < hig2 */
      if (okA_i && dj_key[b] < di && !part_inhibited(s, j) && r2 < hig2) hit = 1;
      /* pass B (:1978-2230): found from pj's side with hig2 <= r2 < hjg2 */
      if (!hit) {
        const int inB_j =
            (dj_key[b] - hj_max * kernel_gamma - dx_max < di_max - rshift) &&
            !part_inhibited(s, j);
        const double dj = dj_key[b] - hj * kernel_gamma - dx_max;
        if (inB_j && !(dj > di_max - rshift) && (di_key[a] - rshift > dj) &&
            !part_inhibited(s, i) && r2 < hjg2 && r2 >= hig2)
          hit = 1;
      }
      if (!hit) continue;
      if (update_i) interact(s, LOOP_FORCE, r2, dx, i, j);
      if (update_j) {
        const float mdx[3] = {-dx[0], -dx[1], -dx[2]};
        interact(s, LOOP_FORCE, r2, mdx, j, i);
      }
    }
  }
  free(di_key);
  free(dj_key);
}

/* ---------------- cell_split_pairs (cell.c:63) derived geometrically --------
 * progeny index bits: 4 -> x, 2 -> y, 1 -> z (space_split.c:243-245). A
 * sub-pair (pid of ci, pjd of cj) exists iff the two octants touch when cj
 * sits at offset dir(sid) from ci. */
static int sub_pairs(int sid, int pid[16], int pjd[16]) {
  int d[3];
  for (int k = 0; k < 3; k++)
    d[k] = runner_shift[sid][k] > 0 ? 1 : (runner_shift[sid][k] < 0 ? -1 : 0);
  int n = 0;
  for (int a = 0; a < 8; a++)
    for (int b = 0; b < 8; b++) {
      const int ax[3] = {(a >> 2) & 1, (a >> 1) & 1, a & 1};
      const int bx[3] = {(b >> 2) & 1, (b >> 1) & 1, b & 1};
      int ok = 1;
      for (int k = 0; k < 3; k++)
        if (abs(bx[k] + 2 * d[k] - ax[k]) > 1) ok = 0;
      if (ok) {
        pid[n] = a;
        pjd[n] = b;
        n++;
      }
    }
  return n;
}
int port_sub_pairs(int sid, int *pid, int *pjd) { return sub_pairs(sid, pid, pjd); }

/* ---------------- DOSUB_PAIR1/2, DOSUB_SELF1/2: :2933,3037,3102,3203 -------- */
static int can_recurse_subpair(const port_t *s, const swiftgpu_cell *c, int loop) {
  if (loop == LOOP_FORCE) /* cell.h:966 */
    return (kernel_gamma * c->h_max + c->dx_max_part) < 0.5f * c->dmin;
  return (kernel_gamma * c->h_max_active + c->dx_max_part_old) < 0.5f * c->dmin; /* :951 */
}
static int can_recurse_subself(const port_t *s, const swiftgpu_cell *c, int loop) {
  if (loop == LOOP_FORCE) /* cell.h:1007 */
    return c->split && (kernel_gamma * c->h_max < 0.5f * c->dmin);
  return (kernel_gamma * c->h_max_active < 0.5f * c->dmin); /* :992 */
}

static void dosub_pair(port_t *s, int loop, int ci_, int cj_, int below) {
  if (!cell_active(s, &s->cells[ci_]) && !cell_active(s, &s->cells[cj_])) return;
  if (s->cells[ci_].count == 0 || s->cells[cj_].count == 0) return;
  double shift[3];
  const int sid = getsid(s, &ci_, &cj_, shift);
  const swiftgpu_cell *ci = &s->cells[ci_], *cj = &s->cells[cj_];
  if (!ci->split || ci->count < space_recurse_size_pair_hydro || !cj->split ||
      cj->count < space_recurse_size_pair_hydro) {
    if (loop == LOOP_FORCE)
      dopair2(s, ci_, cj_, sid, shift, 0, below);
    else
      dopair1(s, loop, ci_, cj_, sid, shift, 0, below);
  } else {
    if (!below && (!can_recurse_subpair(s, ci, loop) || !can_recurse_subpair(s, cj, loop)))
      below = 1;
    if (below) {
      if (loop == LOOP_FORCE)
        dopair2(s, ci_, cj_, sid, shift, 1, 1);
      else
        dopair1(s, loop, ci_, cj_, sid, shift, 1, 1);
    }
    int pid[16], pjd[16];
    const int n = sub_pairs(sid, pid, pjd);
    for (int k = 0; k < n; k++)
      if (ci->progeny[pid[k]] >= 0 && cj->progeny[pjd[k]] >= 0)
        dosub_pair(s, loop, ci->progeny[pid[k]], cj->progeny[pjd[k]], below);
  }
}

static void dosub_self(port_t *s, int loop, int c_, int below) {
  const swiftgpu_cell *c = &s->cells[c_];
  if (c->count == 0 || !cell_active(s, c)) return;
  if (!c->split || c->count < space_recurse_size_self_hydro) {
    doself(s, loop, c_, 0, below);
  } else {
    if (!below && !can_recurse_subself(s, c, loop)) below = 1;
    if (below) doself(s, loop, c_, 1, 1);
    for (int k = 0; k < 8; k++)
      if (c->progeny[k] >= 0) {
        dosub_self(s, loop, c->progeny[k], below);
        for (int j = k + 1; j < 8; j++)
          if (c->progeny[j] >= 0) dosub_pair(s, loop, c->progeny[k], c->progeny[j], below);
      }
  }
}

/* Top-level task list: one self per top cell and one pair per couple of touching
 * top cells with at least one local side (engine_maketasks.c:3501-3569). */
static int top_neighbours(const port_t *s, int a, int *out) {
  const swiftgpu_cell *ca = &s->cells[s->top[a]];
  int n = 0;
  for (int b = 0; b < s->ntop; b++) {
    if (b == a) continue;
    const swiftgpu_cell *cb = &s->cells[s->top[b]];
    int ok = 1;
    for (int k = 0; k < 3; k++) {
      double dx = fabs(cb->loc[k] - ca->loc[k]);
      if (s->cfg.periodic && dx > 0.5 * s->cfg.dim[k]) dx = s->cfg.dim[k] - dx;
      if (dx > 1.0001 * ca->width[k]) ok = 0;
    }
    if (ok) out[n++] = b;
  }
  return n;
}

static void run_loop(port_t *s, int loop) {
  int *ngb = (int *)malloc(sizeof(int) * s->ntop);
  for (int a = 0; a < s->ntop; a++) {
    if (s->cells[s->top[a]].nodeID == s->cfg.rank) dosub_self(s, loop, s->top[a], 0);
    const int n = top_neighbours(s, a, ngb);
    for (int q = 0; q < n; q++) {
      const int b = ngb[q];
      if (b < a) continue;
      if (s->cells[s->top[a]].nodeID != s->cfg.rank &&
          s->cells[s->top[b]].nodeID != s->cfg.rank)
        continue;
      dosub_pair(s, loop, s->top[a], s->top[b], 0);
    }
  }
  free(ngb);
}

/* ---------------- per-particle finalisers ---------------- */
static void init_part(port_t *s, long long p) {
  /* hydro_init_part: Minimal hydro.h:517, Gadget2 :499, SPHENIX :564 */
  s->wcount[p] = 0.f;
  s->wcount_dh[p] = 0.f;
  s->rho[p] = 0.f;
  s->rho_dh[p] = 0.f;
  s->div_v[p] = 0.f;
  s->g_div[p] = s->g_rho_dh[p] = s->g_lap[p] = 0.f;
  s->g2_lap[p] = 0.f;
  s->rot_v[3 * p] = s->rot_v[3 * p + 1] = s->rot_v[3 * p + 2] = 0.f;
#if PORT_SCHEME == SCH_SPHENIX
  s->laplace_u[p] = 0.f;
#endif
  s->nd[p] = 0;
}

static void end_density(port_t *s, long long p) {
  /* hydro_end_density: Minimal :543, Gadget2 :526, SPHENIX :613 */
  const float h = s->h[p];
  const float h_inv = 1.0f / h;
  const float h_inv_dim = h_inv * h_inv * h_inv;
  const float h_inv_dim_plus_one = h_inv_dim * h_inv;
  s->rho[p] += s->m[p] * kernel_root;
  s->rho_dh[p] -= hydro_dimension * s->m[p] * kernel_root;
  s->wcount[p] += kernel_root;
  s->wcount_dh[p] -= hydro_dimension * kernel_root;
  s->rho[p] *= h_inv_dim;
  s->rho_dh[p] *= h_inv_dim_plus_one;
  s->g_rho_dh[p] = (s->g_rho_dh[p] + hydro_dimension * s->m[p] * kernel_root) * h_inv_dim_plus_one;
  s->wcount[p] *= h_inv_dim;
  s->wcount_dh[p] *= h_inv_dim_plus_one;
  const float rho_inv = 1.f / s->rho[p];
  const float a_inv2 = 1.f / (s->step.a * s->step.a);
  s->rot_v[3 * p + 0] *= h_inv_dim_plus_one * a_inv2 * rho_inv;
  s->rot_v[3 * p + 1] *= h_inv_dim_plus_one * a_inv2 * rho_inv;
  s->rot_v[3 * p + 2] *= h_inv_dim_plus_one * a_inv2 * rho_inv;
  s->g_div[p] *= h_inv_dim_plus_one * rho_inv * a_inv2;
#if PORT_SCHEME == SCH_SPHENIX
  s->div_v[p] *= h_inv_dim_plus_one * rho_inv * a_inv2;
  s->div_v[p] += s->step.H * hydro_dimension;
#else
  s->div_v[p] *= h_inv_dim_plus_one * a_inv2 * rho_inv;
#endif
}

static void has_no_neighbours(port_t *s, long long p) {
  /* hydro_part_has_no_neighbours: Minimal :626, Gadget2 :604, SPHENIX :791 */
  const float h = s->h[p];
  const float h_inv = 1.0f / h;
  const float h_inv_dim = h_inv * h_inv * h_inv;
  s->rho[p] = s->m[p] * kernel_root * h_inv_dim;
  s->wcount[p] = kernel_root * h_inv_dim;
  s->rho_dh[p] = 0.f;
  s->wcount_dh[p] = 0.f;
  s->div_v[p] = 0.f;
  s->g_div[p] = s->g_rho_dh[p] = s->g_lap[p] = 0.f;
  s->g2_lap[p] = 0.f;
  s->rot_v[3 * p] = s->rot_v[3 * p + 1] = s->rot_v[3 * p + 2] = 0.f;
#if PORT_SCHEME == SCH_SPHENIX
  s->v_sig[p] = 0.f;
  s->laplace_u[p] = 0.f;
#endif
}

#if PORT_SCHEME == SCH_SPHENIX
/* hydro_prepare_gradient + hydro_reset_gradient: SPHENIX hydro.h:671-755 */
static void prepare_gradient(port_t *s, long long p) {
  const float fac_B = 1.f;
  const float curl_v = sqrtf(s->rot_v[3 * p] * s->rot_v[3 * p] +
                             s->rot_v[3 * p + 1] * s->rot_v[3 * p + 1] +
                             s->rot_v[3 * p + 2] * s->rot_v[3 * p + 2]);
  const float abs_div_v = fabsf(s->div_v[p]);
  const float pressure = hydro_gamma_minus_one * s->u[p] * s->rho[p];
  const float soundspeed = sqrtf(hydro_gamma * pressure / s->rho[p]);
  const float balsara =
      abs_div_v / (abs_div_v + curl_v + 0.0001f * soundspeed * fac_B / s->h[p]);
  const float common_factor = s->h[p] * hydro_dimension_inv / s->wcount[p];
  float grad_h_term;
  if (s->h[p] > 0.9999f * s->cfg.h_max) {
    grad_h_term = 0.f;
  } else {
    const float grad_W_term = common_factor * s->wcount_dh[p];
    if (grad_W_term < -0.9999f)
      grad_h_term = 0.f;
    else
      grad_h_term = common_factor * s->rho_dh[p] / (1.f + grad_W_term);
  }
  s->f[p] = grad_h_term;
  s->P[p] = pressure;
  s->cs[p] = soundspeed;
  s->balsara[p] = balsara;
  /* d balsara = d(div_v) / (|div| + |curl| + eps), d(div_v) = 1e-5 x 0.5 x (un-cancelled sum) */
  s->nz_B[p] = 5.e-6f * s->g_div[p] / (abs_div_v + curl_v + 0.0001f * soundspeed * fac_B / s->h[p]);
  /* reset_gradient */
  s->v_sig[p] = 2.f * s->cs[p];
  s->alpha_max_ngb[p] = s->alpha[p];
  s->ng[p] = 0;
}

/* hydro_end_gradient + hydro_prepare_force + hydro_reset_acceleration:
 * SPHENIX hydro.h:762-775, 840-950, 961-972 (runner_do_extra_ghost) */
static void extra_ghost_part(port_t *s, long long p, float dt_alpha) {
  const float h = s->h[p];
  const float h_inv = 1.0f / h;
  const float h_inv_dim = h_inv * h_inv * h_inv;
  const float h_inv_dim_plus_one = h_inv_dim * h_inv;
  s->laplace_u[p] *= 2.f * h_inv_dim_plus_one;
  s->g_lap[p] *= 2.f * h_inv_dim_plus_one;
  s->g2_lap[p] *= (2.f * h_inv_dim_plus_one) * (2.f * h_inv_dim_plus_one);
  /* relative noise of the diffusion alpha this laplace_u produces (it starts from alpha_diff ~ 0) */
  s->nz_ad[p] = fminf(1.f, 1.e-5f * (0.5f * s->g_lap[p] + 0.05f * sqrtf(s->g2_lap[p])) /
                               fmaxf(fabsf(s->laplace_u[p]), 1.e-30f));

  const float a = s->step.a;
  const float kernel_support_physical = s->h[p] * a * kernel_gamma;
  const float kernel_support_physical_inv = 1.f / kernel_support_physical;
  const float v_sig_physical = s->v_sig[p] * 1.f;
  const float pressure = hydro_gamma_minus_one * s->u[p] * s->rho[p];
  const float soundspeed_physical = sqrtf(hydro_gamma * pressure / s->rho[p]) * 1.f;
  const float sound_crossing_time_inverse =
      soundspeed_physical * kernel_support_physical_inv;
  const float div_v_dt =
      dt_alpha == 0.f ? 0.f : (s->div_v[p] - s->div_v_prev[p]) / dt_alpha;
  const float S = s->div_v[p] < 0.f
                      ? kernel_support_physical * kernel_support_physical *
                            pmax(0.f, -1.f * div_v_dt)
                      : 0.f;
  const float soundspeed_square = soundspeed_physical * soundspeed_physical;
  const float alpha_loc = s->cfg.viscosity_alpha_max * S / (soundspeed_square + S);
  /* absolute noise of the viscosity alpha: d alpha <= alpha_max (h gamma)^2 / c^2 d(div_v) / dt, capped */
  s->nz_al[p] = dt_alpha == 0.f ? 0.f
                                : fminf(s->cfg.viscosity_alpha_max,
                                        s->cfg.viscosity_alpha_max * kernel_support_physical * kernel_support_physical /
                                            soundspeed_square * 5.e-6f * s->g_div[p] / dt_alpha);
  if (alpha_loc > s->alpha[p]) {
    s->alpha[p] = alpha_loc;
  } else {
    const float timescale_ratio =
        dt_alpha * sound_crossing_time_inverse * s->cfg.viscosity_length;
    s->alpha[p] += alpha_loc * timescale_ratio;
    s->alpha[p] /= (1.f + timescale_ratio);
  }
  s->alpha[p] = pmax(s->alpha[p], s->cfg.viscosity_alpha_min);
  s->div_v_prev[p] = s->div_v[p];
  s->div_v_dt[p] = div_v_dt;
  const float diffusion_timescale_physical_inverse =
      v_sig_physical * kernel_support_physical_inv;
  const float sqrt_u_inv = 1.f / sqrtf(s->u[p]);
  float alpha_diff_dt = s->cfg.diffusion_beta * kernel_support_physical *
                        s->laplace_u[p] * 1.f * sqrt_u_inv * (1.f / (a * a));
  alpha_diff_dt -= (s->alpha_diff[p] - s->cfg.diffusion_alpha_min) *
                   diffusion_timescale_physical_inverse;
  float new_diffusion_alpha = s->alpha_diff[p];
  new_diffusion_alpha += alpha_diff_dt * dt_alpha;
  new_diffusion_alpha = pmax(new_diffusion_alpha, s->cfg.diffusion_alpha_min);
  const float viscous_diffusion_limit =
      s->cfg.diffusion_alpha_max *
      (1.f - s->alpha_max_ngb[p] / s->cfg.viscosity_alpha_max);
  new_diffusion_alpha = pmin(new_diffusion_alpha, viscous_diffusion_limit);
  s->alpha_diff[p] = new_diffusion_alpha;
  /* reset_acceleration + timestep_limiter_prepare_force */
  s->a[3 * p] = s->a[3 * p + 1] = s->a[3 * p + 2] = 0.f;
  s->u_dt[p] = 0.f;
  s->h_dt[p] = 0.f;
  s->g_a[p] = s->g_u[p] = s->g_hdt[p] = 0.f;
  s->g2_a[p] = s->g2_u[p] = 0.f;
  s->min_ngb[p] = num_time_bins + 1;
  s->nf[p] = 0;
}
#else
/* hydro_prepare_force + hydro_reset_acceleration: Minimal hydro.h:669-766,
 * Gadget2 hydro.h:648-744 */
static void prepare_force(port_t *s, long long p) {
  const float fac_Balsara_eps = 1.f;
  const float h_inv = 1.f / s->h[p];
  const float curl_v = sqrtf(s->rot_v[3 * p] * s->rot_v[3 * p] +
                             s->rot_v[3 * p + 1] * s->rot_v[3 * p + 1] +
                             s->rot_v[3 * p + 2] * s->rot_v[3 * p + 2]);
  const float div_physical_v = s->div_v[p] + hydro_dimension * s->step.H;
  const float abs_div_physical_v = fabsf(div_physical_v);
#if PORT_SCHEME == SCH_MINIMAL
  const float pressure = hydro_gamma_minus_one * s->u[p] * s->rho[p];
  const float soundspeed = sqrtf(hydro_gamma * pressure / s->rho[p]);
  const float common_factor = s->h[p] * hydro_dimension_inv / s->wcount[p];
  float grad_h_term;
  if (s->h[p] > 0.9999f * s->cfg.h_max) {
    grad_h_term = 0.f;
  } else {
    const float grad_W_term = common_factor * s->wcount_dh[p];
    if (grad_W_term < -0.9999f)
      grad_h_term = 0.f;
    else
      grad_h_term = common_factor * s->rho_dh[p] / (1.f + grad_W_term);
  }
  const float balsara = s->cfg.viscosity_alpha * abs_div_physical_v /
                        (abs_div_physical_v + curl_v +
                         0.0001f * fac_Balsara_eps * soundspeed * h_inv);
  s->f[p] = grad_h_term;
  s->P[p] = pressure;
  s->cs[p] = soundspeed;
  s->balsara[p] = balsara;
  s->nz_B[p] = s->cfg.viscosity_alpha * 5.e-6f * s->g_div[p] /
               (abs_div_physical_v + curl_v + 0.0001f * fac_Balsara_eps * soundspeed * h_inv);
#else /* Gadget2 */
  const float rho_inv = 1.f / s->rho[p];
  /* gas_pressure_from_entropy = entropy * pow_gamma(rho); pow_gamma = cbrt^2*x */
  const float cbrt_rho = cbrtf(s->rho[p]);
  const float comoving_pressure = s->u[p] * (cbrt_rho * cbrt_rho * s->rho[p]);
  const float soundspeed = sqrtf(hydro_gamma * comoving_pressure / s->rho[p]);
  const float P_over_rho2 = comoving_pressure * rho_inv * rho_inv;
  const float balsara = s->cfg.viscosity_alpha * abs_div_physical_v /
                        (abs_div_physical_v + curl_v +
                         0.0001f * fac_Balsara_eps * soundspeed * h_inv);
  float rho_dh = s->rho_dh[p];
  if (s->h[p] > 0.9999f * s->cfg.h_max) rho_dh = 0.f;
  const float grad_rho_term = hydro_dimension_inv * s->h[p] * rho_dh * rho_inv;
  float omega_inv;
  if (grad_rho_term < -0.9999f)
    omega_inv = 1.f;
  else
    omega_inv = 1.f / (1.f + grad_rho_term);
  s->f[p] = omega_inv;
  s->P[p] = P_over_rho2;
  s->cs[p] = soundspeed;
  s->balsara[p] = balsara;
  s->nz_B[p] = s->cfg.viscosity_alpha * 5.e-6f * s->g_div[p] /
               (abs_div_physical_v + curl_v + 0.0001f * fac_Balsara_eps * soundspeed * h_inv);
#endif
  s->min_ngb[p] = num_time_bins + 1; /* timestep_limiter_prepare_force */
  s->a[3 * p] = s->a[3 * p + 1] = s->a[3 * p + 2] = 0.f;
  s->u_dt[p] = 0.f;
  s->h_dt[p] = 0.f;
  s->g_a[p] = s->g_u[p] = s->g_hdt[p] = 0.f;
  s->g2_a[p] = s->g2_u[p] = 0.f;
  s->v_sig[p] = 2.f * s->cs[p];
  s->nf[p] = 0;
}
#endif

/* cell_set_part_h_depth: cell.h:1787-1815 */
static void set_part_h_depth(port_t *s, long long p, int leaf) {
  const float h = s->h[p];
  const swiftgpu_cell *c = &s->cells[leaf];
  if (h < c->h_min_allowed) {
    s->depth_h[p] = (signed char)c->depth;
    return;
  }
  int ci = leaf;
  while (ci >= 0) {
    c = &s->cells[ci];
    if (h >= c->h_min_allowed && h < c->h_max_allowed) {
      s->depth_h[p] = (signed char)c->depth;
      return;
    }
    ci = c->parent;
  }
}

/* ---------------- subset loops for the ghost's redo ---------------- */
/* DOSELF_SUBSET: :1108-1200 */
static void doself_subset(port_t *s, int c_, const long long *ind, int count) {
  const swiftgpu_cell *c = &s->cells[c_];
  for (int k = 0; k < count; k++) {
    const long long i = ind[k];
    const float pix[3] = {(float)(s->x[3 * i] - c->loc[0]),
                          (float)(s->x[3 * i + 1] - c->loc[1]),
                          (float)(s->x[3 * i + 2] - c->loc[2])};
    const float hi = s->h[i];
    const float hig2 = hi * hi * kernel_gamma2;
    for (int b = 0; b < c->count; b++) {
      const long long j = c->first_part + b;
      if (j == i || part_inhibited(s, j)) continue;
      const float pjx[3] = {(float)(s->x[3 * j] - c->loc[0]),
                            (float)(s->x[3 * j + 1] - c->loc[1]),
                            (float)(s->x[3 * j + 2] - c->loc[2])};
      const float dx[3] = {pix[0] - pjx[0], pix[1] - pjx[1], pix[2] - pjx[2]};
      const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
      if (r2 < hig2) iact_density(s, r2, dx, hi, i, j);
    }
  }
}

/* DOPAIR_SUBSET_BRANCH + DOPAIR_SUBSET: :1036-1100, :855-1030 (cells sorted) */
static void dopair_subset(port_t *s, int ci_, const long long *ind, int count, int cj_) {
  const swiftgpu_cell *ci = &s->cells[ci_], *cj = &s->cells[cj_];
  if (cj->count == 0) return;
  double shift[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < 3; k++) {
    if (cj->loc[k] - ci->loc[k] < -s->cfg.dim[k] / 2)
      shift[k] = s->cfg.dim[k];
    else if (cj->loc[k] - ci->loc[k] > s->cfg.dim[k] / 2)
      shift[k] = -s->cfg.dim[k];
  }
  int sid = 0;
  for (int k = 0; k < 3; k++)
    sid = 3 * sid + ((cj->loc[k] - ci->loc[k] + shift[k] < 0)   ? 0
                     : (cj->loc[k] - ci->loc[k] + shift[k] > 0) ? 2
                                                                : 1);
  const int flipped = sid < 13; /* runner_flip */
  sid = sortlistID[sid];
  const float dxj = cj->dx_max_sort;
  for (int k = 0; k < count; k++) {
    const long long i = ind[k];
    const double pix = s->x[3 * i + 0] - (shift[0]);
    const double piy = s->x[3 * i + 1] - (shift[1]);
    const double piz = s->x[3 * i + 2] - (shift[2]);
    const float hi = s->h[i];
    const float hig2 = hi * hi * kernel_gamma2;
    double di;
    if (!flipped)
      di = hi * kernel_gamma + dxj + pix * runner_shift[sid][0] +
           piy * runner_shift[sid][1] + piz * runner_shift[sid][2];
    else
      di = -hi * kernel_gamma - dxj + pix * runner_shift[sid][0] +
           piy * runner_shift[sid][1] + piz * runner_shift[sid][2];
    for (int b = 0; b < cj->count; b++) {
      const long long j = cj->first_part + b;
      const float dj = sort_key(s, j, sid);
      if (!flipped) {
        if (!(dj < di)) continue;
      } else {
        if (!(di < dj)) continue;
      }
      if (part_inhibited(s, j)) continue;
      const float dx[3] = {(float)(pix - s->x[3 * j]), (float)(piy - s->x[3 * j + 1]),
                           (float)(piz - s->x[3 * j + 2])};
      const float r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
      if (r2 < hig2) iact_density(s, r2, dx, hi, i, j);
    }
  }
}

static int cell_contains(const swiftgpu_cell *c, long long p) {
  return p >= c->first_part && p < c->first_part + c->count;
}
static int can_recurse_pair_task(const swiftgpu_cell *c) { /* cell.h:933 */
  return c->split && ((kernel_gamma * c->h_max_old + c->dx_max_part_old) < 0.5f * c->dmin);
}
static int can_recurse_self_task(const swiftgpu_cell *c) { /* cell.h:978 */
  return c->split && (kernel_gamma * c->h_max_old < 0.5f * c->dmin);
}
static int find_sub(const port_t *s, const swiftgpu_cell *c, long long p) {
  for (int k = 0; k < 8; k++)
    if (c->progeny[k] >= 0 && cell_contains(&s->cells[c->progeny[k]], p))
      return c->progeny[k];
  return -1;
}

/* DOSUB_PAIR_SUBSET: :3290-3339 */
static void dosub_pair_subset(port_t *s, int ci_, const long long *ind, int count, int cj_) {
  const swiftgpu_cell *ci = &s->cells[ci_], *cj = &s->cells[cj_];
  if (ci->count == 0 || cj->count == 0) return;
  if (!cell_active(s, ci)) return;
  if (can_recurse_pair_task(ci) && can_recurse_pair_task(cj)) {
    const int sub = find_sub(s, ci, ind[0]);
    int a = ci_, b = cj_;
    double shift[3];
    const int sid = getsid(s, &a, &b, shift);
    const swiftgpu_cell *ca = &s->cells[a], *cb = &s->cells[b];
    int pid[16], pjd[16];
    const int n = sub_pairs(sid, pid, pjd);
    for (int k = 0; k < n; k++) {
      if (ca->progeny[pid[k]] == sub && cb->progeny[pjd[k]] >= 0)
        dosub_pair_subset(s, ca->progeny[pid[k]], ind, count, cb->progeny[pjd[k]]);
      if (ca->progeny[pid[k]] >= 0 && cb->progeny[pjd[k]] == sub)
        dosub_pair_subset(s, cb->progeny[pjd[k]], ind, count, ca->progeny[pid[k]]);
    }
  } else if (cell_active(s, ci)) {
    dopair_subset(s, ci_, ind, count, cj_);
  }
}

/* DOSUB_SELF_SUBSET: :3341-3367 */
static void dosub_self_subset(port_t *s, int ci_, const long long *ind, int count) {
  const swiftgpu_cell *ci = &s->cells[ci_];
  if (ci->count == 0) return;
  if (!cell_active(s, ci)) return;
  if (ci->split && can_recurse_self_task(ci)) {
    const int sub = find_sub(s, ci, ind[0]);
    dosub_self_subset(s, sub, ind, count);
    for (int j = 0; j < 8; j++)
      if (ci->progeny[j] != sub && ci->progeny[j] >= 0)
        dosub_pair_subset(s, sub, ind, count, ci->progeny[j]);
  } else
    doself_subset(s, ci_, ind, count);
}

/* ---------------- runner_do_ghost: runner_ghost.c:1113-1635 ---------------- */
static void ghost_leaf(port_t *s, int c_, int *ngb_top) {
  swiftgpu_cell *c = &s->cells[c_];
  const float hydro_h_max = s->cfg.h_max;
  const float hydro_h_min = s->cfg.h_min;
  const float eps = s->cfg.h_tolerance;
  const float eta = s->cfg.eta_neighbours;
  const float hydro_eta_dim = eta * eta * eta;
  const int max_smoothing_iter = s->cfg.max_smoothing_iterations;
  float h_max = c->h_max;
  float h_max_active = 0.f;
  int count = 0, redo = 0;
  long long *pid = (long long *)malloc(sizeof(long long) * c->count);
  float *h_0 = (float *)malloc(sizeof(float) * c->count);
  float *left = (float *)malloc(sizeof(float) * c->count);
  float *right = (float *)malloc(sizeof(float) * c->count);
  for (int k = 0; k < c->count; k++)
    if (part_active(s, c->first_part + k)) {
      pid[count] = c->first_part + k;
      h_0[count] = s->h[c->first_part + k];
      left[count] = 0.f;
      right[count] = hydro_h_max;
      ++count;
    }
  int num_reruns;
  for (num_reruns = 0; count > 0 && num_reruns < max_smoothing_iter; num_reruns++) {
    redo = 0;
    for (int i = 0; i < count; i++) {
      const long long p = pid[i];
      const float h_old = s->h[p];
      const float h_old_dim = h_old * h_old * h_old;
      const float h_old_dim_minus_one = h_old * h_old;
      float h_new;
      int has_no_ngb = 0;
      if (s->wcount[p] < 1.e-5 * kernel_root) {
        has_no_ngb = 1;
        h_new = 2.f * h_old;
      } else {
        end_density(s, p);
        if (s->cfg.use_mass_weighted_num_ngb) {
          const float inv_mass = 1.f / s->m[p];
          s->wcount[p] = s->rho[p] * inv_mass;
          s->wcount_dh[p] = s->rho_dh[p] * inv_mass;
        }
        const float n_sum = s->wcount[p] * h_old_dim;
        const float n_target = hydro_eta_dim;
        const float f = n_sum - n_target;
        const float f_prime = s->wcount_dh[p] * h_old_dim +
                              hydro_dimension * s->wcount[p] * h_old_dim_minus_one;
        if (n_sum < n_target)
          left[i] = pmax(left[i], h_old);
        else if (n_sum > n_target)
          right[i] = pmin(right[i], h_old);
        if (((s->h[p] >= hydro_h_max) && (f < 0.f)) ||
            ((s->h[p] <= hydro_h_min) && (f > 0.f))) {
#if PORT_SCHEME == SCH_SPHENIX
          prepare_gradient(s, p);
#else
          prepare_force(s, p);
#endif
          h_max = pmax(h_max, s->h[p]);
          h_max_active = pmax(h_max_active, s->h[p]);
          continue;
        }
        h_new = h_old - f / (f_prime + FLT_MIN);
        h_new = pmin(h_new, 2.f * h_old);
        h_new = pmax(h_new, 0.5f * h_old);
        h_new = pmax(h_new, left[i]);
        h_new = pmin(h_new, right[i]);
      }
      if (fabsf(h_new - h_old) > eps * h_old) {
        if ((h_new == left[i] && h_old == right[i]) ||
            (h_old == left[i] && h_new == right[i])) {
          const float l = left[i], r = right[i];
          s->h[p] = cbrtf(0.5f * (l * l * l + r * r * r));
        } else {
          s->h[p] = h_new;
        }
        if (s->h[p] < hydro_h_max && s->h[p] > hydro_h_min) {
          pid[redo] = pid[i];
          h_0[redo] = h_0[i];
          left[redo] = left[i];
          right[redo] = right[i];
          redo += 1;
          init_part(s, p);
          continue;
        } else if (s->h[p] <= hydro_h_min) {
          s->h[p] = hydro_h_min;
        } else if (s->h[p] >= hydro_h_max) {
          s->h[p] = hydro_h_max;
          if (has_no_ngb) has_no_neighbours(s, p);
        }
      }
      set_part_h_depth(s, p, c_);
      h_max = pmax(h_max, s->h[p]);
      h_max_active = pmax(h_max_active, s->h[p]);
#if PORT_SCHEME == SCH_SPHENIX
      prepare_gradient(s, p);
#else
      prepare_force(s, p);
#endif
    }
    count = redo;
    if (count > 0) {
      s->ghost_redo[num_reruns < 31 ? num_reruns : 31] += count;
      /* climb to the top level, where the density tasks are linked (:1548) */
      int finger = c_;
      while (s->cells[finger].parent >= 0) finger = s->cells[finger].parent;
      dosub_self_subset(s, finger, pid, count);
      int a = -1;
      for (int q = 0; q < s->ntop; q++)
        if (s->top[q] == finger) a = q;
      const int n = top_neighbours(s, a, ngb_top);
      for (int q = 0; q < n; q++)
        dosub_pair_subset(s, finger, pid, count, s->top[ngb_top[q]]);
    }
    if (num_reruns + 1 > s->ghost_iterations) s->ghost_iterations = num_reruns + 1;
  }
  if (count) s->ghost_failed += count;
  free(pid);
  free(h_0);
  free(left);
  free(right);
  /* atomic_max_f on this cell and all parents (:1621-1632) */
  for (int f = c_; f >= 0; f = s->cells[f].parent) {
    s->cells[f].h_max = pmax(s->cells[f].h_max, h_max);
    s->cells[f].h_max_active = pmax(s->cells[f].h_max_active, h_max_active);
  }
}

static void ghost_recurse(port_t *s, int c_, int *ngb_top) {
  const swiftgpu_cell *c = &s->cells[c_];
  if (c->count == 0 || !cell_active(s, c)) return;
  if (c->split) {
    for (int k = 0; k < 8; k++)
      if (c->progeny[k] >= 0) ghost_recurse(s, c->progeny[k], ngb_top);
  } else
    ghost_leaf(s, c_, ngb_top);
}

/* ---------------- AoS <-> SoA ---------------- */
#define RD(T, off, p) (*(const T *)(base + (size_t)L->size * (p) + (off)))
#define WR(T, off, p) (*(T *)(base + (size_t)L->size * (p) + (off)))

static float *falloc(long long n) { return (float *)calloc((size_t)n, sizeof(float)); }

port_t *port_create(const swiftgpu_config *cfg, const swiftgpu_step *step,
                    const swiftgpu_cell *cells, int ncells, const int *top,
                    int ntop, const void *parts_aos, long long n) {
  if (cfg->scheme != PORT_SCHEME) return NULL;
  port_t *s = (port_t *)calloc(1, sizeof(port_t));
  s->cfg = *cfg;
  s->step = *step;
  s->ncells = ncells;
  s->cells = (swiftgpu_cell *)malloc(sizeof(swiftgpu_cell) * ncells);
  memcpy(s->cells, cells, sizeof(swiftgpu_cell) * ncells);
  int *t = (int *)malloc(sizeof(int) * ntop);
  memcpy(t, top, sizeof(int) * ntop);
  s->top = t;
  s->ntop = ntop;
  s->n = n;
  s->x = (double *)calloc((size_t)3 * n, sizeof(double));
  s->v = falloc(3 * n); s->a = falloc(3 * n); s->rot_v = falloc(3 * n);
  s->m = falloc(n); s->h = falloc(n); s->u = falloc(n); s->u_dt = falloc(n);
  s->rho = falloc(n); s->wcount = falloc(n); s->wcount_dh = falloc(n);
  s->rho_dh = falloc(n); s->div_v = falloc(n); s->f = falloc(n); s->P = falloc(n);
  s->cs = falloc(n); s->balsara = falloc(n); s->v_sig = falloc(n); s->h_dt = falloc(n);
  s->alpha = falloc(n); s->alpha_diff = falloc(n); s->div_v_prev = falloc(n);
  s->div_v_dt = falloc(n); s->laplace_u = falloc(n); s->alpha_max_ngb = falloc(n);
  s->g_a = falloc(n); s->g_u = falloc(n); s->g_hdt = falloc(n); s->g_div = falloc(n);
  s->g_rho_dh = falloc(n); s->g_lap = falloc(n);
  s->g2_a = falloc(n); s->g2_u = falloc(n); s->g2_lap = falloc(n);
  s->nz_B = falloc(n); s->nz_ad = falloc(n); s->nz_al = falloc(n);
  s->time_bin = (signed char *)calloc(n, 1);
  s->depth_h = (signed char *)calloc(n, 1);
  s->min_ngb = (signed char *)calloc(n, 1);
  s->nd = (int *)calloc(n, sizeof(int));
  s->ng = (int *)calloc(n, sizeof(int));
  s->nf = (int *)calloc(n, sizeof(int));
  const swiftgpu_part_layout *L = &s->cfg.layout;
  const char *base = (const char *)parts_aos;
  for (long long p = 0; p < n; p++) {
    for (int k = 0; k < 3; k++) {
      s->x[3 * p + k] = RD(double, L->x + 8 * k, p);
      s->v[3 * p + k] = RD(float, L->v + 4 * k, p);
    }
    s->m[p] = RD(float, L->mass, p);
    s->h[p] = RD(float, L->h, p);
    s->u[p] = RD(float, PORT_SCHEME == SCH_GADGET2 ? L->entropy : L->u, p);
    s->rho[p] = RD(float, L->rho, p);
    s->time_bin[p] = RD(signed char, L->time_bin, p);
    s->depth_h[p] = RD(signed char, L->depth_h, p);
    /* The density/force union holds the force members of the last step the
     * particle was active in: that is what an INACTIVE neighbour contributes
     * to the force loop (functions_hydro.h:1757-1849). */
    s->f[p] = RD(float, L->f, p);
    s->P[p] = RD(float, PORT_SCHEME == SCH_GADGET2 ? L->P_over_rho2 : L->pressure, p);
    s->cs[p] = RD(float, L->soundspeed, p);
    s->balsara[p] = RD(float, L->balsara, p);
    s->v_sig[p] = RD(float, L->v_sig, p);
    s->h_dt[p] = RD(float, L->h_dt, p);
    for (int k = 0; k < 3; k++) s->a[3 * p + k] = RD(float, L->a_hydro + 4 * k, p);
    s->u_dt[p] = RD(float, PORT_SCHEME == SCH_GADGET2 ? L->entropy_dt : L->u_dt, p);
    s->min_ngb[p] = RD(signed char, L->min_ngb_time_bin, p);
#if PORT_SCHEME == SCH_SPHENIX
    s->alpha[p] = RD(float, L->visc_alpha, p);
    s->alpha_diff[p] = RD(float, L->diff_alpha, p);
    s->div_v_prev[p] = RD(float, L->div_v_previous_step, p);
    s->div_v[p] = RD(float, L->div_v, p);
    s->div_v_dt[p] = RD(float, L->div_v_dt, p);
    s->laplace_u[p] = RD(float, L->laplace_u, p);
    s->alpha_max_ngb[p] = RD(float, L->alpha_visc_max_ngb, p);
#endif
  }
  return s;
}

void port_destroy(port_t *s) {
  if (!s) return;
  free(s->cells); free((void *)s->top); free(s->x); free(s->v); free(s->a);
  free(s->rot_v); free(s->m); free(s->h); free(s->u); free(s->u_dt); free(s->rho);
  free(s->wcount); free(s->wcount_dh); free(s->rho_dh); free(s->div_v); free(s->f);
  free(s->P); free(s->cs); free(s->balsara); free(s->v_sig); free(s->h_dt);
  free(s->g2_a); free(s->g2_u); free(s->g2_lap); free(s->nz_B); free(s->nz_ad); free(s->nz_al);
  free(s->g_a); free(s->g_u); free(s->g_hdt); free(s->g_div); free(s->g_rho_dh); free(s->g_lap);
  free(s->alpha); free(s->alpha_diff); free(s->div_v_prev); free(s->div_v_dt);
  free(s->laplace_u); free(s->alpha_max_ngb); free(s->time_bin); free(s->depth_h);
  free(s->min_ngb); free(s->nd); free(s->ng); free(s->nf);
  free(s);
}

int port_run(port_t *s, unsigned mask) {
  const long long n = s->n;
  if (mask & SWIFTGPU_PHASE_DENSITY) {
    for (long long p = 0; p < n; p++)
      if (part_active(s, p)) init_part(s, p);
    run_loop(s, LOOP_DENSITY);
  }
  if (mask & SWIFTGPU_PHASE_GHOST) {
    int *ngb = (int *)malloc(sizeof(int) * s->ntop);
    s->ghost_iterations = 0;
    memset(s->ghost_redo, 0, sizeof(s->ghost_redo));
    s->ghost_failed = 0;
    for (int a = 0; a < s->ntop; a++)
      if (s->cells[s->top[a]].nodeID == s->cfg.rank) ghost_recurse(s, s->top[a], ngb);
    free(ngb);
  }
#if PORT_SCHEME == SCH_SPHENIX
  if (mask & SWIFTGPU_PHASE_GRADIENT) run_loop(s, LOOP_GRADIENT);
  if (mask & SWIFTGPU_PHASE_EXTRA_GHOST) {
    for (int ci = 0; ci < s->ncells; ci++) {
      const swiftgpu_cell *c = &s->cells[ci];
      if (c->split || c->nodeID != s->cfg.rank || !cell_active(s, c)) continue;
      for (int k = 0; k < c->count; k++) {
        const long long p = c->first_part + k;
        if (!part_active(s, p)) continue;
        /* get_timestep (timeline.h:91) -> double -> float argument */
        const int bin = s->time_bin[p];
        const double dt = (bin <= 0 ? 0LL : 1LL << (bin + 1)) * s->step.time_base;
        extra_ghost_part(s, p, (float)dt);
      }
    }
  }
#endif
  if (mask & SWIFTGPU_PHASE_FORCE) run_loop(s, LOOP_FORCE);
  if (mask & SWIFTGPU_PHASE_END_FORCE) {
    /* hydro_end_force: Minimal :878, Gadget2 :868, SPHENIX :1097 */
    for (int ci = 0; ci < s->ncells; ci++) {
      const swiftgpu_cell *c = &s->cells[ci];
      if (c->split || c->nodeID != s->cfg.rank || !cell_active(s, c)) continue;
      for (int k = 0; k < c->count; k++) {
        const long long p = c->first_part + k;
        if (!part_active(s, p)) continue;
        s->h_dt[p] *= s->h[p] * hydro_dimension_inv;
        s->g_hdt[p] *= s->h[p] * hydro_dimension_inv;
#if PORT_SCHEME == SCH_GADGET2
        /* 0.5 * gas_entropy_from_internal_energy(rho, entropy_dt) */
        const float cbrt_inv = 1.f / cbrtf(s->rho[p]);
        const float pow_mgm1 = cbrt_inv * cbrt_inv; /* rho^-(gamma-1) */
        s->u_dt[p] = 0.5f * (hydro_gamma_minus_one * s->u_dt[p] * pow_mgm1);
        s->g_u[p] = 0.5f * (hydro_gamma_minus_one * s->g_u[p] * pow_mgm1);
        s->g2_u[p] *= (0.5f * hydro_gamma_minus_one * pow_mgm1) * (0.5f * hydro_gamma_minus_one * pow_mgm1);
#endif
      }
    }
  }
  return s->ghost_failed;
}

/* Writes back the fields valid after the phases of `mask` were run. */
void port_get_parts(const port_t *s, unsigned mask, void *parts_aos) {
  const swiftgpu_part_layout *L = &s->cfg.layout;
  char *base = (char *)parts_aos;
  /* Once the ghost ran, the density/force union holds the force members. */
  const int force_valid = (mask & ~(unsigned)(SWIFTGPU_PHASE_SORT | SWIFTGPU_PHASE_DENSITY)) != 0;
  for (long long p = 0; p < s->n; p++) {
    if (!part_active(s, p)) continue; /* inactive particles are read-only */
    WR(float, L->h, p) = s->h[p];
    WR(float, L->rho, p) = s->rho[p];
    WR(signed char, L->depth_h, p) = s->depth_h[p];
    if (!force_valid) {
      WR(float, L->wcount, p) = s->wcount[p];
      WR(float, L->wcount_dh, p) = s->wcount_dh[p];
      WR(float, L->rho_dh, p) = s->rho_dh[p];
      for (int k = 0; k < 3; k++) WR(float, L->rot_v + 4 * k, p) = s->rot_v[3 * p + k];
      WR(float, L->div_v, p) = s->div_v[p];
    } else {
#if PORT_SCHEME == SCH_SPHENIX
      WR(float, L->div_v, p) = s->div_v[p];
      WR(float, L->v_sig, p) = s->v_sig[p];
      WR(float, L->laplace_u, p) = s->laplace_u[p];
      WR(float, L->visc_alpha, p) = s->alpha[p];
      WR(float, L->diff_alpha, p) = s->alpha_diff[p];
      WR(float, L->div_v_previous_step, p) = s->div_v_prev[p];
      WR(float, L->div_v_dt, p) = s->div_v_dt[p];
      WR(float, L->alpha_visc_max_ngb, p) = s->alpha_max_ngb[p];
      WR(float, L->pressure, p) = s->P[p];
#elif PORT_SCHEME == SCH_MINIMAL
      WR(float, L->pressure, p) = s->P[p];
      WR(float, L->v_sig, p) = s->v_sig[p];
#else
      WR(float, L->P_over_rho2, p) = s->P[p];
      WR(float, L->v_sig, p) = s->v_sig[p];
#endif
      WR(float, L->f, p) = s->f[p];
      WR(float, L->soundspeed, p) = s->cs[p];
      WR(float, L->balsara, p) = s->balsara[p];
      WR(float, L->h_dt, p) = s->h_dt[p];
      for (int k = 0; k < 3; k++) WR(float, L->a_hydro + 4 * k, p) = s->a[3 * p + k];
      WR(float, PORT_SCHEME == SCH_GADGET2 ? L->entropy_dt : L->u_dt, p) = s->u_dt[p];
      WR(signed char, L->min_ngb_time_bin, p) = s->min_ngb[p];
    }
  }
}

void port_get_cells(const port_t *s, swiftgpu_cell *cells) {
  for (int i = 0; i < s->ncells; i++) {
    cells[i].h_max = s->cells[i].h_max;
    cells[i].h_max_active = s->cells[i].h_max_active;
  }
}
void port_get_counts(const port_t *s, int *nd, int *ng, int *nf) {
  if (nd) memcpy(nd, s->nd, sizeof(int) * s->n);
  if (ng) memcpy(ng, s->ng, sizeof(int) * s->n);
  if (nf) memcpy(nf, s->nf, sizeof(int) * s->n);
}
int port_ghost_iterations(const port_t *s) { return s->ghost_iterations; }
/* particles that entered re-run k of the ghost (k = 0..31), summed over the leaves */
void port_ghost_redo(const port_t *s, long long out[32]) { memcpy(out, s->ghost_redo, sizeof(s->ghost_redo)); }
/* The un-cancelled sums (sum over neighbours of |pair term|, finalised with the
 * same factors as the sums) of a_hydro (norm), u_dt | entropy_dt, h_dt, div_v,
 * rho_dh and laplace_u: what the parity metric floors its relative errors with. */
void port_get_gross(const port_t *s, float *a, float *u, float *hdt, float *div, float *rho_dh,
                    float *lap, float *a2, float *u2, float *lap2) {
  const size_t b = sizeof(float) * (size_t)s->n;
  if (a) memcpy(a, s->g_a, b);
  if (u) memcpy(u, s->g_u, b);
  if (hdt) memcpy(hdt, s->g_hdt, b);
  if (div) memcpy(div, s->g_div, b);
  if (rho_dh) memcpy(rho_dh, s->g_rho_dh, b);
  if (lap) memcpy(lap, s->g_lap, b);
  if (a2) memcpy(a2, s->g2_a, b);
  if (u2) memcpy(u2, s->g2_u, b);
  if (lap2) memcpy(lap2, s->g2_lap, b);
}
