/*
 * sph_math.cuh - device-side SPH arithmetic of the three schemes.
 *
 * Two kinds of arithmetic live here and they are kept apart on purpose:
 *
 *  (1) EXACT path - everything that decides WHETHER two particles interact
 *      (frame positions, r2, h^2 gamma^2, sort keys, pruning thresholds). The
 *      reference evaluates these with separate IEEE multiplies and adds (its
 *      oracle build has no FMA), so they are written with the __f*_rn /
 *      __d*_rn intrinsics, which nvcc never contracts. Neighbour sets are
 *      therefore bit-identical to the reference's.
 *
 *  (2) FAST path - the interaction bodies (runner_iact_nonsym_*), free to use
 *      FMA: results differ from the reference by summation order and last-bit
 *      rounding only (tolerance 1e-5 relative, tests/).
 *
 * Expression order follows the reference files cited at each function.
 */
#ifndef SWIFTGPU_SPH_MATH_CUH
#define SWIFTGPU_SPH_MATH_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace swiftgpu {

enum { SCH_MINIMAL = 0, SCH_GADGET2 = 1, SCH_SPHENIX = 2 };
enum { LOOP_DENSITY = 0, LOOP_GRADIENT = 1, LOOP_FORCE = 2, LOOP_LIMITER = 3 };

/* kernel_hydro.h:41-55,205-237 (cubic spline, 3D); values pinned against the
 * reference build in tests/golden/reference_constants.json. */
#define KERNEL_GAMMA 0x1.d363d4p+0f
#define KERNEL_GAMMA2 0x1.aaaab0p+1f
#define KERNEL_ROOT 0x1.ac78aep-2f
#define KERNEL_CONSTANT 0x1.45f306p+2f
#define KERNEL_GAMMA_INV_DIM 0x1.50854ap-3f
#define KERNEL_GAMMA_INV_DIM_PLUS_ONE 0x1.70a3d0p-4f
#define KERNEL_GAMMA_INV ((float)(1. / (double)KERNEL_GAMMA))
#define HYDRO_DIMENSION 3.f
#define HYDRO_DIMENSION_INV 0.3333333333f
#define HYDRO_GAMMA 1.66666666666666667f
#define HYDRO_GAMMA_MINUS_ONE 0.66666666666666667f
#define CONST_VISCOSITY_BETA 3.0f
#define NUM_TIME_BINS 56

__constant__ double c_runner_shift[13][3] = {
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

/* ------------------------- EXACT path helpers ------------------------- */

/* r2 = dx[0]*dx[0] + dx[1]*dx[1] + dx[2]*dx[2], functions_hydro.h:1347 */
__device__ __forceinline__ float r2_exact(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
/* hig2 = hi * hi * kernel_gamma2, functions_hydro.h:1326 */
__device__ __forceinline__ float hg2_exact(float h) {
  return __fmul_rn(__fmul_rn(h, h), KERNEL_GAMMA2);
}
/* sort key of runner_do_hydro_sort, runner_sort.c:411-413 */
__device__ __forceinline__ float sort_key(double x, double y, double z, int sid) {
  const double d = __dadd_rn(__dadd_rn(__dmul_rn(x, c_runner_shift[sid][0]),
                                       __dmul_rn(y, c_runner_shift[sid][1])),
                             __dmul_rn(z, c_runner_shift[sid][2]));
  return __double2float_rn(d);
}
/* (float)(a - b) on doubles */
__device__ __forceinline__ float dsubf(double a, double b) {
  return __double2float_rn(__dsub_rn(a, b));
}

/* ------------------------- kernel_deval ------------------------- */
/* kernel_hydro.h:257-285 with the cubic-spline coefficient rows of :61-66.
 * The row select is done arithmetically (no table load). */
__device__ __forceinline__ void kernel_deval(float u, float &W, float &dW_dx) {
  const float x = u * KERNEL_GAMMA_INV;
  const int temp = (int)(x * 2.f);
  const int ind = temp > 2 ? 2 : temp;
  /* rows: {3,-3,0,0.5}, {-1,3,-3,1}, {0,0,0,0} */
  const float c0 = ind == 0 ? 3.f : (ind == 1 ? -1.f : 0.f);
  const float c1 = ind == 0 ? -3.f : (ind == 1 ? 3.f : 0.f);
  const float c2 = ind == 1 ? -3.f : 0.f;
  const float c3 = ind == 0 ? 0.5f : (ind == 1 ? 1.f : 0.f);
  float w = c0 * x + c1;
  float dw_dx = c0;
  dw_dx = dw_dx * x + w;
  w = x * w + c2;
  dw_dx = dw_dx * x + w;
  w = x * w + c3;
  w = fmaxf(w, 0.f);
  dw_dx = fminf(dw_dx, 0.f);
  W = w * KERNEL_CONSTANT * KERNEL_GAMMA_INV_DIM;
  dW_dx = dw_dx * KERNEL_CONSTANT * KERNEL_GAMMA_INV_DIM_PLUS_ONE;
}

/* ------------------------- interaction bodies ------------------------- */

/* r = sqrt(r2) and 1/r without the out-of-line IEEE sqrt/divide subroutines
 * nvcc emits for sqrtf() and '/': MUFU.RSQ + one Newton step (|error| < 1 ulp)
 * and MUFU.RCP + one Newton step. Quotients a / b of the
 * reference become a * rcp_rn(b) (<= 1 ulp apart, the same order as the FMA
 * contractions of this path). r2 == 0 (coincident particles) gives r_inv = 0
 * like the reference's `r ? 1/r : 0`. */
__device__ __forceinline__ void sqrt_and_inverse(float r2, float &r, float &r_inv) {
  const float y = rsqrtf(r2);
  const float r0 = r2 * y;
  const float e = fmaf(-r0, r0, r2);
  const float rr = fmaf(0.5f * y, e, r0);
  r = r2 > 0.f ? rr : 0.f;
  /* 1/r from the refined root: y is 1/sqrt(r2) to 2 ulp, one Newton step on it */
  const float yi = fmaf(y, fmaf(-rr, y, 1.f), y);
  r_inv = r2 > 0.f ? yi : 0.f;
}
/* 1/x: MUFU.RCP + one Newton step (|error| < 1 ulp; __frcp_rn costs ~10 instructions) */
__device__ __forceinline__ float rcp_rn(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}
__device__ __forceinline__ float sqrt_newton(float x) {
  float r, r_inv;
  sqrt_and_inverse(x, r, r_inv);
  return r;
}

struct DensityAcc {
  float rho, rho_dh, wcount, wcount_dh, div_v, rot[3];
  __device__ __forceinline__ void zero() {
    rho = rho_dh = wcount = wcount_dh = div_v = rot[0] = rot[1] = rot[2] = 0.f;
  }
};

/* runner_iact_nonsym_density: Minimal hydro_iact.h:137-200, Gadget2 :158-222,
 * SPHENIX :141-197 (same arithmetic in all three). */
__device__ __forceinline__ void iact_density(DensityAcc &a, float r2, float dx, float dy,
                                             float dz, float hi_inv, float vix, float viy,
                                             float viz, float mj, float vjx, float vjy,
                                             float vjz) {
  float wi, wi_dx;
  float r, r_inv;
  sqrt_and_inverse(r2, r, r_inv);
  const float ui = r * hi_inv;
  kernel_deval(ui, wi, wi_dx);
  const float t = HYDRO_DIMENSION * wi + ui * wi_dx;
  a.rho += mj * wi;
  a.rho_dh -= mj * t;
  a.wcount += wi;
  a.wcount_dh -= t;
  const float faci = mj * wi_dx * r_inv;
  const float dvx = vix - vjx, dvy = viy - vjy, dvz = viz - vjz;
  const float dvdr = dvx * dx + dvy * dy + dvz * dz;
  a.div_v -= faci * dvdr;
  a.rot[0] += faci * (dvy * dz - dvz * dy);
  a.rot[1] += faci * (dvz * dx - dvx * dz);
  a.rot[2] += faci * (dvx * dy - dvy * dx);
}

/* The same for a candidate that may be masked out (evaluated unconditionally, discarded if !ok). */
__device__ __forceinline__ void iact_density_masked(DensityAcc &a, float r2, float dx, float dy, float dz,
                                                    float hi_inv, float vix, float viy, float viz, float mj,
                                                    float vjx, float vjy, float vjz, bool ok) {
  const DensityAcc old = a;
  iact_density(a, r2, dx, dy, dz, hi_inv, vix, viy, viz, mj, vjx, vjy, vjz);
  if (!ok) a = old;
}

struct GradientAcc {
  float v_sig, laplace_u, alpha_max;
};

/* runner_iact_nonsym_gradient: SPHENIX hydro_iact.h:291-350 */
__device__ __forceinline__ void iact_gradient(GradientAcc &a, float r2, float dx, float dy,
                                              float dz, float hi, float vix, float viy,
                                              float viz, float ui_, float csi, float mj,
                                              float vjx, float vjy, float vjz, float uj,
                                              float rhoj, float csj, float alphaj,
                                              float a2_Hubble) {
  float r, r_inv;
  sqrt_and_inverse(r2, r, r_inv);
  const float fac_mu = 1.f; /* pow_three_gamma_minus_five_over_two, gamma = 5/3 */
  const float dvdr = (vix - vjx) * dx + (viy - vjy) * dy + (viz - vjz) * dz;
  const float dvdr_Hubble = dvdr + a2_Hubble * r2;
  const float omega_ij = fminf(dvdr_Hubble, 0.f);
  const float mu_ij = fac_mu * r_inv * omega_ij;
  const float new_v_sig = csi + csj - CONST_VISCOSITY_BETA * mu_ij;
  a.v_sig = fmaxf(a.v_sig, new_v_sig);
  float wi, wi_dx;
  const float ui = r * rcp_rn(hi);
  kernel_deval(ui, wi, wi_dx);
  const float delta_u_factor = (ui_ - uj) * r_inv;
  a.laplace_u += mj * delta_u_factor * wi_dx * rcp_rn(rhoj);
  a.alpha_max = fmaxf(a.alpha_max, alphaj);
}

struct ForceAcc {
  float ax, ay, az, u_dt, h_dt, v_sig;
  int min_ngb;
};

/* Quantities of one particle read by the force interaction. */
struct ForceQ {
  float m, vx, vy, vz;
  float rho, P, f, cs; /* P = pressure (Minimal, SPHENIX) or P_over_rho2 (Gadget2) */
  float balsara, h, u, alpha_visc, alpha_diff;
  int time_bin;
};

/* runner_iact_nonsym_force: Minimal hydro_iact.h:378-520, Gadget2 :632-760,
 * SPHENIX :507-640; runner_iact_nonsym_timebin timestep_limiter_iact.h:41-55 */
template <int SCHEME>
__device__ __forceinline__ void iact_force(ForceAcc &a, float r2, float dx, float dy, float dz,
                                           const ForceQ &pi, const ForceQ &pj, float a2_Hubble) {
  const float fac_mu = 1.f;
  float r, r_inv;
  sqrt_and_inverse(r2, r, r_inv);
  const float mj = pj.m;
  const float rhoi = pi.rho, rhoj = pj.rho;
  const float rhoj_inv = rcp_rn(rhoj);
  const float hi_inv = rcp_rn(pi.h);
  const float hid_inv = hi_inv * hi_inv * hi_inv * hi_inv;
  const float xi = r * hi_inv;
  float wi, wi_dx;
  kernel_deval(xi, wi, wi_dx);
  const float wi_dr = hid_inv * wi_dx;
  const float hj_inv = rcp_rn(pj.h);
  const float hjd_inv = hj_inv * hj_inv * hj_inv * hj_inv;
  const float xj = r * hj_inv;
  float wj, wj_dx;
  kernel_deval(xj, wj, wj_dx);
  const float wj_dr = hjd_inv * wj_dx;
  const float dvdr = (pi.vx - pj.vx) * dx + (pi.vy - pj.vy) * dy + (pi.vz - pj.vz) * dz;
  const float dvdr_Hubble = dvdr + a2_Hubble * r2;
  const float omega_ij = fminf(dvdr_Hubble, 0.f);
  const float mu_ij = fac_mu * r_inv * omega_ij;
  const float v_sig = pi.cs + pj.cs - CONST_VISCOSITY_BETA * mu_ij;
  const float balsara_i = pi.balsara, balsara_j = pj.balsara;
  if (SCHEME == SCH_MINIMAL) {
    const float mi = pi.m;
    const float f_ij = 1.f - pi.f * rcp_rn(mj);
    const float f_ji = 1.f - pj.f * rcp_rn(mi); /* target-only: hoisted out of the candidate loop */
    const float P_over_rho2_i = pi.P * rcp_rn(rhoi * rhoi) * f_ij;
    const float P_over_rho2_j = pj.P * (rhoj_inv * rhoj_inv) * f_ji;
    const float rho_ij = 0.5f * (rhoi + rhoj);
    const float visc = -0.25f * v_sig * (balsara_i + balsara_j) * mu_ij * rcp_rn(rho_ij);
    const float visc_acc_term = 0.5f * visc * (wi_dr * f_ij + wj_dr * f_ji) * r_inv;
    const float sph_acc_term = (P_over_rho2_i * wi_dr + P_over_rho2_j * wj_dr) * r_inv;
    const float acc = sph_acc_term + visc_acc_term;
    a.ax -= mj * acc * dx;
    a.ay -= mj * acc * dy;
    a.az -= mj * acc * dz;
    const float sph_du_term_i = P_over_rho2_i * dvdr * r_inv * wi_dr;
    const float visc_du_term = 0.5f * visc_acc_term * dvdr_Hubble;
    a.u_dt += (sph_du_term_i + visc_du_term) * mj;
    a.h_dt -= mj * dvdr * r_inv * rhoj_inv * wi_dr * f_ij;
    a.v_sig = fmaxf(a.v_sig, v_sig);
  } else if (SCHEME == SCH_GADGET2) {
    const float rho_ij = 0.5f * (rhoi + rhoj);
    const float visc = -0.25f * v_sig * mu_ij * (balsara_i + balsara_j) * rcp_rn(rho_ij);
    const float visc_term = 0.5f * visc * (wi_dr + wj_dr) * r_inv;
    const float sph_term = (pi.f * pi.P * wi_dr + pj.f * pj.P * wj_dr) * r_inv;
    const float acc = visc_term + sph_term;
    a.ax -= mj * acc * dx;
    a.ay -= mj * acc * dy;
    a.az -= mj * acc * dz;
    a.h_dt -= mj * dvdr * r_inv * rhoj_inv * wi_dr;
    a.v_sig = fmaxf(a.v_sig, v_sig);
    a.u_dt += mj * visc_term * dvdr_Hubble; /* entropy_dt */
  } else {
    const float mi = pi.m;
    const float f_ij = 1.f - pi.f * rcp_rn(mj);
    const float f_ji = 1.f - pj.f * rcp_rn(mi); /* target-only: hoisted out of the candidate loop */
    const float rho_ij = rhoi + rhoj;
    const float rho_ij_inv = rcp_rn(rho_ij);
    const float alpha = pi.alpha_visc + pj.alpha_visc;
    const float visc = -0.25f * alpha * v_sig * mu_ij * (balsara_i + balsara_j) * rho_ij_inv;
    const float visc_acc_term = 0.5f * visc * (wi_dr * f_ij + wj_dr * f_ji) * r_inv;
    const float P_over_rho2_i = pi.P * rcp_rn(rhoi * rhoi) * f_ij;
    const float P_over_rho2_j = pj.P * (rhoj_inv * rhoj_inv) * f_ji;
    const float sph_acc_term = (P_over_rho2_i * wi_dr + P_over_rho2_j * wj_dr) * r_inv;
    const float acc = sph_acc_term + visc_acc_term;
    a.ax -= mj * acc * dx;
    a.ay -= mj * acc * dy;
    a.az -= mj * acc * dz;
    const float sph_du_term_i = P_over_rho2_i * dvdr * r_inv * wi_dr;
    const float visc_du_term = 0.5f * visc_acc_term * dvdr_Hubble;
    const float alpha_diff = (pi.P * pi.alpha_diff + pj.P * pj.alpha_diff) * rcp_rn(pi.P + pj.P);
    const float v_diff = alpha_diff * 0.5f *
                         (sqrt_newton(2.f * fabsf(pi.P - pj.P) * rho_ij_inv) +
                          fabsf(fac_mu * r_inv * dvdr_Hubble));
    const float diff_du_term =
        v_diff * (pi.u - pj.u) * (f_ij * wi_dr * rcp_rn(rhoi) + f_ji * wj_dr * rhoj_inv);
    a.u_dt += (sph_du_term_i + visc_du_term + diff_du_term) * mj;
    a.h_dt -= mj * dvdr * r_inv * rhoj_inv * wi_dr;
  }
  if (pj.time_bin > 0) a.min_ngb = min(a.min_ngb, pj.time_bin);
}

/* iact_force for a candidate that may be masked out: evaluated unconditionally (so that two
 * candidates interleave in one basic block), the accumulators keep their old values if !ok. */
template <int SCHEME>
__device__ __forceinline__ void iact_force_masked(ForceAcc &a, float r2, float dx, float dy, float dz,
                                                  const ForceQ &pi, const ForceQ &pj, float a2_Hubble, bool ok) {
  const ForceAcc old = a;
  iact_force<SCHEME>(a, r2, dx, dy, dz, pi, pj, a2_Hubble);
  if (!ok) a = old;
}

}  // namespace swiftgpu
#endif
