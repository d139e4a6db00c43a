/*
 * kernels_drift.cuh - cell_drift_part on the device (included by swiftgpu.cu).
 *
 * One warp per LOCAL leaf cell, lane = particle: drift_part (src/drift.h:141-215) + the per-particle
 * tail of cell_drift_part (src/cell_drift.c:255-375) directly on the device copies of the caller's
 * struct part[] / struct xpart[] (host order, addressed through the layouts), then the cell
 * reductions h_max, h_max_active, dx_max_part, dx_max_sort of the leaf and, by atomic max, of all
 * its ancestors (:219-232, :380-390). HBM-bound: one read-modify-write sweep of both AoS arrays.
 *
 * Arithmetic follows the reference's C expression by expression (double products for x and v, the
 * float Horner form of approx_expf, IEEE division and square root, no contraction), so that
 * positions, velocities, offsets, h, u and rho come out bit-identical; only cbrtf (Gadget2's
 * pow_gamma) and expf (|w| >= 0.2, a particle changing h by > 20 % in one drift) are the CUDA
 * library's instead of glibc's.
 */
#ifndef SWIFTGPU_KERNELS_DRIFT_CUH
#define SWIFTGPU_KERNELS_DRIFT_CUH

struct DriftArgs {
  char *aos;
  char *xaos;
  DevLayout D;
  swiftgpu_xpart_layout X;
  DevCell *cells;
  int ncells;
  const int32_t *d2h;
  double dt_drift, dt_kick_hydro, dt_therm;
  float min_u;
  float h_max, h_min;
  int init_particles;
  int max_active_bin;
  int64_t n_host; /* rows of the AoS copies that exist (local particles) */
};

/* approx_expf, src/approx_math.h:35-37 */
__device__ __forceinline__ float approx_expf_rn(float x) {
  const float c6 = 1.f / 6.f, c24 = 1.f / 24.f;
  float t = __fadd_rn(c6, __fmul_rn(c24, x));
  t = __fadd_rn(0.5f, __fmul_rn(x, t));
  t = __fadd_rn(1.f, __fmul_rn(x, t));
  return __fadd_rn(1.f, __fmul_rn(x, t));
}

/* Cells that will be recomputed start from zero; empty cells keep their values (cell_drift.c:192-200). */
__global__ void k_drift_begin(DevCell *cells, int ncells) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  DevCell &C = cells[c];
  if (!(C.flags & 2) || C.count == 0) return;
  C.h_max = 0.f;
  C.h_max_active = 0.f;
  C.dx_max_part = 0.f;
  C.dx_max_sort = 0.f;
}

template <int SCHEME>
__global__ void __launch_bounds__(128) k_drift(const DriftArgs A) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= A.ncells) return;
  const DevCell C = A.cells[c];
  if ((C.flags & 4) || !(C.flags & 2) || C.count == 0) return; /* split, foreign or empty */
  const swiftgpu_part_layout &L = A.D.L;
  const float dtd = (float)A.dt_drift, dtt = (float)A.dt_therm; /* hydro_predict_extra takes floats */
  float dx2_max = 0.f, dx2_max_sort = 0.f, h_max = 0.f, h_max_active = 0.f;
  for (int k = lane; k < C.count; k += 32) {
    const int64_t row = A.d2h[C.first + k];
    if (row >= A.n_host) continue;
    char *b = A.aos + (size_t)L.size * (size_t)row;
    char *xb = A.xaos + (size_t)A.X.size * (size_t)row;
    const int tb = rd<int8_t>(b, L.time_bin);
    if (tb == 58) continue; /* time_bin_inhibited: part_is_inhibited */
    /* ---- drift_part ---- */
    float vfull[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      vfull[a] = rd<float>(xb, A.X.v_full + 4 * a);
      const double x = rd<double>(b, L.x + 8 * a);
      wr<double>(b, L.x + 8 * a, __dadd_rn(x, __dmul_rn((double)vfull[a], A.dt_drift)));
      const float v = rd<float>(b, L.v + 4 * a), acc = rd<float>(b, L.a_hydro + 4 * a);
      wr<float>(b, L.v + 4 * a, (float)__dadd_rn((double)v, __dmul_rn((double)acc, A.dt_kick_hydro)));
    }
    /* ---- hydro_predict_extra ---- */
    float h = rd<float>(b, L.h), rho = rd<float>(b, L.rho);
    const float h_inv = __fdiv_rn(1.f, h);
    const float w1 = __fmul_rn(__fmul_rn(rd<float>(b, L.h_dt), h_inv), dtd);
    h = __fmul_rn(h, fabsf(w1) < 0.2f ? approx_expf_rn(w1) : expf(w1));
    const float w2 = __fmul_rn(-HYDRO_DIMENSION, w1);
    rho = __fmul_rn(rho, fabsf(w2) < 0.2f ? approx_expf_rn(w2) : expf(w2));
    float cs;
    if (SCHEME == SCH_GADGET2) {
      float A_ent = __fadd_rn(rd<float>(b, L.entropy), __fmul_rn(rd<float>(b, L.entropy_dt), dtt));
      /* entropy floor none: floor_A = 0; min_A = gas_entropy_from_internal_energy(rho, min_u) */
      A_ent = fmaxf(A_ent, 0.f);
      if (A.min_u > 0.f) {
        const float cb = cbrtf(rho);
        A_ent = fmaxf(A_ent, __fmul_rn(__fmul_rn(HYDRO_GAMMA_MINUS_ONE, A.min_u), __fdiv_rn(1.f, __fmul_rn(cb, cb))));
      }
      const float cb = cbrtf(rho);
      const float P = __fmul_rn(A_ent, __fmul_rn(__fmul_rn(cb, cb), rho)); /* entropy * pow_gamma(rho) */
      cs = __fsqrt_rn(__fdiv_rn(__fmul_rn(HYDRO_GAMMA, P), rho));
      const float rho_inv = __fdiv_rn(1.f, rho);
      wr<float>(b, L.entropy, A_ent);
      wr<float>(b, L.P_over_rho2, __fmul_rn(__fmul_rn(P, rho_inv), rho_inv));
    } else {
      float u = __fadd_rn(rd<float>(b, L.u), __fmul_rn(rd<float>(b, L.u_dt), dtt));
      u = fmaxf(u, 0.f); /* entropy floor none: floor_u = 0 */
      u = fmaxf(u, A.min_u);
      const float P = __fmul_rn(__fmul_rn(HYDRO_GAMMA_MINUS_ONE, u), rho);
      cs = __fsqrt_rn(__fdiv_rn(__fmul_rn(HYDRO_GAMMA, P), rho));
      wr<float>(b, L.u, u);
      wr<float>(b, L.pressure, P);
    }
    wr<float>(b, L.soundspeed, cs);
    wr<float>(b, L.v_sig, fmaxf(rd<float>(b, L.v_sig), __fmul_rn(2.f, cs)));
    /* ---- offsets since the last rebuild / sort ---- */
    float dx2 = 0.f, dx2s = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float dx = (float)__dmul_rn((double)vfull[a], A.dt_drift);
      const float xd = __fsub_rn(rd<float>(xb, A.X.x_diff + 4 * a), dx);
      const float xs = __fsub_rn(rd<float>(xb, A.X.x_diff_sort + 4 * a), dx);
      wr<float>(xb, A.X.x_diff + 4 * a, xd);
      wr<float>(xb, A.X.x_diff_sort + 4 * a, xs);
      dx2 = a == 0 ? __fmul_rn(xd, xd) : __fadd_rn(dx2, __fmul_rn(xd, xd));
      dx2s = a == 0 ? __fmul_rn(xs, xs) : __fadd_rn(dx2s, __fmul_rn(xs, xs));
    }
    /* ---- cell_drift_part's tail ---- */
    h = fminf(h, A.h_max);
    h = fmaxf(h, A.h_min);
    wr<float>(b, L.h, h);
    wr<int8_t>(b, L.depth_h, (int8_t)part_h_depth(A.cells, c, h, rd<int8_t>(b, L.depth_h)));
    dx2_max = fmaxf(dx2_max, dx2);
    dx2_max_sort = fmaxf(dx2_max_sort, dx2s);
    h_max = fmaxf(h_max, h);
    const bool active = tb <= A.max_active_bin;
    if (active) h_max_active = fmaxf(h_max_active, h);
    if (A.init_particles && active) {
      /* hydro_init_part: Minimal hydro.h:574, Gadget2 :560, SPHENIX :587 */
      rho = 0.f;
      wr<float>(b, L.wcount, 0.f);
      wr<float>(b, L.wcount_dh, 0.f);
      wr<float>(b, L.rho_dh, 0.f);
      wr<float>(b, L.div_v, 0.f);
      wr<float>(b, L.rot_v, 0.f);
      wr<float>(b, L.rot_v + 4, 0.f);
      wr<float>(b, L.rot_v + 8, 0.f);
      if (SCHEME == SCH_SPHENIX) wr<float>(b, L.laplace_u, 0.f);
    }
    wr<float>(b, L.rho, rho);
  }
  dx2_max = warp_max(dx2_max);
  dx2_max_sort = warp_max(dx2_max_sort);
  h_max = warp_max(h_max);
  h_max_active = warp_max(h_max_active);
  if (lane == 0) {
    const float dxm = __fsqrt_rn(dx2_max), dxs = __fsqrt_rn(dx2_max_sort);
    /* non-negative floats order like their bit patterns */
    for (int f = c; f >= 0; f = A.cells[f].parent) {
      DevCell *F = &A.cells[f];
      atomicMax((int *)&F->h_max, __float_as_int(h_max));
      atomicMax((int *)&F->h_max_active, __float_as_int(h_max_active));
      atomicMax((int *)&F->dx_max_part, __float_as_int(dxm));
      atomicMax((int *)&F->dx_max_sort, __float_as_int(dxs));
    }
  }
}

__global__ void k_get_cell_drift(const DevCell *cells, int ncells, float *out /* 4 x ncells */,
                                 float *dx_max_part /* the array k_pred_bits reads */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  dx_max_part[c] = cells[c].dx_max_part;
  out[c] = cells[c].h_max;
  out[ncells + c] = cells[c].h_max_active;
  out[2 * ncells + c] = cells[c].dx_max_part;
  out[3 * ncells + c] = cells[c].dx_max_sort;
}

/* ---- kick (runner_do_kick1 / runner_do_kick2, src/runner_time_integration.c:87,360) ---- */
struct KickArgs {
  char *aos;
  char *xaos;
  DevLayout D;
  swiftgpu_xpart_layout X;
  int64_t n;
  int which; /* 1: kick1 (part_is_starting), 2: kick2 (part_is_active) + hydro_reset_predicted_values */
  int max_active_bin;
  double time_base;
  float min_u;
};

template <int SCHEME>
__global__ void __launch_bounds__(256) k_kick(const KickArgs A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const swiftgpu_part_layout &L = A.D.L;
  char *b = A.aos + (size_t)L.size * (size_t)i;
  char *xb = A.xaos + (size_t)A.X.size * (size_t)i;
  const int tb = rd<int8_t>(b, L.time_bin);
  /* part_is_starting / part_is_active (src/active.h:349,472): time_bin <= max_active_bin */
  if (tb > A.max_active_bin) return;
  /* half of the particle's own step: get_integer_timestep (timeline.h:59), no cosmology */
  const long long ti_step = tb <= 0 ? 0ll : (1ll << (tb + 1));
  const double dt = (double)(ti_step / 2) * A.time_base;
  const float dt_therm = (float)dt;
  /* kick_part, src/kick.h:113 */
  float vf[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    vf[a] = (float)__dadd_rn((double)rd<float>(xb, A.X.v_full + 4 * a),
                             __dmul_rn((double)rd<float>(b, L.a_hydro + 4 * a), dt));
    wr<float>(xb, A.X.v_full + 4 * a, vf[a]);
  }
  /* hydro_kick_extra: Minimal hydro.h:880, Gadget2 :862, SPHENIX :1093 */
  const int rate_off = SCHEME == SCH_GADGET2 ? L.entropy_dt : L.u_dt;
  const float rho = rd<float>(b, L.rho);
  float full = rd<float>(xb, A.X.u_full);
  const float delta = __fmul_rn(rd<float>(b, rate_off), dt_therm);
  full = fmaxf(__fadd_rn(full, delta), __fmul_rn(0.5f, full));
  float floor_v = 0.f; /* entropy floor none */
  if (SCHEME == SCH_GADGET2) {
    if (A.min_u > 0.f) {
      const float cb = cbrtf(rho);
      floor_v = fmaxf(floor_v, __fmul_rn(__fmul_rn(HYDRO_GAMMA_MINUS_ONE, A.min_u), __fdiv_rn(1.f, __fmul_rn(cb, cb))));
    }
  } else {
    floor_v = fmaxf(A.min_u, 0.f);
  }
  if (full < floor_v) {
    full = floor_v;
    wr<float>(b, rate_off, 0.f);
  }
  wr<float>(xb, A.X.u_full, full);
  if (A.which != 2) return;
  /* hydro_reset_predicted_values: Minimal hydro.h:786, Gadget2 :765, SPHENIX :983 */
#pragma unroll
  for (int a = 0; a < 3; a++) wr<float>(b, L.v + 4 * a, vf[a]);
  if (SCHEME == SCH_GADGET2) {
    wr<float>(b, L.entropy, full);
    const float cb = cbrtf(rho);
    const float P = __fmul_rn(full, __fmul_rn(__fmul_rn(cb, cb), rho));
    const float cs = __fsqrt_rn(__fdiv_rn(__fmul_rn(HYDRO_GAMMA, P), rho));
    const float rho_inv = __fdiv_rn(1.f, rho);
    wr<float>(b, L.soundspeed, cs);
    wr<float>(b, L.P_over_rho2, __fmul_rn(__fmul_rn(P, rho_inv), rho_inv));
  } else {
    wr<float>(b, L.u, full);
    const float P = __fmul_rn(__fmul_rn(HYDRO_GAMMA_MINUS_ONE, full), rho);
    const float cs = __fsqrt_rn(__fdiv_rn(__fmul_rn(HYDRO_GAMMA, P), rho));
    wr<float>(b, L.pressure, P);
    wr<float>(b, L.soundspeed, cs);
    wr<float>(b, L.v_sig, fmaxf(rd<float>(b, L.v_sig), __fmul_rn(2.f, cs)));
  }
}

#endif
