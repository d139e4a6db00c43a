/*
 * kernels_finalise.cuh - the per-particle finalisers of the path (included by swiftgpu.cu): the ghost
 * (hydro_end_density + the Newton-Raphson / bisection h iteration of runner_do_ghost +
 * hydro_prepare_force | hydro_prepare_gradient), the extra ghost (SPHENIX hydro_end_gradient +
 * hydro_prepare_force), end_force + the CFL time-step, hydro_init_part, and the recursion-predicate
 * check that decides whether a worklist must follow the ghost.
 */
#ifndef SWIFTGPU_KERNELS_FINALISE_CUH
#define SWIFTGPU_KERNELS_FINALISE_CUH

/* ======================================================================== */
/* Per-particle finalisers                                                   */
/* ======================================================================== */
struct GhostArgs {
  const Group *groups; /* subset groups: one per active local leaf */
  int ngroups;
  DevCell *cells;
  int32_t *redo_list;  /* L_subset.tgt_list, indexed by particle slot */
  int32_t *redo_count; /* L_subset.tgt_count */
  Soa S;
  float *left, *right;
  int32_t *nd, *ng, *nf;
  unsigned long long *n_redo;
  int first_pass;
  int max_active_bin;
  float h_max, h_min, eps, eta_dim;
  int use_mass_weighted;
  float visc_alpha; /* hydro_props->viscosity.alpha (Minimal/Gadget2 Balsara prefactor) */
  float H, a;
  float num_reruns;
};

/* cell_set_part_h_depth, cell.h:1787-1815 */
__device__ __forceinline__ int part_h_depth(const DevCell *cells, int leaf, float h, int current) {
  const DevCell *c = &cells[leaf];
  if (h < c->h_min_allowed) return c->depth;
  int ci = leaf;
  while (ci >= 0) {
    c = &cells[ci];
    if (h >= c->h_min_allowed && h < c->h_max_allowed) return c->depth;
    ci = c->parent;
  }
  return current;
}

/* hydro_prepare_force + hydro_reset_acceleration (Minimal hydro.h:669-766,
 * Gadget2 hydro.h:648-744) or hydro_prepare_gradient + hydro_reset_gradient
 * (SPHENIX hydro.h:671-755), from the finished density sums. */
template <int SCHEME>
__device__ __forceinline__ void ghost_finalise(const GhostArgs &G, int p, float h, float rho,
                                               float rho_dh, float wcount, float wcount_dh,
                                               float div_v, float rx, float ry, float rz) {
  const Soa &S = G.S;
  const float u = S.u[p];
  const int tb = S.time_bin[p];
  const float curl_v = sqrtf(rx * rx + ry * ry + rz * rz);
  float f, P, cs, balsara;
  if (SCHEME == SCH_GADGET2) {
    const float rho_inv = 1.f / rho;
    const float h_inv = 1.f / h;
    const float abs_div = fabsf(div_v + HYDRO_DIMENSION * G.H);
    const float cb = cbrtf(rho);
    const float pressure = u * (cb * cb * rho); /* entropy * pow_gamma(rho) */
    cs = sqrtf(HYDRO_GAMMA * pressure / rho);
    P = pressure * rho_inv * rho_inv;
    balsara = G.visc_alpha * abs_div / (abs_div + curl_v + 0.0001f * cs * h_inv);
    float rdh = rho_dh;
    if (h > 0.9999f * G.h_max) rdh = 0.f;
    const float grad_rho_term = HYDRO_DIMENSION_INV * h * rdh * rho_inv;
    f = (grad_rho_term < -0.9999f) ? 1.f : 1.f / (1.f + grad_rho_term);
  } else {
    P = HYDRO_GAMMA_MINUS_ONE * u * rho;
    cs = sqrtf(HYDRO_GAMMA * P / rho);
    const float common_factor = h * HYDRO_DIMENSION_INV / wcount;
    if (h > 0.9999f * G.h_max) {
      f = 0.f;
    } else {
      const float grad_W_term = common_factor * wcount_dh;
      f = (grad_W_term < -0.9999f) ? 0.f : common_factor * rho_dh / (1.f + grad_W_term);
    }
    if (SCHEME == SCH_MINIMAL) {
      const float h_inv = 1.f / h;
      const float abs_div = fabsf(div_v + HYDRO_DIMENSION * G.H);
      balsara = G.visc_alpha * abs_div / (abs_div + curl_v + 0.0001f * cs * h_inv);
    } else {
      const float abs_div = fabsf(div_v);
      balsara = abs_div / (abs_div + curl_v + 0.0001f * cs * 1.f / h);
    }
  }
  S.rho[p] = rho;
  S.fq1[p] = make_float4(rho, P, f, cs);
  S.fq2[p] = make_float4(balsara, h, SCHEME == SCH_SPHENIX ? u : hg2_exact(h), __int_as_float(tb));
  if (SCHEME == SCH_SPHENIX) {
    const float al = S.alpha[p];
    S.fq3[p] = make_float4(al, S.alpha_diff[p], hg2_exact(h), 0.f);
    S.div_v[p] = div_v;
    S.g_vsig[p] = 2.f * cs; /* hydro_reset_gradient */
    S.g_amax[p] = al;
    G.ng[p] = 0;
  } else {
    S.fo1[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    S.f_hdt[p] = 0.f;
    S.f_vsig[p] = 2.f * cs;
    S.f_minngb[p] = NUM_TIME_BINS + 1; /* timestep_limiter_prepare_force */
    G.nf[p] = 0;
  }
  /* keep the finished density members for a density-level download */
  S.dA[p] = make_float4(rho, rho_dh, wcount, wcount_dh);
  S.dB[p] = make_float4(div_v, rx, ry, rz);
}

/* runner_do_ghost leaf loop, runner_ghost.c:1197-1538. One warp per leaf. */
template <int SCHEME>
__global__ void __launch_bounds__(128) k_ghost(const GhostArgs G) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= G.ngroups) return;
  const int leaf = G.groups[g].tcell;
  const DevCell c = G.cells[leaf];
  const Soa &S = G.S;
  const int n_in = G.first_pass ? c.count : G.redo_count[g];
  int32_t *list = G.redo_list + c.first;
  int nout = 0;
  float hmax_conv = 0.f;
  bool any_conv = false;
  for (int base = 0; base < n_in; base += 32) {
    const int k = base + lane;
    bool valid = k < n_in;
    int p = -1;
    if (valid) p = G.first_pass ? c.first + k : list[k];
    if (valid && G.first_pass) valid = S.time_bin[p] <= G.max_active_bin;
    bool redo = false;
    if (valid) {
      float left = G.first_pass ? 0.f : G.left[p];
      float right = G.first_pass ? G.h_max : G.right[p];
      const float h_old = S.h[p];
      const float h_old_dim = h_old * h_old * h_old;
      const float h_old_dim_minus_one = h_old * h_old;
      const float4 a = S.dA[p], b = S.dB[p];
      const float m = S.mv[p].x;
      float rho = a.x, rho_dh = a.y, wcount = a.z, wcount_dh = a.w;
      float div_v = b.x, rx = b.y, ry = b.z, rz = b.w;
      float h_new = 0.f;
      bool has_no_ngb = false;
      bool done_early = false;
      if (wcount < 1.e-5 * (double)KERNEL_ROOT) {
        has_no_ngb = true;
        h_new = 2.f * h_old;
      } else {
        /* hydro_end_density: Minimal :543, Gadget2 :526, SPHENIX :613 */
        const float h_inv = 1.0f / h_old;
        const float h_inv_dim = h_inv * h_inv * h_inv;
        const float h_inv_dim_plus_one = h_inv_dim * h_inv;
        rho += m * KERNEL_ROOT;
        rho_dh -= HYDRO_DIMENSION * m * KERNEL_ROOT;
        wcount += KERNEL_ROOT;
        wcount_dh -= HYDRO_DIMENSION * KERNEL_ROOT;
        rho *= h_inv_dim;
        rho_dh *= h_inv_dim_plus_one;
        wcount *= h_inv_dim;
        wcount_dh *= h_inv_dim_plus_one;
        const float rho_inv = 1.f / rho;
        const float a_inv2 = 1.f / (G.a * G.a);
        const float fac = h_inv_dim_plus_one * a_inv2 * rho_inv;
        rx *= fac;
        ry *= fac;
        rz *= fac;
        if (SCHEME == SCH_SPHENIX) {
          div_v *= h_inv_dim_plus_one * rho_inv * a_inv2;
          div_v += G.H * HYDRO_DIMENSION;
        } else {
          div_v *= fac;
        }
        if (G.use_mass_weighted) {
          const float inv_mass = 1.f / m;
          wcount = rho * inv_mass;
          wcount_dh = rho_dh * inv_mass;
        }
        const float n_sum = wcount * h_old_dim;
        const float n_target = G.eta_dim;
        const float f = n_sum - n_target;
        const float f_prime = wcount_dh * h_old_dim + HYDRO_DIMENSION * wcount * h_old_dim_minus_one;
        if (n_sum < n_target)
          left = fmaxf(left, h_old);
        else if (n_sum > n_target)
          right = fminf(right, h_old);
        if (((h_old >= G.h_max) && (f < 0.f)) || ((h_old <= G.h_min) && (f > 0.f))) {
          /* already at the limit: tidy up as if converged (:1271-1352) */
          ghost_finalise<SCHEME>(G, p, h_old, rho, rho_dh, wcount, wcount_dh, div_v, rx, ry, rz);
          hmax_conv = fmaxf(hmax_conv, h_old);
          any_conv = true;
          done_early = true;
        } else {
          h_new = h_old - f / (f_prime + 1.17549435e-38f);
          h_new = fminf(h_new, 2.f * h_old);
          h_new = fmaxf(h_new, 0.5f * h_old);
          h_new = fmaxf(h_new, left);
          h_new = fminf(h_new, right);
        }
      }
      if (!done_early) {
        float h_final = h_old;
        if (fabsf(h_new - h_old) > G.eps * h_old) {
          float h_set;
          if ((h_new == left && h_old == right) || (h_old == left && h_new == right)) {
            h_set = cbrtf(0.5f * (left * left * left + right * right * right));
          } else {
            h_set = h_new;
          }
          if (h_set < G.h_max && h_set > G.h_min) {
            redo = true;
            S.h[p] = h_set;
            G.left[p] = left;
            G.right[p] = right;
            /* hydro_init_part */
            S.dA[p] = make_float4(0.f, 0.f, 0.f, 0.f);
            S.dB[p] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (SCHEME == SCH_SPHENIX) S.g_lap[p] = 0.f;
            G.nd[p] = 0;
          } else if (h_set <= G.h_min) {
            h_final = G.h_min;
          } else {
            h_final = G.h_max;
            if (has_no_ngb) {
              /* hydro_part_has_no_neighbours */
              const float h_inv = 1.0f / h_final;
              const float h_inv_dim = h_inv * h_inv * h_inv;
              rho = m * KERNEL_ROOT * h_inv_dim;
              wcount = KERNEL_ROOT * h_inv_dim;
              rho_dh = wcount_dh = div_v = rx = ry = rz = 0.f;
            }
          }
        }
        if (!redo) {
          S.h[p] = h_final;
          S.depth_h[p] = (int8_t)part_h_depth(G.cells, leaf, h_final, S.depth_h[p]);
          hmax_conv = fmaxf(hmax_conv, h_final);
          any_conv = true;
          ghost_finalise<SCHEME>(G, p, h_final, rho, rho_dh, wcount, wcount_dh, div_v, rx, ry, rz);
        }
      }
    }
    __syncwarp();
    const unsigned m = __ballot_sync(FULL_MASK, redo);
    if (redo) list[nout + __popc(m & ((1u << lane) - 1u))] = p;
    nout += __popc(m);
    __syncwarp();
  }
  hmax_conv = warp_max(hmax_conv);
  const bool anyc = __any_sync(FULL_MASK, any_conv);
  if (lane == 0) {
    G.redo_count[g] = nout;
    if (nout) atomicAdd(G.n_redo, (unsigned long long)nout);
    if (anyc) {
      /* atomic_max_f on the leaf and all its parents (:1621-1632) */
      for (int ci = leaf; ci >= 0; ci = G.cells[ci].parent) {
        atomic_max_pos(&G.cells[ci].h_max, hmax_conv);
        atomic_max_pos(&G.cells[ci].h_max_active, hmax_conv);
      }
    }
  }
}

/* runner_do_extra_ghost (runner_ghost.c:1016): hydro_end_gradient +
 * hydro_prepare_force + hydro_reset_acceleration, SPHENIX hydro.h:762-972. */
struct ExtraArgs {
  const Group *groups;
  int ngroups;
  const DevCell *cells;
  Soa S;
  int32_t *nf;
  int max_active_bin;
  double time_base;
  float a;
  float alpha_max, alpha_min, length, beta, diff_alpha_max, diff_alpha_min;
};
__global__ void __launch_bounds__(128) k_extra_ghost(const ExtraArgs E) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= E.ngroups) return;
  const DevCell c = E.cells[E.groups[g].tcell];
  const Soa &S = E.S;
  for (int k = lane; k < c.count; k += 32) {
    const int p = c.first + k;
    const int bin = S.time_bin[p];
    if (bin > E.max_active_bin) continue;
    const float h = S.h[p];
    const float h_inv = 1.0f / h;
    const float h_inv_dim_plus_one = h_inv * h_inv * h_inv * h_inv;
    float laplace_u = S.g_lap[p] * (2.f * h_inv_dim_plus_one);
    /* get_timestep (timeline.h:91), passed as a float argument */
    const float dt_alpha = (float)((double)(bin <= 0 ? 0LL : 1LL << (bin + 1)) * E.time_base);
    const float4 q1 = S.fq1[p];
    const float rho = q1.x;
    const float u = S.u[p];
    const float div_v = S.div_v[p];
    const float kernel_support_physical = h * E.a * KERNEL_GAMMA;
    const float kernel_support_physical_inv = 1.f / kernel_support_physical;
    const float v_sig_physical = S.g_vsig[p];
    const float pressure = HYDRO_GAMMA_MINUS_ONE * u * rho;
    const float soundspeed_physical = sqrtf(HYDRO_GAMMA * pressure / rho);
    const float sound_crossing_time_inverse = soundspeed_physical * kernel_support_physical_inv;
    const float div_v_dt = dt_alpha == 0.f ? 0.f : (div_v - S.div_v_prev[p]) / dt_alpha;
    const float Sterm = div_v < 0.f ? kernel_support_physical * kernel_support_physical *
                                          fmaxf(0.f, -1.f * div_v_dt)
                                    : 0.f;
    const float soundspeed_square = soundspeed_physical * soundspeed_physical;
    const float alpha_loc = E.alpha_max * Sterm / (soundspeed_square + Sterm);
    float alpha = S.alpha[p];
    if (alpha_loc > alpha) {
      alpha = alpha_loc;
    } else {
      const float timescale_ratio = dt_alpha * sound_crossing_time_inverse * E.length;
      alpha += alpha_loc * timescale_ratio;
      alpha /= (1.f + timescale_ratio);
    }
    alpha = fmaxf(alpha, E.alpha_min);
    S.alpha[p] = alpha;
    S.div_v_prev[p] = div_v;
    S.div_v_dt[p] = div_v_dt;
    const float diffusion_timescale_physical_inverse = v_sig_physical * kernel_support_physical_inv;
    const float sqrt_u_inv = 1.f / sqrtf(u);
    float alpha_diff_dt = E.beta * kernel_support_physical * laplace_u * sqrt_u_inv * (1.f / (E.a * E.a));
    const float ad = S.alpha_diff[p];
    alpha_diff_dt -= (ad - E.diff_alpha_min) * diffusion_timescale_physical_inverse;
    float new_ad = ad + alpha_diff_dt * dt_alpha;
    new_ad = fmaxf(new_ad, E.diff_alpha_min);
    const float viscous_diffusion_limit = E.diff_alpha_max * (1.f - S.g_amax[p] / E.alpha_max);
    new_ad = fminf(new_ad, viscous_diffusion_limit);
    S.alpha_diff[p] = new_ad;
    S.g_lap[p] = laplace_u;
    S.fq3[p] = make_float4(alpha, new_ad, hg2_exact(h), 0.f);
    /* hydro_reset_acceleration + timestep_limiter_prepare_force */
    S.fo1[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    S.f_hdt[p] = 0.f;
    S.f_minngb[p] = NUM_TIME_BINS + 1;
    E.nf[p] = 0;
  }
}

/* runner_do_end_hydro_force (runner_others.c:815): hydro_end_force, and in the
 * same pass hydro_compute_timestep (Minimal hydro.h:440, Gadget2 :444, SPHENIX
 * :475) with the reference's order of operations (separate IEEE multiplies and
 * one divide), so that dt is bit-identical for identical h and v_sig. */
__global__ void __launch_bounds__(128)
    k_end_force(const Group *groups, int ngroups, const DevCell *cells, Soa S, int max_active_bin,
                int scheme, float cfl, float a, float a_factor_sound_speed, float *dt_cfl) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  const DevCell c = cells[groups[g].tcell];
  for (int k = lane; k < c.count; k += 32) {
    const int p = c.first + k;
    if (S.time_bin[p] > max_active_bin) {
      dt_cfl[p] = -1.f;
      continue;
    }
    const float h = S.h[p];
    S.f_hdt[p] *= h * HYDRO_DIMENSION_INV;
    if (scheme == SCH_GADGET2) {
      /* 0.5 * gas_entropy_from_internal_energy(rho, entropy_dt) */
      const float cbrt_inv = 1.f / cbrtf(S.rho[p]);
      float4 o = S.fo1[p];
      o.w = 0.5f * (HYDRO_GAMMA_MINUS_ONE * o.w * (cbrt_inv * cbrt_inv));
      S.fo1[p] = o;
    }
    const float v_sig = scheme == SCH_SPHENIX ? S.g_vsig[p] : S.f_vsig[p];
    const float num = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.f, KERNEL_GAMMA), cfl), a), h);
    dt_cfl[p] = __fdiv_rn(num, __fmul_rn(a_factor_sound_speed, v_sig));
  }
}

/* hydro_init_part for the active particles (cell_drift.c:361) */
__global__ void k_init_parts(Soa S, int32_t *nd, int64_t n, int max_active_bin, int scheme) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (S.time_bin[p] > max_active_bin) return;
  S.dA[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  S.dB[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scheme == SCH_SPHENIX) S.g_lap[p] = 0.f;
  nd[p] = 0;
}

/* Which h_max-dependent recursion predicates does each cell satisfy NOW?
 * loop 2: cell.h:966 subpair2, :1007 subself2 (h_max, dx_max_part);
 * loop 1: cell.h:951 subpair, :992 subself (h_max_active, dx_max_part_old).
 * Compared with the bits the worklist was built with (Flattener::subpair*);
 * a mismatch triggers a rebuild of that list on the host. */
__global__ void k_pred_bits(const DevCell *cells, const float *dmin, const float *dx_max_part,
                            const uint8_t *bits, int ncells, int use_active, int32_t *flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint8_t b = 0;
  if (((cells[c].flags >> 2) & 1) && cells[c].count >= 100) { /* Flattener::recursable */
    const float hm = use_active ? cells[c].h_max_active : cells[c].h_max;
    const float gh = __fmul_rn(KERNEL_GAMMA, hm);
    const float half = __fmul_rn(0.5f, dmin[c]);
    b = (uint8_t)((__fadd_rn(gh, dx_max_part[c]) < half) ? 1 : 0) | (uint8_t)((gh < half) ? 2 : 0);
  }
  if (b != bits[c]) *flag = 1;
}

__global__ void k_get_cell_hmax(const DevCell *cells, int ncells, float *h_max, float *h_max_active) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  h_max[c] = cells[c].h_max;
  h_max_active[c] = cells[c].h_max_active;
}


#endif
