/*
 * loops_direct.cuh - the density loop for SPARSE target sets: one warp per
 * target particle, sources read straight from global memory (L2).
 *
 * The late iterations of the ghost re-run the density loop for the handful of
 * particles per leaf whose h has not converged (runner_ghost.c:1548-1572:
 * DOSELF_SUBSET / DOPAIR_SUBSET over the cells of the leaf's density tasks).
 * Streaming all 27 source leaves through the TMA ring of loops_pipe.cuh for
 * five targets costs as much as for sixty-four; here the cost is per target:
 * lane = source particle, 32 at a time, after a cell-level and a 4-octet box
 * cull. The frame floats and double columns are the same arrays the pipeline
 * stages, the accept arithmetic is the same (functions_hydro.h:891-1000,
 * :1108-1200, :1327-1347): frame modes test tp - F exactly; double modes
 * prefilter on the source cell's own frame and then evaluate
 * (float)((x_t - shift) - x_s) on the doubles; the sorted-axis conditions are
 * checked (exact_type1) only within keyE of the cut-off. One warp owns a
 * target, so its sums are reduced by shuffles and stored without atomics.
 */
#ifndef SWIFTGPU_LOOPS_DIRECT_CUH
#define SWIFTGPU_LOOPS_DIRECT_CUH

#include "loops_pipe.cuh"

namespace swiftgpu {

/* flat list of the targets of a launch: (particle, group), compacted by k_flat_targets */
__global__ void __launch_bounds__(128)
    k_flat_targets(const Group *groups, int ngroups, const int32_t *tgt_first, const int32_t *tgt_count,
                   const int32_t *tgt_list, int2 *flat, unsigned int *nflat, const unsigned long long *gate,
                   unsigned long long gate_hi) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  if (gate && *gate >= gate_hi) return; /* too many targets for this kernel: the pipeline takes the pass */
  const int nt = tgt_count[g];
  if (nt <= 0) return;
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(nflat, (unsigned)nt);
  base = __shfl_sync(FULL_MASK, base, 0);
  for (int k = lane; k < nt; k += 32) flat[base + k] = make_int2(tgt_list[tgt_first[g] + k], g);
}

template <int LOOP>
__global__ void __launch_bounds__(256) k_direct(const LoopArgs A, const int2 *flat, const unsigned int *nflat) {
  static_assert(LOOP == LOOP_DENSITY, "direct loop: density (and its ghost re-runs) only");
  const int lane = threadIdx.x & 31;
  const unsigned nt = *nflat;
  for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nt; w += (gridDim.x * blockDim.x) >> 5) {
    const int2 tg = flat[w];
    const int ti = tg.x;
    const Group G = A.groups[tg.y];
    const DevCell tcell = A.cells[G.tcell];
    const double tx = A.xs0[ti], ty = A.xs1[ti], tz = A.xs2[ti];
    const float th = A.h[ti];
    const float4 tmv = A.mv[ti];
    const int tdepth = A.depth_h[ti];
    const float thg2 = hg2_exact(th);
    const float th_inv = 1.f / th;
    const float thg = __fmul_rn(th, KERNEL_GAMMA);
    const float tsure2 = sure_r2(thg, A.keyE);
    const float re = fmaf(thg, PREFILTER_REL, A.margin);
    const float r2e = re * re;
    /* own-frame position of the target in its group's cell (culls) */
    const float tox = dsubf(tx, tcell.loc[0]), toy = dsubf(ty, tcell.loc[1]), toz = dsubf(tz, tcell.loc[2]);
    DensityAcc acc;
    acc.zero();
    int nhit = 0, ntests = 0;
    for (int k = 0; k < G.item_count; k++) {
      const int item = G.item_first + k;
      const Item I = A.items[item];
      if (tdepth < I.min_depth || tdepth > I.max_depth) continue;
      const DevCell sc = A.cells[I.scell];
      if (sc.count <= 0) continue;
      const int mode = I.mode;
      const double shx = I.shift[0] * A.dim[0], shy = I.shift[1] * A.dim[1], shz = I.shift[2] * A.dim[2];
      double o0, o1, o2, s0 = 0., s1 = 0., s2 = 0., ex = 0., ey = 0., ez = 0.;
      float hcap = 3.402823466e+38f;
      bool dbl = false, nokey = false;
      if (mode == MODE_PAIR_L || mode == MODE_PAIR_R) {
        if (mode == MODE_PAIR_L) {
          o0 = __dadd_rn(sc.loc[0], shx); o1 = __dadd_rn(sc.loc[1], shy); o2 = __dadd_rn(sc.loc[2], shz);
          ex = shx; ey = shy; ez = shz;
        } else {
          o0 = tcell.loc[0]; o1 = tcell.loc[1]; o2 = tcell.loc[2];
          ex = -shx; ey = -shy; ez = -shz;
        }
        const float ci_hma = (mode == MODE_PAIR_L) ? tcell.h_max_allowed : sc.h_max_allowed;
        const float h_max_lim = (I.flags & 1) ? ci_hma : 3.402823466e+38f;
        hcap = __fmul_rn(fminf(h_max_lim, tcell.h_max_active), KERNEL_GAMMA);
      } else if (mode == MODE_SUB_SELF) {
        o0 = sc.loc[0]; o1 = sc.loc[1]; o2 = sc.loc[2];
        nokey = true;
      } else {
        dbl = true;
        if (mode != MODE_SELF) {
          s0 = shx; s1 = shy; s2 = shz;
          ex = shx; ey = shy; ez = shz;
        } else {
          nokey = true;
        }
        o0 = __dadd_rn(sc.loc[0], s0); o1 = __dadd_rn(sc.loc[1], s1); o2 = __dadd_rn(sc.loc[2], s2);
      }
      /* cell-level cull in the source cell's own frame */
      const float d0 = (float)__dsub_rn(__dadd_rn(sc.loc[0], ex), tcell.loc[0]);
      const float d1 = (float)__dsub_rn(__dadd_rn(sc.loc[1], ey), tcell.loc[1]);
      const float d2 = (float)__dsub_rn(__dadd_rn(sc.loc[2], ez), tcell.loc[2]);
      const float px = tox - d0, py = toy - d1, pz = toz - d2; /* target in the source cell's own frame */
      const float rc = fmaf(thg, PREFILTER_REL, A.margin);
      {
        const float r = rc + sc.dx_max_part;
        const float gx = fmaxf(0.f, fmaxf(-px, px - sc.width));
        const float gy = fmaxf(0.f, fmaxf(-py, py - sc.width));
        const float gz = fmaxf(0.f, fmaxf(-pz, pz - sc.width));
        if (fmaf(gz, gz, fmaf(gy, gy, gx * gx)) >= r * r) continue;
      }
      const float tpx = dsubf(tx, o0), tpy = dsubf(ty, o1), tpz = dsubf(tz, o2);
      const float lim = dbl ? r2e : thg2;
      const float4 *const F = A.frames + (size_t)I.sframe;
      const float4 *const B = A.boxes + 2 * (size_t)A.cell_box_first[I.scell];
      for (int base = 0; base < sc.count; base += 32) {
        /* 4-octet box cull of this chunk (lanes 0-3) */
        bool oacc = false;
        if (lane < 4 && base + 8 * lane < sc.count) {
          const float4 lo = B[2 * ((base >> 3) + lane)], hi = B[2 * ((base >> 3) + lane) + 1];
          const float gx = fmaxf(0.f, fmaxf(lo.x - px, px - hi.x));
          const float gy = fmaxf(0.f, fmaxf(lo.y - py, py - hi.y));
          const float gz = fmaxf(0.f, fmaxf(lo.z - pz, pz - hi.z));
          oacc = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < rc * rc;
        }
        const unsigned om = __ballot_sync(FULL_MASK, oacc);
        if (!om) continue;
        const int s = base + lane;
        bool cand = false;
        float dx = 0.f, dy = 0.f, dz = 0.f, r2 = 0.f;
        if (s < sc.count && ((om >> (lane >> 3)) & 1u)) {
          const float4 f = F[s];
          dx = __fsub_rn(tpx, f.x);
          dy = __fsub_rn(tpy, f.y);
          dz = __fsub_rn(tpz, f.z);
          r2 = r2_exact(dx, dy, dz);
          cand = r2 < lim;
          ntests++;
        }
        if (!__any_sync(FULL_MASK, cand)) continue;
        if (cand) {
          const int gi = sc.first + s;
          double Xx = 0., Xy = 0., Xz = 0.;
          if (dbl) {
            Xx = A.xs0[gi]; Xy = A.xs1[gi]; Xz = A.xs2[gi];
            dx = dsubf(__dsub_rn(tx, s0), Xx);
            dy = dsubf(__dsub_rn(ty, s1), Xy);
            dz = dsubf(__dsub_rn(tz, s2), Xz);
            r2 = r2_exact(dx, dy, dz);
          }
          bool hit = (gi != ti) && (r2 < thg2);
          if (hit && !nokey && !(r2 < tsure2 && thg <= hcap)) {
            if (!dbl) {
              Xx = A.xs0[gi]; Xy = A.xs1[gi]; Xz = A.xs2[gi];
            }
            SlowArgs SA;
            SA.items = A.items; SA.cells = A.cells; SA.ext = A.ext;
            SA.dim[0] = A.dim[0]; SA.dim[1] = A.dim[1]; SA.dim[2] = A.dim[2];
            hit = exact_type1(SA, item, tx, ty, tz, thg, Xx, Xy, Xz);
          }
          if (hit) {
            const float4 q = A.mv[gi];
            iact_density(acc, r2, dx, dy, dz, th_inv, tmv.y, tmv.z, tmv.w, q.x, q.y, q.z, q.w);
            nhit++;
          }
        }
      }
    }
    /* warp reduction, plain stores: this warp is the only writer of the target in this launch */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.rho += __shfl_xor_sync(FULL_MASK, acc.rho, o);
      acc.rho_dh += __shfl_xor_sync(FULL_MASK, acc.rho_dh, o);
      acc.wcount += __shfl_xor_sync(FULL_MASK, acc.wcount, o);
      acc.wcount_dh += __shfl_xor_sync(FULL_MASK, acc.wcount_dh, o);
      acc.div_v += __shfl_xor_sync(FULL_MASK, acc.div_v, o);
      acc.rot[0] += __shfl_xor_sync(FULL_MASK, acc.rot[0], o);
      acc.rot[1] += __shfl_xor_sync(FULL_MASK, acc.rot[1], o);
      acc.rot[2] += __shfl_xor_sync(FULL_MASK, acc.rot[2], o);
      nhit += __shfl_xor_sync(FULL_MASK, nhit, o);
      ntests += __shfl_xor_sync(FULL_MASK, ntests, o);
    }
    if (lane == 0) {
      /* += : in a multi-level tree a main-loop target can also be served by another group */
      float4 a = A.dA[ti], b = A.dB[ti];
      a.x += acc.rho; a.y += acc.rho_dh; a.z += acc.wcount; a.w += acc.wcount_dh;
      b.x += acc.div_v; b.y += acc.rot[0]; b.z += acc.rot[1]; b.w += acc.rot[2];
      A.dA[ti] = a;
      A.dB[ti] = b;
      A.count[ti] += nhit;
      if (nhit) atomicAdd(A.total, (unsigned long long)nhit);
      if (ntests) atomicAdd(A.tests, (unsigned long long)ntests);
    }
  }
}

}  // namespace swiftgpu
#endif
