/*
 * kernels_records.cuh - the streaming kernels that keep the device records of libswiftgpu (included
 * by swiftgpu.cu after the handle is defined): AoS <-> SoA transposes, Morton order inside the leaves,
 * the 13-axis sort and its key extrema, tile records, frame arrays and octet boxes of the frame
 * pipeline, target lists and the TaskRecs of a launch. All HBM-bound, together a few per cent of a step.
 */
#ifndef SWIFTGPU_KERNELS_RECORDS_CUH
#define SWIFTGPU_KERNELS_RECORDS_CUH

/* ======================================================================== */
/* Kernels: AoS <-> SoA                                                      */
/* ======================================================================== */
struct DevLayout {
  swiftgpu_part_layout L;
  int scheme;
};

template <typename T>
__device__ __forceinline__ T rd(const char *p, int off) {
  return *(const T *)(p + off);
}
template <typename T>
__device__ __forceinline__ void wr(char *p, int off, T v) {
  *(T *)(p + off) = v;
}

struct Soa {
  double *x;
  float4 *mv, *dA, *dB, *fq1, *fq2, *fq3, *fo1;
  float *h, *u, *rho, *f_hdt, *f_vsig, *g_vsig, *g_lap, *g_amax, *alpha, *alpha_diff, *div_v_prev,
      *div_v_dt, *div_v;
  int8_t *time_bin, *depth_h;
  int32_t *f_minngb;
};

__global__ void k_iota2(int32_t *a, int32_t *b, int64_t n) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  a[p] = (int32_t)p;
  b[p] = (int32_t)p;
}

/* Morton order inside every leaf (one CTA per leaf, bitonic sort of
 * (15-bit Morton code, index) in shared memory). Leaves above LEAF_SORT_MAX
 * particles keep the host order. */
#define LEAF_SORT_MAX 1024
__device__ __forceinline__ uint32_t spread5(uint32_t v) {
  /* 5 bits -> every third bit */
  v = (v | (v << 8)) & 0x0000100fu;
  v = (v | (v << 4)) & 0x000010c3u;
  v = (v | (v << 2)) & 0x00001249u;
  return v;
}
__global__ void __launch_bounds__(128)
    k_leaf_order(const LeafRec *leaves, int nleaves, const char *aos, int part_size, int x_off,
                 int32_t *d2h, int32_t *h2d) {
  __shared__ uint32_t skey[LEAF_SORT_MAX];
  const int l = blockIdx.x;
  if (l >= nleaves) return;
  const LeafRec R = leaves[l];
  const int n = R.count;
  if (n <= 1 || n > LEAF_SORT_MAX) return;
  int N = 1;
  while (N < n) N <<= 1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    uint32_t key = 0xffffffffu;
    if (i < n) {
      const char *b = aos + (size_t)part_size * (size_t)(R.first + i);
      uint32_t q[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float f = (float)(*(const double *)(b + x_off + 8 * k) - R.loc[k]) * R.iwidth[k];
        q[k] = (uint32_t)min(31, max(0, (int)f));
      }
      const uint32_t mort = (spread5(q[0]) << 2) | (spread5(q[1]) << 1) | spread5(q[2]);
      key = (mort << 10) | (uint32_t)i;
    }
    skey[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= N; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const uint32_t a = skey[i], c = skey[j];
          const bool up = ((i & size) == 0);
          if ((a > c) == up) {
            skey[i] = c;
            skey[j] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int i = (int)(skey[r] & 1023u);
    d2h[R.first + r] = R.first + i;
    h2d[R.first + i] = R.first + r;
  }
}

__global__ void k_scatter_i32(const int32_t *src, const int32_t *d2h, int64_t n, int32_t *dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  dst[d2h[p]] = src[p];
}

__global__ void k_aos_to_soa(const char *aos, DevLayout D, Soa S, int64_t n, const int32_t *d2h) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const swiftgpu_part_layout &L = D.L;
  const char *b = aos + (size_t)L.size * (size_t)d2h[p];
  S.x[3 * p + 0] = rd<double>(b, L.x);
  S.x[3 * p + 1] = rd<double>(b, L.x + 8);
  S.x[3 * p + 2] = rd<double>(b, L.x + 16);
  const float m = rd<float>(b, L.mass);
  S.mv[p] = make_float4(m, rd<float>(b, L.v), rd<float>(b, L.v + 4), rd<float>(b, L.v + 8));
  const float h = rd<float>(b, L.h);
  S.h[p] = h;
  const float u = rd<float>(b, D.scheme == SCH_GADGET2 ? L.entropy : L.u);
  S.u[p] = u;
  const float rho = rd<float>(b, L.rho);
  S.rho[p] = rho;
  const int8_t tb = rd<int8_t>(b, L.time_bin);
  S.time_bin[p] = tb;
  S.depth_h[p] = rd<int8_t>(b, L.depth_h);
  /* The density/force union holds the force members of the last step the
   * particle was active in: they are what inactive neighbours contribute. */
  const float P = rd<float>(b, D.scheme == SCH_GADGET2 ? L.P_over_rho2 : L.pressure);
  S.fq1[p] = make_float4(rho, P, rd<float>(b, L.f), rd<float>(b, L.soundspeed));
  /* the force loop's test reads the exact h^2 gamma^2 of a source from the spare lane of its payload:
   * fq2.z (Minimal, Gadget2: u is not read by their force interaction) or fq3.z (SPHENIX) */
  S.fq2[p] = make_float4(rd<float>(b, L.balsara), h, D.scheme == SCH_SPHENIX ? u : hg2_exact(h),
                         __int_as_float((int)tb));
  S.f_hdt[p] = rd<float>(b, L.h_dt);
  S.f_vsig[p] = rd<float>(b, L.v_sig);
  S.f_minngb[p] = rd<int8_t>(b, L.min_ngb_time_bin);
  S.fo1[p] = make_float4(rd<float>(b, L.a_hydro), rd<float>(b, L.a_hydro + 4),
                         rd<float>(b, L.a_hydro + 8),
                         rd<float>(b, D.scheme == SCH_GADGET2 ? L.entropy_dt : L.u_dt));
  if (D.scheme == SCH_SPHENIX) {
    const float al = rd<float>(b, L.visc_alpha), ad = rd<float>(b, L.diff_alpha);
    S.alpha[p] = al;
    S.alpha_diff[p] = ad;
    S.fq3[p] = make_float4(al, ad, hg2_exact(h), 0.f);
    S.div_v_prev[p] = rd<float>(b, L.div_v_previous_step);
    S.div_v_dt[p] = rd<float>(b, L.div_v_dt);
    S.div_v[p] = rd<float>(b, L.div_v);
    S.g_vsig[p] = rd<float>(b, L.v_sig);
    S.g_lap[p] = rd<float>(b, L.laplace_u);
    S.g_amax[p] = rd<float>(b, L.alpha_visc_max_ngb);
  }
}

/* Writes back the fields the phases run so far have made valid, for ACTIVE
 * particles only (inactive particles are read-only on this path). */
__global__ void k_soa_to_aos(char *aos, DevLayout D, Soa S, int64_t n, int max_active_bin,
                             int density_only, const int32_t *d2h) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (S.time_bin[p] > max_active_bin) return;
  const swiftgpu_part_layout &L = D.L;
  char *b = aos + (size_t)L.size * (size_t)d2h[p];
  wr<float>(b, L.h, S.h[p]);
  wr<int8_t>(b, L.depth_h, S.depth_h[p]);
  const float4 a = S.dA[p], c = S.dB[p];
  if (density_only) {
    wr<float>(b, L.rho, a.x);
    wr<float>(b, L.rho_dh, a.y);
    wr<float>(b, L.wcount, a.z);
    wr<float>(b, L.wcount_dh, a.w);
    wr<float>(b, L.div_v, c.x);
    wr<float>(b, L.rot_v, c.y);
    wr<float>(b, L.rot_v + 4, c.z);
    wr<float>(b, L.rot_v + 8, c.w);
    return;
  }
  const float4 q1 = S.fq1[p], q2 = S.fq2[p], o = S.fo1[p];
  wr<float>(b, L.rho, q1.x);
  wr<float>(b, D.scheme == SCH_GADGET2 ? L.P_over_rho2 : L.pressure, q1.y);
  wr<float>(b, L.f, q1.z);
  wr<float>(b, L.soundspeed, q1.w);
  wr<float>(b, L.balsara, q2.x);
  wr<float>(b, L.h_dt, S.f_hdt[p]);
  wr<float>(b, L.a_hydro, o.x);
  wr<float>(b, L.a_hydro + 4, o.y);
  wr<float>(b, L.a_hydro + 8, o.z);
  wr<float>(b, D.scheme == SCH_GADGET2 ? L.entropy_dt : L.u_dt, o.w);
  wr<int8_t>(b, L.min_ngb_time_bin, (int8_t)S.f_minngb[p]);
  if (D.scheme == SCH_SPHENIX) {
    wr<float>(b, L.div_v, S.div_v[p]);
    wr<float>(b, L.v_sig, S.g_vsig[p]);
    wr<float>(b, L.laplace_u, S.g_lap[p]);
    wr<float>(b, L.alpha_visc_max_ngb, S.g_amax[p]);
    wr<float>(b, L.visc_alpha, S.alpha[p]);
    wr<float>(b, L.diff_alpha, S.alpha_diff[p]);
    wr<float>(b, L.div_v_previous_step, S.div_v_prev[p]);
    wr<float>(b, L.div_v_dt, S.div_v_dt[p]);
  } else {
    wr<float>(b, L.v_sig, S.f_vsig[p]);
  }
}

/* ======================================================================== */
/* Kernel: 13-axis sort (runner_do_hydro_sort, runner_sort.c:203)            */
/* One CTA per (cell, sid) segment. Keys are (float)(x . runner_shift[sid])  */
/* of absolute double positions (:411-413); the order of equal keys is       */
/* irrelevant to the neighbour sets. All-ascending bitonic network with      */
/* virtual +inf padding.                                                     */
/* ======================================================================== */
#define SORT_SMEM_MAX 2048
__global__ void __launch_bounds__(256)
    k_sort(const SortSeg *segs, const DevCell *cells, const double *x, uint32_t *sort_idx,
           float *gkeys /* scratch for segments larger than SORT_SMEM_MAX, may be null */) {
  __shared__ float skey[SORT_SMEM_MAX];
  __shared__ uint32_t sidx[SORT_SMEM_MAX];
  const SortSeg seg = segs[blockIdx.x];
  const DevCell c = cells[seg.cell];
  const int n = c.count;
  uint32_t *out = sort_idx + seg.off;
  int N = 1;
  while (N < n) N <<= 1;
  const bool in_smem = n <= SORT_SMEM_MAX;
  float *keys = in_smem ? skey : gkeys + seg.off;
  uint32_t *idx = in_smem ? sidx : out;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t p = (size_t)c.first + i;
    keys[i] = sort_key(x[3 * p], x[3 * p + 1], x[3 * p + 2], seg.sid);
    idx[i] = (uint32_t)i;
  }
  __syncthreads();
  for (int size = 2; size <= N; size <<= 1) {
    for (int stride = size >> 1, first = 1; stride > 0; stride >>= 1, first = 0) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int j = first ? (i ^ (size - 1)) : (i ^ stride);
        if (j > i && j < n) {
          const float ki = keys[i], kj = keys[j];
          if (kj < ki) {
            keys[i] = kj;
            keys[j] = ki;
            const uint32_t t = idx[i];
            idx[i] = idx[j];
            idx[j] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sidx[i];
}

/* ======================================================================== */
/* Kernel: key extrema. The loops only need sort[0].d and sort[count-1].d of  */
/* every (cell, sid) array (dj_min / di_max, functions_hydro.h:1286,1418): the */
/* candidate culling is done with boxes, not with the sorted order. One warp  */
/* per cell computes the extrema of all its requested sids in one pass.       */
/* ======================================================================== */
__global__ void __launch_bounds__(128)
    k_extrema(const int32_t *ext_cells, int ncells, const DevCell *cells, const double *x, float2 *ext) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= ncells) return;
  const DevCell c = cells[ext_cells[w]];
  const unsigned mask = c.sort_mask;
  float mn[13], mx[13];
#pragma unroll
  for (int s = 0; s < 13; s++) {
    mn[s] = 3.402823466e+38f;
    mx[s] = -3.402823466e+38f;
  }
  for (int k = lane; k < c.count; k += 32) {
    const size_t p = (size_t)c.first + k;
    const double px = x[3 * p], py = x[3 * p + 1], pz = x[3 * p + 2];
#pragma unroll
    for (int s = 0; s < 13; s++) {
      if ((mask >> s) & 1u) {
        const float key = sort_key(px, py, pz, s);
        mn[s] = fminf(mn[s], key);
        mx[s] = fmaxf(mx[s], key);
      }
    }
  }
  int rank = 0;
#pragma unroll
  for (int s = 0; s < 13; s++) {
    if ((mask >> s) & 1u) {
      const float a = warp_min(mn[s]), b = warp_max(mx[s]);
      if (lane == 0) ext[c.seg_base + rank] = make_float2(a, b);
      rank++;
    }
  }
}

/* ======================================================================== */
/* Kernels: the TMA-copyable source records of the tile pipeline             */
/* (loops_tile.cuh). xf/x4 follow the positions (once per upload / xv halo), */
/* xf.w follows h (again before the force loop), the octet boxes follow xf.  */
/* ======================================================================== */
__device__ __forceinline__ float reach2(float h, float margin) {
  const float re = fmaf(__fmul_rn(h, KERNEL_GAMMA), PREFILTER_REL, margin);
  return re * re;
}
__global__ void k_prep_tiles(const double *x, const float *h, int64_t n, float margin, float4 *xf,
                             double *xs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double px = x[3 * p], py = x[3 * p + 1], pz = x[3 * p + 2];
  xf[p] = make_float4(__double2float_rn(px), __double2float_rn(py), __double2float_rn(pz),
                      reach2(h[p], margin));
  xs[p] = px;
  xs[(n + 4) + p] = py;
  xs[2 * (n + 4) + p] = pz;
}
__global__ void k_refresh_reach(const float *h, int64_t n, float margin, float4 *xf) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  xf[p].w = reach2(h[p], margin);
}
/* gradient payload of every particle: (u, rho, soundspeed, alpha_visc) */
__global__ void k_prep_gq(const float4 *fq1, const float4 *fq2, const float4 *fq3, int64_t n, float4 *gq) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float4 q1 = fq1[p];
  gq[p] = make_float4(fq2[p].z, q1.x, q1.w, fq3[p].x);
}
/* One warp per cell (every level): the axis-aligned box of each octet of 8
 * consecutive particles of the cell, in absolute floats. */
__global__ void __launch_bounds__(128)
    k_octet_boxes(const DevCell *cells, int ncells, const int32_t *box_first, const float4 *xf,
                  float4 *boxes) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= ncells) return;
  const int first = cells[c].first, count = cells[c].count;
  const int noct = (count + 7) >> 3;
  float4 *out = boxes + 2 * (size_t)box_first[c];
  for (int o = lane; o < noct; o += 32) {
    float lo0 = 3.0e30f, lo1 = 3.0e30f, lo2 = 3.0e30f, hi0 = -3.0e30f, hi1 = -3.0e30f, hi2 = -3.0e30f;
    const int k1 = min(count, 8 * o + 8);
    for (int k = 8 * o; k < k1; k++) {
      const float4 f = xf[first + k];
      lo0 = fminf(lo0, f.x); lo1 = fminf(lo1, f.y); lo2 = fminf(lo2, f.z);
      hi0 = fmaxf(hi0, f.x); hi1 = fmaxf(hi1, f.y); hi2 = fmaxf(hi2, f.z);
    }
    out[2 * o] = make_float4(lo0, lo1, lo2, 0.f);
    out[2 * o + 1] = make_float4(hi0, hi1, hi2, 0.f);
  }
}

/* ======================================================================== */
/* Kernels of the frame pipeline (loops_pipe.cuh)                            */
/* ======================================================================== */
/* One warp per frame: F[k] = (float)(x_k - origin) for the particles of the
 * frame's cell - the reference's pix / pjx of functions_hydro.h:1327-1338,
 * evaluated once per step instead of once per candidate pair. */
__global__ void __launch_bounds__(128)
    k_frames(const swiftgpu_handle::FrameRec *recs, int64_t nframes, const double *xs0, const double *xs1,
             const double *xs2, float4 *frames) {
  const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= nframes) return;
  const swiftgpu_handle::FrameRec R = recs[f];
  float4 *out = frames + R.off;
  for (int k = lane; k < R.count; k += 32) {
    const size_t p = (size_t)R.first + k;
    out[k] = make_float4(dsubf(xs0[p], R.o[0]), dsubf(xs1[p], R.o[1]), dsubf(xs2[p], R.o[2]), 0.f);
  }
}
/* Octet boxes in the cell's OWN frame (float)(x - loc): what the culls of the
 * frame pipeline compare, shifted by the per-item float offset `d`. */
__global__ void __launch_bounds__(128)
    k_octet_boxes_own(const DevCell *cells, int ncells, const int32_t *box_first, const double *xs0,
                      const double *xs1, const double *xs2, float4 *boxes) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= ncells) return;
  const DevCell C = cells[c];
  const int first = C.first, count = C.count;
  const int noct = (count + 7) >> 3;
  float4 *out = boxes + 2 * (size_t)box_first[c];
  for (int o = lane; o < noct; o += 32) {
    float lo0 = 3.0e30f, lo1 = 3.0e30f, lo2 = 3.0e30f, hi0 = -3.0e30f, hi1 = -3.0e30f, hi2 = -3.0e30f;
    const int k1 = min(count, 8 * o + 8);
    for (int k = 8 * o; k < k1; k++) {
      const size_t p = (size_t)first + k;
      const float fx = dsubf(xs0[p], C.loc[0]), fy = dsubf(xs1[p], C.loc[1]), fz = dsubf(xs2[p], C.loc[2]);
      lo0 = fminf(lo0, fx); lo1 = fminf(lo1, fy); lo2 = fminf(lo2, fz);
      hi0 = fmaxf(hi0, fx); hi1 = fmaxf(hi1, fy); hi2 = fmaxf(hi2, fz);
    }
    out[2 * o] = make_float4(lo0, lo1, lo2, 0.f);
    out[2 * o + 1] = make_float4(hi0, hi1, hi2, 0.f);
  }
}
/* One warp per (group, 64-target chunk) of the host task list: the TaskRec of
 * every NON-EMPTY task, compacted (the order of the heaviest-first list is kept
 * up to the scheduling of the warps). */
__global__ void __launch_bounds__(128)
    k_task_recs(const int32_t *task_group, const int32_t *task_chunk, int ntasks, const Group *groups,
                const DevCell *cells, const int32_t *tgt_first, const int32_t *tgt_count,
                const int32_t *tgt_list, const double *xs0, const double *xs1, const double *xs2,
                const float *h, TaskRec *recs, unsigned int *ntask_dev, const unsigned long long *gate,
                unsigned long long gate_lo, unsigned long long gate_hi, int chunk /* targets per task */,
                const unsigned long long *gate_den) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= ntasks) return;
  /* ghost re-runs: this launch only if the number of unconverged particles (a device counter the
   * host never reads in between) is in [gate_lo, gate_hi) - else the other kernel takes the pass */
  if (gate) {
    /* with gate_den the bounds are per unit of *gate_den (targets per 64-target chunk of the list) */
    const unsigned long long v = *gate, den = gate_den ? *gate_den : 1ull;
    if (v < gate_lo * den) return;
    if (gate_hi != ~0ull && v >= gate_hi * den) return;
  }
  const int g = task_group[w];
  const int nt = tgt_count[g];
  /* the host's task list is cut for the smallest task size; a launch with larger tasks uses its head */
  const int t0 = task_chunk[w] * chunk;
  if (t0 >= nt) return;
  const int n = min(chunk, nt - t0);
  const Group G = groups[g];
  const DevCell C = cells[G.tcell];
  const int off = tgt_first[g] + t0;
  float lo[3] = {3.0e30f, 3.0e30f, 3.0e30f}, hi[3] = {-3.0e30f, -3.0e30f, -3.0e30f}, rmax = 0.f;
  for (int k = lane; k < n; k += 32) {
    const int ti = tgt_list[off + k];
    const float f[3] = {dsubf(xs0[ti], C.loc[0]), dsubf(xs1[ti], C.loc[1]), dsubf(xs2[ti], C.loc[2])};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(lo[a], f[a]);
      hi[a] = fmaxf(hi[a], f[a]);
    }
    rmax = fmaxf(rmax, __fmul_rn(h[ti], KERNEL_GAMMA));
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    lo[a] = warp_min(lo[a]);
    hi[a] = warp_max(hi[a]);
  }
  rmax = warp_max(rmax);
  if (lane == 0) {
    TaskRec R;
    R.item_first = G.item_first;
    R.item_count = G.item_count;
    R.tgt_off = off;
    R.ntgt = n;
    R.tcell = G.tcell;
    for (int a = 0; a < 3; a++) {
      R.lo[a] = lo[a];
      R.hi[a] = hi[a];
    }
    R.rmax = rmax;
    recs[atomicAdd(ntask_dev, 1u)] = R;
  }
}

/* ======================================================================== */
/* Kernel: target lists (active particles of each group's cell)              */
/* ======================================================================== */
__global__ void k_build_targets(const Group *groups, int ngroups, const DevCell *cells,
                                const int8_t *time_bin, int max_active_bin, const int32_t *tgt_first,
                                int32_t *tgt_count, int32_t *tgt_list, const Item *items,
                                const int8_t *depth_h, unsigned long long *totals) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  const Group G = groups[g];
  const DevCell c = cells[G.tcell];
  /* A particle takes part in an item only if its depth_h lies in the item's
   * range (limit_min_h / limit_max_h of the reference's recursion): targets
   * outside the union of the group's ranges have nothing to do here. In a
   * multi-level tree most particles of a non-leaf target cell are such. */
  int lo = 127, hi = 0;
  for (int k = lane; k < G.item_count; k += 32) {
    const Item I = items[G.item_first + k];
    lo = min(lo, (int)I.min_depth);
    hi = max(hi, (int)I.max_depth);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
  }
  int32_t *out = tgt_list + tgt_first[g];
  int nout = 0;
  for (int base = 0; base < c.count; base += 32) {
    const int k = base + lane;
    bool act = (k < c.count) && (time_bin[c.first + k] <= max_active_bin);
    if (act) {
      const int d = depth_h[c.first + k];
      act = d >= lo && d <= hi;
    }
    const unsigned m = __ballot_sync(FULL_MASK, act);
    if (act) out[nout + __popc(m & ((1u << lane) - 1u))] = c.first + k;
    nout += __popc(m);
  }
  if (lane == 0) {
    tgt_count[g] = nout;
    if (nout) { /* totals[0] targets, totals[1] non-empty 64-target tasks: the launch picks the CTA size */
      atomicAdd(totals, (unsigned long long)nout);
      atomicAdd(totals + 1, (unsigned long long)((nout + TASK_TARGETS - 1) / TASK_TARGETS));
    }
  }
}


#endif
