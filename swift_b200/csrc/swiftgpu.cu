/*
 * swiftgpu.cu - libswiftgpu: C ABI (include/swiftgpu.h), device state and the
 * per-particle kernels (device order, AoS<->SoA transposes, tile records, 13-axis
 * sort, ghost, extra ghost, end force + CFL time-step). The neighbour loops live
 * in loops_tile.cuh (default), loops_cta.cuh and loops.cuh (A/B generations).
 *
 * Device state is SoA. Every cell is a contiguous index range exactly as on the
 * host (progeny partition their parent's range); inside a leaf the particles are
 * in Morton order (d2h / h2d map device <-> host indices at the AoS boundary):
 *   x[3n] f64 | mv[n] f32x4 (m,vx,vy,vz) | h,u,rho f32 | time_bin,depth_h i8
 *   tile records xf (float position, reach^2), xs (3 f64 columns), octet boxes
 *   density sums dA (rho,rho_dh,wcount,wcount_dh), dB (div_v,rot_v)
 *   loop inputs  fq1 (rho,P,f,cs), fq2 (balsara,h,u,time_bin), fq3 (alpha,alpha_diff)
 *   force sums   fo1 (a_hydro,u_dt|entropy_dt), h_dt, v_sig, min_ngb_time_bin
 * No CPU fallback exists: every entry point runs CUDA kernels or fails.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swiftgpu.h"
#include "loops_pipe.cuh"
#ifdef SWIFTGPU_LEGACY_LOOPS /* `make legacy`: the three superseded generations, for A/B runs (SWIFTGPU_LOOPS=tile|cta|warp) */
#include "legacy/loops_cta.cuh"
#include "legacy/loops_tile.cuh"
#endif
#include "loops_direct.cuh"

using namespace swiftgpu;

#include "halo.cuh"

#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess)                                                         \
      return h->fail("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

/* ======================================================================== */
/* Device-side worklist of one loop family                                   */
/* ======================================================================== */
struct DevList {
  Item *items = nullptr;
  Group *groups = nullptr;
  int32_t *task_group = nullptr, *task_chunk = nullptr;
  int32_t *tgt_first = nullptr, *tgt_count = nullptr, *tgt_list = nullptr;
  TaskRec *task_recs = nullptr; /* compacted per launch by k_task_recs */
  int ngroups = 0, ntasks = 0;
  size_t nitems = 0;
  int64_t tgt_total = 0;
  void release() {
    cudaFree(items); cudaFree(groups); cudaFree(task_group); cudaFree(task_chunk);
    cudaFree(tgt_first); cudaFree(tgt_count); cudaFree(tgt_list); cudaFree(task_recs);
    task_recs = nullptr;
    items = nullptr; groups = nullptr; task_group = task_chunk = nullptr;
    tgt_first = tgt_count = tgt_list = nullptr;
    ngroups = ntasks = 0; nitems = 0; tgt_total = 0;
  }
};

struct LeafRec {
  double loc[3];
  float iwidth[3]; /* 32 / width */
  int32_t first, count;
};

struct SortSeg {
  int32_t cell;
  int32_t sid;
  int64_t off;
};

struct swiftgpu_handle {
  swiftgpu_config cfg;
  swiftgpu_step step;
  std::string err;
  int has_step = 0;

  std::vector<swiftgpu_cell> cells;
  std::vector<float> up_hmax, up_hmax_active; /* values as uploaded */
  float cells_uploaded_hmax(int c) const { return up_hmax[c]; }
  float cells_uploaded_hmax_active(int c) const { return up_hmax_active[c]; }
  std::vector<int32_t> top;
  std::vector<uint8_t> force_bits; /* recursion predicate bits the force list was built with */
  std::vector<uint8_t> loop1_bits; /* ... and the density list (cell.h:951,992), revalidated for the gradient loop */
  int64_t n = 0;
  int64_t n_host = 0; /* particles that cross the host boundary: [0, n_host) (all, or the local ones) */
  int ncells = 0;

  /* device */
  DevCell *d_cells = nullptr;
  DevCell *d_cells_init = nullptr; /* pristine copy: the ghost raises h_max in d_cells */
  float *d_dmin = nullptr, *d_dxp = nullptr;
  std::vector<uint64_t> req_density, req_subset, req_gradient, req_force; /* sort requests of the lists */
  float *d_dxp_old = nullptr, *d_hmax_tmp = nullptr;
  bool ext_valid = false; /* d_ext holds the key extrema of the current segments */
  uint8_t *d_loop1_bits = nullptr, *d_grad_bits = nullptr;
  bool gradient_own = false; /* L_gradient was rebuilt after the ghost (else the density list is used) */
  char *d_aos = nullptr;
  char *d_xaos = nullptr; /* device copy of the caller's struct xpart[] (drift) */
  int32_t *d_wakeup = nullptr; /* limiter loop: limiter_data.wakeup per particle, device order */
  swiftgpu_xpart_layout xlayout = {0, 0, 0, 0, -1};
  int64_t n_x = 0;
  size_t aos_bytes = 0;
  double *x = nullptr;
  float4 *mv = nullptr, *dA = nullptr, *dB = nullptr, *fq1 = nullptr, *fq2 = nullptr, *fq3 = nullptr,
         *fo1 = nullptr;
  float *hh = nullptr, *u = nullptr, *rho = nullptr, *f_hdt = nullptr, *f_vsig = nullptr,
        *g_vsig = nullptr, *g_lap = nullptr, *g_amax = nullptr, *alpha = nullptr,
        *alpha_diff = nullptr, *div_v_prev = nullptr, *div_v_dt = nullptr, *div_v = nullptr,
        *gleft = nullptr, *gright = nullptr;
  int8_t *time_bin = nullptr, *depth_h = nullptr;
  int32_t *f_minngb = nullptr, *nd = nullptr, *ng = nullptr, *nf = nullptr;
  /* device order: particles of every LEAF are kept in Morton order of their
   * position inside the leaf (compact octets for the box culling of the
   * loops); d2h[device index] = host index, h2d the inverse. Cells stay
   * contiguous ranges, so nothing but the AoS boundary sees the permutation. */
  int32_t *d_d2h = nullptr, *d_h2d = nullptr;
  int32_t *d_cnt_tmp = nullptr; /* download_counts scratch */
  float *dt_cfl = nullptr;      /* hydro_compute_timestep of the active particles (end_force epilogue) */
  struct LeafRec *d_leaves = nullptr;
  int nleaves = 0;
  bool leaves_valid = false;
  bool perm_stale = false; /* cells changed after the particles were transposed */
  /* frame arrays (loops_pipe.cuh): one per (cell, origin) the items of the lists read */
  struct FrameRec {
    int32_t first, count;
    double o[3];
    uint32_t off;
    uint32_t pad;
  };
  std::vector<FrameRec> frames_host;
  std::vector<int32_t> frame_idx; /* [cell * 14 + slot] -> index into frames_host (slot 0: own frame, 1 + sid: ci frames) */
  float4 *d_frames = nullptr;
  int2 *d_flat_tgt = nullptr; /* (target, group) list of a sparse launch (loops_direct.cuh) */
  FrameRec *d_frame_recs = nullptr;
  uint64_t frames_total = 0, frames_cap = 0;
  size_t frame_recs_uploaded = 0;
  bool frames_valid = false; /* d_frames holds the frames of the current positions */
  /* tile pipeline records (loops_tile.cuh) */
  float4 *xf = nullptr, *gq = nullptr, *boxes = nullptr;
  double *xs = nullptr; /* 3 columns of n + 4 doubles */
  int32_t *d_box_first = nullptr;
  int64_t nboxes = 0;
  uint32_t *sort_idx = nullptr;
  float *d_sort_keys = nullptr; /* key scratch, only when a requested segment exceeds SORT_SMEM_MAX */
  int64_t sort_total = 0;
  SortSeg *d_segs = nullptr;
  int nsegs = 0;
  float2 *d_ext = nullptr;        /* key extrema per segment */
  int32_t *d_ext_cells = nullptr; /* cells that own at least one segment */
  int n_ext_cells = 0;
  bool full_sorted = false;       /* sort_idx holds the full sorted arrays */
  unsigned long long *d_counters = nullptr; /* [0] density [1] gradient [2] force [3] redo */
  int32_t *d_flag = nullptr;
  uint8_t *d_force_bits = nullptr;

  DevList L_density, L_subset, L_force, L_gradient;

  /* multi-GPU */
  std::vector<HaloPeer> halo;
  sg_ncclComm_t comm = nullptr;
  bool halo_ready = false;
  int64_t halo_bytes[3] = {0, 0, 0}; /* bytes sent per exchange of each phase */
  bool lists_built = false;
  bool sorted = false;
  uint32_t phases_done = 0;

  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pev[7][2] = {}; /* per phase: begin / end of its last run */
  uint32_t pev_pending = 0;   /* phases whose time has not been read back yet */
  int cur_phase = -1;
  bool counters_pending = false;
  swiftgpu_stats stats;

  int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return 1;
  }
};
typedef swiftgpu_handle H;

static thread_local std::string g_err;

#include "kernels_records.cuh"
#include "kernels_finalise.cuh"
#include "kernels_drift.cuh"

/* ======================================================================== */
/* Host side                                                                 */
/* ======================================================================== */
static Soa soa_of(H *h) {
  Soa S;
  S.x = h->x; S.mv = h->mv; S.dA = h->dA; S.dB = h->dB; S.fq1 = h->fq1; S.fq2 = h->fq2;
  S.fq3 = h->fq3; S.fo1 = h->fo1; S.h = h->hh; S.u = h->u; S.rho = h->rho; S.f_hdt = h->f_hdt;
  S.f_vsig = h->f_vsig; S.g_vsig = h->g_vsig; S.g_lap = h->g_lap; S.g_amax = h->g_amax;
  S.alpha = h->alpha; S.alpha_diff = h->alpha_diff; S.div_v_prev = h->div_v_prev;
  S.div_v_dt = h->div_v_dt; S.div_v = h->div_v; S.time_bin = h->time_bin; S.depth_h = h->depth_h;
  S.f_minngb = h->f_minngb;
  return S;
}

static const int32_t kLayouts[3][33] = {
    /* offsetof() of the reference's default builds (tests/golden/part_layouts.json) */
    {128, 0, 16, 40, 52, 64, 68, 72, 76, -1, -1, 80, 84, 88, 92, 100, 96, 84, 88, -1, 92, 96, 100,
     104, -1, -1, -1, -1, -1, -1, 113, 114, 116},
    {128, 0, 16, 40, 52, 68, 64, -1, -1, 76, 80, 72, 84, 88, 92, 96, 108, 88, -1, 92, 96, 100, 104,
     84, -1, -1, -1, -1, -1, -1, 113, 114, 116},
    {160, 0, 16, 40, 52, 64, 68, 72, 76, -1, -1, 80, 112, 116, 120, 124, 84, 112, 116, -1, 120, 100,
     124, 128, 88, 92, 96, 104, 108, 132, 138, 137, 140}};

extern "C" int swiftgpu_default_layout(int scheme, swiftgpu_part_layout *out) {
  if (scheme < 0 || scheme > 2 || !out) return 1;
  static_assert(sizeof(swiftgpu_part_layout) == 33 * sizeof(int32_t), "layout struct");
  memcpy(out, kLayouts[scheme], sizeof(*out));
  return 0;
}

extern "C" int swiftgpu_default_config(int scheme, swiftgpu_config *cfg) {
  if (scheme < 0 || scheme > 2 || !cfg) return 1;
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = SWIFTGPU_ABI_VERSION;
  cfg->scheme = scheme;
  cfg->periodic = 1;
  cfg->dim[0] = cfg->dim[1] = cfg->dim[2] = 1.0;
  cfg->eta_neighbours = 1.2348f;   /* hydro_props_init_no_hydro */
  cfg->h_tolerance = 1e-4f;        /* hydro_properties.c:45 */
  cfg->h_max = 3.402823466e+38f;   /* hydro_props_default_h_max = FLT_MAX */
  cfg->h_min = 0.f;
  cfg->max_smoothing_iterations = 30; /* hydro_properties.c:42 */
  cfg->CFL_condition = 0.1f;
  cfg->viscosity_alpha = scheme == SWIFTGPU_SCHEME_SPHENIX ? 0.1f : 0.8f;
  cfg->viscosity_alpha_max = 2.0f;
  cfg->viscosity_alpha_min = 0.0f;
  cfg->viscosity_length = 0.05f;
  cfg->diffusion_alpha = 0.0f;
  cfg->diffusion_beta = 1.0f;
  cfg->diffusion_alpha_max = 1.0f;
  cfg->diffusion_alpha_min = 0.0f;
  cfg->rank = 0;
  cfg->nranks = 1;
  return swiftgpu_default_layout(scheme, &cfg->layout);
}

extern "C" const char *swiftgpu_last_error(const swiftgpu_t *h) {
  return h ? h->err.c_str() : g_err.c_str();
}

extern "C" int swiftgpu_init(swiftgpu_t **out, const swiftgpu_config *cfg) {
  if (!out || !cfg) return 1;
  *out = nullptr;
  if (cfg->abi_version != SWIFTGPU_ABI_VERSION) {
    g_err = "ABI version mismatch";
    return 1;
  }
  if (cfg->scheme < 0 || cfg->scheme > 2) {
    g_err = "unknown scheme";
    return 1;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_err = std::string("no CUDA device: ") + cudaGetErrorString(e) +
            " (libswiftgpu has no CPU fallback)";
    return 2;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_err = "bad device ordinal";
    return 1;
  }
  H *h = new H();
  h->cfg = *cfg;
  memset(&h->stats, 0, sizeof(h->stats));
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaStreamCreate(&h->own_stream) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
      cudaMalloc(&h->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&h->d_flag, sizeof(int32_t)) != cudaSuccess) {
    g_err = std::string("CUDA init failed: ") + cudaGetErrorString(cudaGetLastError());
    delete h;
    return 2;
  }
  h->stream = h->own_stream;
  cudaMemset(h->d_counters, 0, 16 * sizeof(unsigned long long));
  *out = h;
  return 0;
}

static void free_parts(H *h) {
  cudaFree(h->d_wakeup);
  h->d_wakeup = nullptr;
  cudaFree(h->d_xaos);
  h->d_xaos = nullptr;
  h->n_x = 0;
  cudaFree(h->d_aos); cudaFree(h->x); cudaFree(h->mv); cudaFree(h->dA); cudaFree(h->dB);
  cudaFree(h->fq1); cudaFree(h->fq2); cudaFree(h->fq3); cudaFree(h->fo1); cudaFree(h->hh);
  cudaFree(h->u); cudaFree(h->rho); cudaFree(h->f_hdt); cudaFree(h->f_vsig); cudaFree(h->g_vsig);
  cudaFree(h->g_lap); cudaFree(h->g_amax); cudaFree(h->alpha); cudaFree(h->alpha_diff);
  cudaFree(h->div_v_prev); cudaFree(h->div_v_dt); cudaFree(h->div_v); cudaFree(h->gleft);
  cudaFree(h->gright); cudaFree(h->time_bin); cudaFree(h->depth_h); cudaFree(h->f_minngb);
  cudaFree(h->nd); cudaFree(h->ng); cudaFree(h->nf);
  cudaFree(h->xf); cudaFree(h->xs); cudaFree(h->gq);
  cudaFree(h->d_d2h); cudaFree(h->d_h2d); cudaFree(h->d_cnt_tmp); cudaFree(h->dt_cfl);
  h->d_d2h = h->d_h2d = h->d_cnt_tmp = nullptr;
  cudaFree(h->d_flat_tgt);
  h->d_flat_tgt = nullptr;
  h->dt_cfl = nullptr;
  h->xf = h->gq = nullptr; h->xs = nullptr;
  h->d_aos = nullptr; h->x = nullptr;
  h->n = 0;
}

static void halo_release(H *h, bool keep_comm = false) {
  for (HaloPeer &P : h->halo) {
    cudaFree(P.d_send_idx); cudaFree(P.d_recv_idx); cudaFree(P.d_sendbuf); cudaFree(P.d_recvbuf);
  }
  h->halo.clear();
  if (h->comm && !keep_comm) {
    std::string e;
    NcclApi *N = nccl_api(e);
    if (N) N->CommDestroy(h->comm);
    h->comm = nullptr;
  }
  h->halo_ready = false;
}

extern "C" void swiftgpu_destroy(swiftgpu_t *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  free_parts(h);
  h->L_density.release(); h->L_subset.release(); h->L_force.release(); h->L_gradient.release();
  halo_release(h);
  cudaFree(h->d_cells); cudaFree(h->d_cells_init); cudaFree(h->d_dmin); cudaFree(h->d_dxp); cudaFree(h->sort_idx); cudaFree(h->d_sort_keys); cudaFree(h->d_segs); cudaFree(h->d_ext); cudaFree(h->d_ext_cells); cudaFree(h->d_counters);
  cudaFree(h->d_flag); cudaFree(h->d_force_bits); cudaFree(h->d_loop1_bits); cudaFree(h->d_grad_bits);
  cudaFree(h->d_dxp_old); cudaFree(h->d_hmax_tmp);
  cudaFree(h->boxes); cudaFree(h->d_box_first); cudaFree(h->d_leaves);
  cudaFree(h->d_frames); cudaFree(h->d_frame_recs); cudaFree(h->d_flat_tgt);
  for (int ph = 0; ph < 7; ph++)
    for (int k = 0; k < 2; k++)
      if (h->pev[ph][k]) cudaEventDestroy(h->pev[ph][k]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

extern "C" int swiftgpu_set_step(swiftgpu_t *h, const swiftgpu_step *step) {
  if (!h || !step) return 1;
  if (step->with_cosmology) return h->fail("cosmological time integration is not supported");
  /* the cosmological factors (a_factor_Balsara_eps, fac_mu, a2_inv of the Gadget2
   * entropy equation, ...) are not carried through: reject instead of computing
   * silently wrong balsara / alpha values */
  if (step->a != 1.0f || step->H != 0.0f)
    return h->fail("cosmology is not supported: set a = 1, H = 0 (got a = %g, H = %g)", (double)step->a, (double)step->H);
  if (h->has_step && (h->step.ti_current != step->ti_current ||
                      h->step.max_active_bin != step->max_active_bin))
    h->lists_built = false; /* activity changed: the worklists depend on it */
  h->step = *step;
  h->has_step = 1;
  return 0;
}

extern "C" int swiftgpu_set_stream(swiftgpu_t *h, void *cuda_stream) {
  if (!h) return 1;
  cudaSetDevice(h->cfg.device);
  CK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

extern "C" int swiftgpu_upload_cells(swiftgpu_t *h, const swiftgpu_cell *cells, int32_t ncells,
                                     const int32_t *top, int32_t ntop) {
  if (!h || !cells || !top || ncells <= 0 || ntop <= 0) return 1;
  cudaSetDevice(h->cfg.device);
  h->cells.assign(cells, cells + ncells);
  h->up_hmax.resize(ncells);
  h->up_hmax_active.resize(ncells);
  for (int c = 0; c < ncells; c++) {
    h->up_hmax[c] = cells[c].h_max;
    h->up_hmax_active[c] = cells[c].h_max_active;
  }
  h->top.assign(top, top + ntop);
  h->ncells = ncells;
  h->lists_built = false;
  h->sorted = false;
  h->halo_ready = false; /* the send/receive lists follow the cells */
  h->leaves_valid = false;
  if (h->n > 0 && h->x) h->perm_stale = true; /* the device order follows the leaves */
  return 0;
}

template <class T>
static cudaError_t to_device(T **dst, const std::vector<T> &v) {
  cudaFree(*dst);
  *dst = nullptr;
  const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void **)dst, bytes);
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

#include "lists_host.inl" /* frame_of, upload_list, pull_cell_hmax, build_lists, ensure_lists */

static int alloc_parts(H *h, int64_t n) {
  if (h->n == n && h->x) return 0;
  free_parts(h);
  const swiftgpu_part_layout &L = h->cfg.layout;
  h->aos_bytes = (size_t)L.size * n;
#define AL(ptr, T, count) CK(cudaMalloc((void **)&(ptr), sizeof(T) * (size_t)(count)))
  AL(h->d_aos, char, h->aos_bytes);
  AL(h->x, double, 3 * n);
  AL(h->mv, float4, n); AL(h->dA, float4, n); AL(h->dB, float4, n); AL(h->fq1, float4, n);
  AL(h->fq2, float4, n); AL(h->fq3, float4, n); AL(h->fo1, float4, n);
  AL(h->hh, float, n); AL(h->u, float, n); AL(h->rho, float, n); AL(h->f_hdt, float, n);
  AL(h->f_vsig, float, n); AL(h->g_vsig, float, n); AL(h->g_lap, float, n); AL(h->g_amax, float, n);
  AL(h->alpha, float, n); AL(h->alpha_diff, float, n); AL(h->div_v_prev, float, n);
  AL(h->div_v_dt, float, n); AL(h->div_v, float, n); AL(h->gleft, float, n); AL(h->gright, float, n);
  AL(h->time_bin, int8_t, n); AL(h->depth_h, int8_t, n);
  AL(h->f_minngb, int32_t, n); AL(h->nd, int32_t, n); AL(h->ng, int32_t, n); AL(h->nf, int32_t, n);
  AL(h->xf, float4, n + 2); AL(h->xs, double, 3 * (n + 4));
  AL(h->d_d2h, int32_t, n); AL(h->d_h2d, int32_t, n); AL(h->dt_cfl, float, n);
  if (h->cfg.scheme == SCH_SPHENIX) AL(h->gq, float4, n + 2);
#undef AL
  h->n = n;
  h->lists_built = false;
  return 0;
}

static bool use_leaf_order() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SWIFTGPU_NO_REORDER");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

/* Device order of the particles now in d_aos (host order). */
static int build_device_order(H *h) {
  const int64_t n = h->n;
  k_iota2<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_d2h, h->d_h2d, n);
  h->stats.n_launches++;
  if (!use_leaf_order() || h->cells.empty()) return 0;
  if (!h->leaves_valid) {
    std::vector<LeafRec> lv;
    for (const swiftgpu_cell &c : h->cells) {
      if (c.split || c.count <= 1) continue;
      if (c.first_part < 0 || c.first_part + c.count > n) return h->fail("cell range outside the particle array");
      /* proxies keep the order they arrive in: the halo exchange ships every cell in the SENDER's
       * device order (its Morton order), the same in all three phases */
      if (h->cfg.nranks > 1 && c.nodeID != h->cfg.rank) continue;
      LeafRec R;
      for (int k = 0; k < 3; k++) {
        R.loc[k] = c.loc[k];
        R.iwidth[k] = c.width[k] > 0. ? (float)(32. / c.width[k]) : 0.f;
      }
      R.first = (int32_t)c.first_part;
      R.count = c.count;
      lv.push_back(R);
    }
    CK(to_device(&h->d_leaves, lv));
    h->nleaves = (int)lv.size();
    h->leaves_valid = true;
  }
  if (h->nleaves > 0) {
    k_leaf_order<<<h->nleaves, 128, 0, h->stream>>>(h->d_leaves, h->nleaves, h->d_aos, h->cfg.layout.size,
                                                    h->cfg.layout.x, h->d_d2h, h->d_h2d);
    h->stats.n_launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

static int transpose_in(H *h) {
  const int64_t n = h->n;
  DevLayout D;
  D.L = h->cfg.layout;
  D.scheme = h->cfg.scheme;
  if (build_device_order(h)) return 1;
  h->perm_stale = false;
  CK(cudaMemsetAsync(h->nd, 0, sizeof(int32_t) * n, h->stream));
  CK(cudaMemsetAsync(h->ng, 0, sizeof(int32_t) * n, h->stream));
  CK(cudaMemsetAsync(h->nf, 0, sizeof(int32_t) * n, h->stream));
  CK(cudaMemsetAsync(h->dA, 0, sizeof(float4) * n, h->stream));
  CK(cudaMemsetAsync(h->dB, 0, sizeof(float4) * n, h->stream));
  CK(cudaMemsetAsync(h->fq3, 0, sizeof(float4) * n, h->stream));
  const int64_t nh = h->n_host > 0 ? h->n_host : n;
  k_aos_to_soa<<<(unsigned)((nh + 255) / 256), 256, 0, h->stream>>>(h->d_aos, D, soa_of(h), nh, h->d_d2h);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  h->phases_done = 0;
  h->sorted = false;
  /* h may have changed: the cell table keeps the uploaded h_max values; the
   * worklists only depend on cells + step. The ghost state restarts. */
  return 0;
}

extern "C" int swiftgpu_upload_parts(swiftgpu_t *h, const void *parts_aos, int64_t nparts) {
  if (!h || !parts_aos || nparts <= 0) return 1;
  cudaSetDevice(h->cfg.device);
  if (alloc_parts(h, nparts)) return 1;
  h->n_host = nparts;
  CK(cudaMemcpyAsync(h->d_aos, parts_aos, h->aos_bytes, cudaMemcpyHostToDevice, h->stream));
  return transpose_in(h);
}

/* Only the rank's OWN particles cross the host boundary: space->parts holds them in [0, nlocal) and
 * the proxies of foreign cells behind them; the proxies' slots are filled by the xv halo exchange
 * (the reference fills them with recv tasks, scheduler.c:1088-1112), never by the host. */
extern "C" int swiftgpu_upload_parts_local(swiftgpu_t *h, const void *parts_aos, int64_t nlocal, int64_t ntotal) {
  if (!h || !parts_aos || nlocal <= 0 || ntotal < nlocal) return 1;
  cudaSetDevice(h->cfg.device);
  if (ntotal > nlocal && h->cfg.nranks <= 1) return h->fail("upload_parts_local: proxies without ranks");
  for (const swiftgpu_cell &c : h->cells)
    if (c.depth == 0 && c.count > 0 && (c.nodeID == h->cfg.rank) != (c.first_part + c.count <= nlocal))
      return h->fail("upload_parts_local: the local top-level cells must own exactly the range [0, nlocal)");
  if (alloc_parts(h, ntotal)) return 1;
  h->n_host = nlocal;
  CK(cudaMemcpyAsync(h->d_aos, parts_aos, (size_t)h->cfg.layout.size * (size_t)nlocal, cudaMemcpyHostToDevice,
                     h->stream));
  return transpose_in(h);
}

extern "C" int swiftgpu_upload_parts_device(swiftgpu_t *h, const void *d_parts_aos, int64_t nparts) {
  if (!h || !d_parts_aos || nparts <= 0) return 1;
  cudaSetDevice(h->cfg.device);
  if (alloc_parts(h, nparts)) return 1;
  CK(cudaMemcpyAsync(h->d_aos, d_parts_aos, h->aos_bytes, cudaMemcpyDeviceToDevice, h->stream));
  return transpose_in(h);
}

/* Phase timing without a host round trip: the begin / end events of every phase are recorded on the
 * stream and read back lazily by swiftgpu_get_stats (sync_stats), together with the interaction
 * counters, which live on the device until then. */
enum { PH_SORT = 0, PH_DENSITY, PH_GHOST, PH_GRADIENT, PH_EXTRA_GHOST, PH_FORCE, PH_END_FORCE };
static int phase_begin(H *h, int ph = -1) {
  cudaSetDevice(h->cfg.device);
  if (ensure_lists(h)) return 1;
  h->cur_phase = ph;
  if (ph >= 0) {
    for (int k = 0; k < 2; k++)
      if (!h->pev[ph][k]) CK(cudaEventCreate(&h->pev[ph][k]));
    CK(cudaEventRecord(h->pev[ph][0], h->stream));
  }
  return 0;
}
static int phase_end(H *h, double *) {
  const int ph = h->cur_phase;
  if (ph >= 0) {
    CK(cudaEventRecord(h->pev[ph][1], h->stream));
    h->pev_pending |= 1u << ph;
  }
  h->cur_phase = -1;
  h->counters_pending = true;
  CK(cudaGetLastError());
  return 0;
}
static int sync_stats(H *h) {
  if (!h->pev_pending && !h->counters_pending) return 0;
  unsigned long long c[16];
  CK(cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.n_host_syncs++;
  double *ms[7] = {&h->stats.ms_sort, &h->stats.ms_density, &h->stats.ms_ghost, &h->stats.ms_gradient,
                   &h->stats.ms_extra_ghost, &h->stats.ms_force, &h->stats.ms_end_force};
  for (int ph = 0; ph < 7; ph++)
    if ((h->pev_pending >> ph) & 1u) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, h->pev[ph][0], h->pev[ph][1]) == cudaSuccess) *ms[ph] = t;
    }
  h->pev_pending = 0;
  h->counters_pending = false;
  h->stats.n_density = (int64_t)(c[0] + c[4]); /* + the ghost's re-runs */
  h->stats.n_gradient = (int64_t)c[1];
  h->stats.n_force = (int64_t)c[2];
  h->stats.t_density = (int64_t)c[8];
  h->stats.t_gradient = (int64_t)c[9];
  h->stats.t_force = (int64_t)c[10];
  h->stats.ghost_iterations = (int32_t)c[5];
  return 0;
}

static bool use_cta_loops();
static int loop_kind();
static int launch_extrema(H *h);
static float tile_maxdim(const H *h) {
  return (float)std::max(h->cfg.dim[0], std::max(h->cfg.dim[1], h->cfg.dim[2]));
}
/* absolute widening of the float prefilter on absolute-position floats (error < 4.2e-7 * dim) */
static float tile_margin(const H *h) { return 1.0e-6f * tile_maxdim(h); }
/* r-margin under which the sorted-axis conditions are implied (key rounding < 5e-7 * dim) */
static float tile_keyE(const H *h) { return 2.0e-6f * tile_maxdim(h); }
/* (Re)computes the frame arrays of the current positions (all of them: positions moved, or the
 * frame table grew with a rebuilt list). */
static int ensure_frames(H *h) {
  if (h->frames_valid) return 0;
  if (h->frames_total > h->frames_cap || !h->d_frames) {
    cudaFree(h->d_frames);
    h->d_frames = nullptr;
    h->frames_cap = h->frames_total + h->frames_total / 16 + 64;
    CK(cudaMalloc((void **)&h->d_frames, sizeof(float4) * (size_t)h->frames_cap));
  }
  if (h->frame_recs_uploaded != h->frames_host.size() || !h->d_frame_recs) {
    CK(to_device(&h->d_frame_recs, h->frames_host));
    h->frame_recs_uploaded = h->frames_host.size();
  }
  const int64_t nf = (int64_t)h->frames_host.size();
  if (nf > 0) {
    k_frames<<<(unsigned)((nf * 32 + 127) / 128), 128, 0, h->stream>>>(h->d_frame_recs, nf, h->xs, h->xs + (h->n + 4),
                                                                     h->xs + 2 * (h->n + 4), h->d_frames);
    h->stats.n_launches++;
    CK(cudaGetLastError());
  }
  h->frames_valid = true;
  return 0;
}
static int prep_tiles(H *h) {
  const int64_t n = h->n;
  k_prep_tiles<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->x, h->hh, n, tile_margin(h), h->xf,
                                                                  h->xs);
  h->stats.n_launches++;
  if (loop_kind() == 3) {
    k_octet_boxes_own<<<(h->ncells * 32 + 127) / 128, 128, 0, h->stream>>>(
        h->d_cells, h->ncells, h->d_box_first, h->xs, h->xs + (n + 4), h->xs + 2 * (n + 4), h->boxes);
    h->stats.n_launches++;
    h->frames_valid = false;
    if (ensure_frames(h)) return 1;
  } else {
    k_octet_boxes<<<(h->ncells * 32 + 127) / 128, 128, 0, h->stream>>>(h->d_cells, h->ncells,
                                                                       h->d_box_first, h->xf, h->boxes);
    h->stats.n_launches++;
  }
  CK(cudaGetLastError());
  return 0;
}
/* full = the 13-axis sorted index arrays of runner_do_hydro_sort; the CTA
 * loops only consume the key extrema. */
static int launch_extrema(H *h) {
  if (h->n_ext_cells > 0) {
    k_extrema<<<(h->n_ext_cells * 32 + 127) / 128, 128, 0, h->stream>>>(h->d_ext_cells, h->n_ext_cells,
                                                                        h->d_cells, h->x, h->d_ext);
    h->stats.n_launches++;
    CK(cudaGetLastError());
  }
  h->ext_valid = true;
  return 0;
}
static int launch_sort(H *h, bool full) {
  if (loop_kind() >= 2 && prep_tiles(h)) return 1;
  if (launch_extrema(h)) return 1;
  if (full && h->nsegs > 0) {
    k_sort<<<h->nsegs, 256, 0, h->stream>>>(h->d_segs, h->d_cells, h->x, h->sort_idx, h->d_sort_keys);
    h->stats.n_launches++;
    CK(cudaGetLastError());
    h->full_sorted = true;
  }
  return 0;
}

extern "C" int swiftgpu_run_sort(swiftgpu_t *h) {
  if (!h) return 1;
  if (phase_begin(h, PH_SORT)) return 1;
  h->full_sorted = false;
  if (launch_sort(h, !use_cta_loops())) return 1;
  h->sorted = true;
  h->phases_done |= SWIFTGPU_PHASE_SORT;
  return phase_end(h, &h->stats.ms_sort);
}

#include "launch_host.inl" /* loop_args, build_targets, build_task_recs, launch_pipe*, run_pipe_loop, launch_loop1/2 */

extern "C" int swiftgpu_run_density(swiftgpu_t *h) {
  if (!h) return 1;
  if (ensure_lists(h)) return 1;
  if (!h->sorted && swiftgpu_run_sort(h)) return 1;
  if (phase_begin(h, PH_DENSITY)) return 1;
  const int64_t n = h->n;
  /* the previous ghost raised h_max / h_max_active: start from the uploaded values */
  CK(cudaMemcpyAsync(h->d_cells, h->d_cells_init, sizeof(DevCell) * h->ncells,
                     cudaMemcpyDeviceToDevice, h->stream));
  k_init_parts<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(soa_of(h), h->nd, n,
                                                                  h->step.max_active_bin, h->cfg.scheme);
  h->stats.n_launches++;
  CK(cudaMemsetAsync(h->d_counters + 0, 0, sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->d_counters + 8, 0, sizeof(unsigned long long), h->stream));
  bool sparse = false;
  if (build_targets(h, h->L_density, &sparse)) return 1;
  if (h->L_density.ntasks > 0 && loop_kind() == 3) {
    if (run_pipe_loop<LOOP_DENSITY, false, 0>(h, h->L_density, h->nd, 0, h->d_counters + 12, 0, main_split(),
                                              h->d_counters + 13))
      return 1;
  } else if (h->L_density.ntasks > 0) {
    LoopArgs A = loop_args(h, h->L_density, h->nd, 0);
    CK((launch_loop1<LOOP_DENSITY, false>(h, A, sparse)));
    h->stats.n_launches++;
  }
  h->phases_done |= SWIFTGPU_PHASE_DENSITY;
  h->phases_done &= ~(uint32_t)(SWIFTGPU_PHASE_GHOST | SWIFTGPU_PHASE_GRADIENT |
                                SWIFTGPU_PHASE_EXTRA_GHOST | SWIFTGPU_PHASE_FORCE |
                                SWIFTGPU_PHASE_END_FORCE);
  return phase_end(h, &h->stats.ms_density);
}

template <int SCHEME>
static int ghost_launch(H *h, const GhostArgs &G) {
  k_ghost<SCHEME><<<(G.ngroups * 32 + 127) / 128, 128, 0, h->stream>>>(G);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  return 0;
}

/* Start of one ghost iteration, on the device: counts it if there is anything left to do and zeroes the
 * redo counter the k_ghost pass fills. c[3] redo, c[5] iterations done, c[6] redo before this pass. */
__global__ void k_ghost_begin(unsigned long long *c, int first) {
  if (first || c[3] > 0) c[5]++;
  c[6] = first ? ~0ull : c[3];
  c[3] = 0;
}

extern "C" int swiftgpu_run_ghost(swiftgpu_t *h) {
  if (!h) return 1;
  if (!(h->phases_done & SWIFTGPU_PHASE_DENSITY)) return h->fail("run_ghost before run_density");
  if (phase_begin(h, PH_GHOST)) return 1;
  CK(cudaMemsetAsync(h->d_counters + 3, 0, 4 * sizeof(unsigned long long), h->stream)); /* redo, re-run hits, iterations */
  DevList &D = h->L_subset;
  GhostArgs G;
  memset(&G, 0, sizeof(G));
  G.groups = D.groups;
  G.ngroups = D.ngroups;
  G.cells = h->d_cells;
  G.redo_list = D.tgt_list;
  G.redo_count = D.tgt_count;
  G.S = soa_of(h);
  G.left = h->gleft;
  G.right = h->gright;
  G.nd = h->nd; G.ng = h->ng; G.nf = h->nf;
  G.n_redo = h->d_counters + 3;
  G.max_active_bin = h->step.max_active_bin;
  G.h_max = h->cfg.h_max;
  G.h_min = h->cfg.h_min;
  G.eps = h->cfg.h_tolerance;
  G.eta_dim = h->cfg.eta_neighbours * h->cfg.eta_neighbours * h->cfg.eta_neighbours;
  G.use_mass_weighted = h->cfg.use_mass_weighted_num_ngb;
  G.visc_alpha = h->cfg.viscosity_alpha;
  G.H = h->step.H;
  G.a = h->step.a;
  int iter = 0;
  int64_t redo = 0;
  const int max_iter = h->cfg.max_smoothing_iterations;
  if (D.ngroups > 0 && loop_kind() == 3) {
    /* The whole loop stays on the device: every iteration enqueues the ghost pass and BOTH re-run
     * kernels, each gated on the redo counter (many unconverged particles: the TMA pipeline; a handful
     * per leaf: one warp per target); passes after convergence find nothing to do. The host looks at
     * the counter only every second iteration (runner_ghost.c:1545-1580 loops until count == 0). */
    static int direct_thr = -1;
    if (direct_thr < 0) {
      const char *e = getenv("SWIFTGPU_DIRECT");
      direct_thr = e ? atoi(e) : 4; /* unconverged particles per leaf below which the per-target kernel wins */
    }
    const unsigned long long thr = (unsigned long long)direct_thr * (unsigned long long)D.ngroups;
    const unsigned long long *gate = h->d_counters + 3;
    /* re-runs with fewer unconverged particles than this fraction of the list's particles: small tasks */
    const double sf = sparse_frac();
    const unsigned long long thr_small =
        sf <= 0. ? 0ull : (sf > 1. ? ~0ull : (unsigned long long)(sf * (double)(h->n_host > 0 ? h->n_host : h->n)));
    if (ensure_frames(h)) return 1;
    for (iter = 0; iter < max_iter; iter++) {
      G.first_pass = iter == 0;
      k_ghost_begin<<<1, 1, 0, h->stream>>>(h->d_counters, iter == 0);
      h->stats.n_launches++;
      int rc = 0;
      switch (h->cfg.scheme) {
        case SCH_MINIMAL: rc = ghost_launch<SCH_MINIMAL>(h, G); break;
        case SCH_GADGET2: rc = ghost_launch<SCH_GADGET2>(h, G); break;
        default: rc = ghost_launch<SCH_SPHENIX>(h, G); break;
      }
      if (rc) return 1;
      const bool last = iter + 1 >= max_iter;
      if ((iter & 1) || last) {
        if (read_counter(h, 3, &redo)) return 1;
        if (getenv("SWIFTGPU_VERBOSE"))
          fprintf(stderr, "ghost iteration %d: %lld to redo in %d leaves\n", iter, (long long)redo, D.ngroups);
        if (redo == 0 || last) {
          iter++;
          break;
        }
      }
      /* re-run the density loop for the unconverged particles
       * (runner_dosub_{self,pair}_subset_density, runner_ghost.c:1548-1572) */
      if (run_pipe_loop<LOOP_DENSITY, true, 0>(h, D, h->nd, 4, gate, thr, std::max(thr, thr_small), nullptr)) return 1;
      {
        LoopArgs A = loop_args(h, D, h->nd, 4);
        if (launch_direct_density(h, D, A, gate, thr)) return 1;
      }
    }
  } else if (D.ngroups > 0) {
    for (iter = 0; iter < max_iter; iter++) {
      G.first_pass = iter == 0;
      CK(cudaMemsetAsync(h->d_counters + 3, 0, sizeof(unsigned long long), h->stream));
      int rc = 0;
      switch (h->cfg.scheme) {
        case SCH_MINIMAL: rc = ghost_launch<SCH_MINIMAL>(h, G); break;
        case SCH_GADGET2: rc = ghost_launch<SCH_GADGET2>(h, G); break;
        default: rc = ghost_launch<SCH_SPHENIX>(h, G); break;
      }
      if (rc) return 1;
      if (read_counter(h, 3, &redo)) return 1;
      if (redo == 0) {
        iter++;
        break;
      }
      if (iter + 1 >= max_iter) {
        iter++;
        break;
      }
      /* re-run the density loop for the unconverged particles
       * (runner_dosub_{self,pair}_subset_density, runner_ghost.c:1548-1572) */
      /* few unconverged particles per leaf: (pipe) one warp per target straight from L2, (tile)
       * smaller CTAs, more of them per SM */
      const bool sparse = redo < (int64_t)sparse_threshold() * D.ngroups;
      static int direct_thr = -1;
      if (direct_thr < 0) {
        const char *e = getenv("SWIFTGPU_DIRECT");
        direct_thr = e ? atoi(e) : 4; /* unconverged particles per leaf below which the per-target kernel wins */
      }
      if (getenv("SWIFTGPU_VERBOSE")) fprintf(stderr, "ghost iteration %d: %lld to redo in %d leaves\n", iter, (long long)redo, D.ngroups);
      if (loop_kind() == 3 && redo < (int64_t)direct_thr * D.ngroups) {
        if (ensure_frames(h)) return 1;
        LoopArgs A = loop_args(h, D, h->nd, 4);
        if (launch_direct_density(h, D, A)) return 1;
      } else {
        if (loop_kind() == 3 && (ensure_frames(h) || build_task_recs(h, D, PL_TARGETS))) return 1;
        LoopArgs A = loop_args(h, D, h->nd, 4);
        CK((launch_loop1<LOOP_DENSITY, true>(h, A, sparse)));
        h->stats.n_launches++;
      }
    }
  }
  h->stats.ghost_iterations = iter; /* (frame pipeline: replaced by the device's own count in sync_stats) */
  h->stats.ghost_unconverged = (int32_t)redo;
  h->phases_done |= SWIFTGPU_PHASE_GHOST;
  if (phase_end(h, &h->stats.ms_ghost)) return 1;
  if (redo > 0)
    return h->fail("Smoothing length failed to converge on %lld particles.", (long long)redo);
  return 0;
}

/* Has the ghost (or the rho halo) changed a recursion predicate the list was
 * built with? Then rebuild it from the h_max / h_max_active the device holds. */
static int revalidate_list(H *h, int which) {
  CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int32_t), h->stream));
  if (which == LISTS_FORCE)
    k_pred_bits<<<(h->ncells + 255) / 256, 256, 0, h->stream>>>(h->d_cells, h->d_dmin, h->d_dxp, h->d_force_bits,
                                                                h->ncells, 0, h->d_flag);
  else
    k_pred_bits<<<(h->ncells + 255) / 256, 256, 0, h->stream>>>(h->d_cells, h->d_dmin, h->d_dxp_old,
                                                                h->d_loop1_bits, h->ncells, 1, h->d_flag);
  h->stats.n_launches++;
  int32_t flag = 0;
  CK(cudaMemcpyAsync(&flag, h->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.n_host_syncs++;
  if (which == LISTS_GRADIENT && h->gradient_own) {
    /* an own gradient list exists: still valid if ITS bits match; dropped if the
     * predicates are back to what the density list assumes */
    if (!flag) {
      h->L_gradient.release();
      h->gradient_own = false;
      return 0;
    }
    CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int32_t), h->stream));
    k_pred_bits<<<(h->ncells + 255) / 256, 256, 0, h->stream>>>(h->d_cells, h->d_dmin, h->d_dxp_old,
                                                                h->d_grad_bits, h->ncells, 1, h->d_flag);
    h->stats.n_launches++;
    CK(cudaMemcpyAsync(&flag, h->d_flag, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->stats.n_host_syncs++;
  }
  if (!flag) return 0;
  if (build_lists(h, which)) return 1;
  if (which == LISTS_FORCE) h->stats.force_list_rebuilds++;
  else h->stats.gradient_list_rebuilds++;
  if (loop_kind() != 2) return launch_sort(h, !use_cta_loops());
  /* new (cell, sid) segments may exist: their key extrema */
  return launch_extrema(h);
}

extern "C" int swiftgpu_run_gradient(swiftgpu_t *h) {
  if (!h) return 1;
  if (h->cfg.scheme != SCH_SPHENIX) return 0;
  if (!(h->phases_done & SWIFTGPU_PHASE_GHOST)) return h->fail("run_gradient before run_ghost");
  if (phase_begin(h, PH_GRADIENT)) return 1;
  /* DOSUB_PAIR1/SELF1 of the gradient task evaluate cell_can_recurse_in_subpair/
   * subself_hydro_task (cell.h:951,992) on the h_max_active the ghost just set */
  if (revalidate_list(h, LISTS_GRADIENT)) return 1;
  DevList &L = h->gradient_own ? h->L_gradient : h->L_density;
  CK(cudaMemsetAsync(h->d_counters + 1, 0, sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->d_counters + 9, 0, sizeof(unsigned long long), h->stream));
  bool sparse = false;
  if (build_targets(h, L, &sparse)) return 1; /* depth_h changed in the ghost */
  if (loop_kind() >= 2) {
    k_prep_gq<<<(unsigned)((h->n + 255) / 256), 256, 0, h->stream>>>(h->fq1, h->fq2, h->fq3, h->n, h->gq);
    h->stats.n_launches++;
  }
  if (L.ntasks > 0 && loop_kind() == 3) {
    if (run_pipe_loop<LOOP_GRADIENT, false, 0>(h, L, h->ng, 1, h->d_counters + 12, 0, main_split(), h->d_counters + 13))
      return 1;
  } else if (L.ntasks > 0) {
    LoopArgs A = loop_args(h, L, h->ng, 1);
    CK((launch_loop1<LOOP_GRADIENT, false>(h, A, sparse)));
    h->stats.n_launches++;
  }
  h->phases_done |= SWIFTGPU_PHASE_GRADIENT;
  return phase_end(h, &h->stats.ms_gradient);
}

extern "C" int swiftgpu_run_extra_ghost(swiftgpu_t *h) {
  if (!h) return 1;
  if (h->cfg.scheme != SCH_SPHENIX) return 0;
  if (!(h->phases_done & SWIFTGPU_PHASE_GRADIENT)) return h->fail("run_extra_ghost before run_gradient");
  if (phase_begin(h, PH_EXTRA_GHOST)) return 1;
  ExtraArgs E;
  memset(&E, 0, sizeof(E));
  E.groups = h->L_subset.groups;
  E.ngroups = h->L_subset.ngroups;
  E.cells = h->d_cells;
  E.S = soa_of(h);
  E.nf = h->nf;
  E.max_active_bin = h->step.max_active_bin;
  E.time_base = h->step.time_base;
  E.a = h->step.a;
  E.alpha_max = h->cfg.viscosity_alpha_max;
  E.alpha_min = h->cfg.viscosity_alpha_min;
  E.length = h->cfg.viscosity_length;
  E.beta = h->cfg.diffusion_beta;
  E.diff_alpha_max = h->cfg.diffusion_alpha_max;
  E.diff_alpha_min = h->cfg.diffusion_alpha_min;
  if (E.ngroups > 0) {
    k_extra_ghost<<<(E.ngroups * 32 + 127) / 128, 128, 0, h->stream>>>(E);
    h->stats.n_launches++;
    CK(cudaGetLastError());
  }
  h->phases_done |= SWIFTGPU_PHASE_EXTRA_GHOST;
  return phase_end(h, &h->stats.ms_extra_ghost);
}

extern "C" int swiftgpu_run_force(swiftgpu_t *h) {
  if (!h) return 1;
  const uint32_t need = h->cfg.scheme == SCH_SPHENIX ? SWIFTGPU_PHASE_EXTRA_GHOST : SWIFTGPU_PHASE_GHOST;
  if (!(h->phases_done & need)) return h->fail("run_force before the ghost phases");
  if (phase_begin(h, PH_FORCE)) return 1;
  if (revalidate_list(h, LISTS_FORCE)) return 1;
  CK(cudaMemsetAsync(h->d_counters + 2, 0, sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(h->d_counters + 10, 0, sizeof(unsigned long long), h->stream));
  bool sparse = false;
  if (build_targets(h, h->L_force, &sparse)) return 1;
  if (loop_kind() == 2) { /* h changed in the ghost (and the rho halo): source reach of the prefilter */
    k_refresh_reach<<<(unsigned)((h->n + 255) / 256), 256, 0, h->stream>>>(h->hh, h->n, tile_margin(h),
                                                                          h->xf);
    h->stats.n_launches++;
  }
  if (h->L_force.ntasks > 0 && loop_kind() == 3) {
    const unsigned long long *g = h->d_counters + 12, *d = h->d_counters + 13;
    int rc;
    switch (h->cfg.scheme) {
      case SCH_MINIMAL: rc = run_pipe_loop<LOOP_FORCE, false, SCH_MINIMAL>(h, h->L_force, h->nf, 2, g, 0, main_split(), d); break;
      case SCH_GADGET2: rc = run_pipe_loop<LOOP_FORCE, false, SCH_GADGET2>(h, h->L_force, h->nf, 2, g, 0, main_split(), d); break;
      default: rc = run_pipe_loop<LOOP_FORCE, false, SCH_SPHENIX>(h, h->L_force, h->nf, 2, g, 0, main_split(), d); break;
    }
    if (rc) return 1;
  } else if (h->L_force.ntasks > 0) {
    LoopArgs A = loop_args(h, h->L_force, h->nf, 2);
    switch (h->cfg.scheme) {
      case SCH_MINIMAL: CK(launch_loop2<SCH_MINIMAL>(h, A, sparse)); break;
      case SCH_GADGET2: CK(launch_loop2<SCH_GADGET2>(h, A, sparse)); break;
      default: CK(launch_loop2<SCH_SPHENIX>(h, A, sparse)); break;
    }
    h->stats.n_launches++;
  }
  h->phases_done |= SWIFTGPU_PHASE_FORCE;
  return phase_end(h, &h->stats.ms_force);
}

extern "C" int swiftgpu_run_end_force(swiftgpu_t *h) {
  if (!h) return 1;
  if (!(h->phases_done & SWIFTGPU_PHASE_FORCE)) return h->fail("run_end_force before run_force");
  if (phase_begin(h, PH_END_FORCE)) return 1;
  if (h->L_subset.ngroups > 0) {
    k_end_force<<<(h->L_subset.ngroups * 32 + 127) / 128, 128, 0, h->stream>>>(
        h->L_subset.groups, h->L_subset.ngroups, h->d_cells, soa_of(h), h->step.max_active_bin,
        h->cfg.scheme, h->cfg.CFL_condition, h->step.a,
        /* cosmology.c: a_factor_sound_speed = a^(-1.5 (gamma - 1)) */
        powf(h->step.a, -1.5f * HYDRO_GAMMA_MINUS_ONE), h->dt_cfl);
    h->stats.n_launches++;
    CK(cudaGetLastError());
  }
  h->phases_done |= SWIFTGPU_PHASE_END_FORCE;
  return phase_end(h, &h->stats.ms_end_force);
}

extern "C" int swiftgpu_halo_exchange(swiftgpu_t *h, int phase);

extern "C" int swiftgpu_run_step(swiftgpu_t *h, uint32_t mask) {
  if (!h) return 1;
  const bool halo = h->cfg.nranks > 1 && h->halo_ready;
  /* lists (and a pending re-transpose after new cells) first: the exchange
   * writes into the SoA columns the transpose would overwrite */
  cudaSetDevice(h->cfg.device);
  if (ensure_lists(h)) return 1;
  if (halo && (mask & (SWIFTGPU_PHASE_SORT | SWIFTGPU_PHASE_DENSITY)) && swiftgpu_halo_exchange(h, 0))
    return 1;
  if ((mask & SWIFTGPU_PHASE_SORT) && swiftgpu_run_sort(h)) return 1;
  if ((mask & SWIFTGPU_PHASE_DENSITY) && swiftgpu_run_density(h)) return 1;
  if ((mask & SWIFTGPU_PHASE_GHOST) && swiftgpu_run_ghost(h)) return 1;
  if (halo && (mask & SWIFTGPU_PHASE_GHOST) && swiftgpu_halo_exchange(h, 1)) return 1;
  if ((mask & SWIFTGPU_PHASE_GRADIENT) && swiftgpu_run_gradient(h)) return 1;
  if ((mask & SWIFTGPU_PHASE_EXTRA_GHOST) && swiftgpu_run_extra_ghost(h)) return 1;
  if (halo && (mask & SWIFTGPU_PHASE_EXTRA_GHOST) && swiftgpu_halo_exchange(h, 2)) return 1;
  if ((mask & SWIFTGPU_PHASE_FORCE) && swiftgpu_run_force(h)) return 1;
  if ((mask & SWIFTGPU_PHASE_END_FORCE) && swiftgpu_run_end_force(h)) return 1;
  return 0;
}

static int transpose_out(H *h) {
  /* nothing ran since the particles were transposed in (or drifted): the AoS copy is current */
  if (!(h->phases_done & SWIFTGPU_PHASE_DENSITY)) return 0;
  DevLayout D;
  D.L = h->cfg.layout;
  D.scheme = h->cfg.scheme;
  const int density_only = (h->phases_done & SWIFTGPU_PHASE_GHOST) ? 0 : 1;
  const int64_t nh = h->n_host > 0 ? h->n_host : h->n;
  k_soa_to_aos<<<(unsigned)((nh + 255) / 256), 256, 0, h->stream>>>(
      h->d_aos, D, soa_of(h), nh, h->step.max_active_bin, density_only, h->d_d2h);
  h->stats.n_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int swiftgpu_download_parts(swiftgpu_t *h, void *parts_aos, int64_t nparts) {
  if (!h || !parts_aos || nparts != h->n) return 1;
  cudaSetDevice(h->cfg.device);
  if (transpose_out(h)) return 1;
  CK(cudaMemcpyAsync(parts_aos, h->d_aos, h->aos_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int swiftgpu_download_parts_local(swiftgpu_t *h, void *parts_aos, int64_t nlocal) {
  if (!h || !parts_aos || nlocal <= 0 || nlocal != h->n_host) return 1;
  cudaSetDevice(h->cfg.device);
  if (transpose_out(h)) return 1;
  CK(cudaMemcpyAsync(parts_aos, h->d_aos, (size_t)h->cfg.layout.size * (size_t)nlocal, cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.n_host_syncs++;
  return 0;
}

extern "C" int swiftgpu_download_parts_device(swiftgpu_t *h, void *d_parts_aos, int64_t nparts) {
  if (!h || !d_parts_aos || nparts != h->n) return 1;
  cudaSetDevice(h->cfg.device);
  if (transpose_out(h)) return 1;
  CK(cudaMemcpyAsync(d_parts_aos, h->d_aos, h->aos_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int swiftgpu_download_cells(swiftgpu_t *h, swiftgpu_cell *cells, int32_t ncells) {
  if (!h || !cells || ncells != h->ncells || !h->d_cells) return 1;
  cudaSetDevice(h->cfg.device);
  std::vector<float> hm, hma;
  if (pull_cell_hmax(h, hm, hma)) return 1;
  for (int c = 0; c < ncells; c++) {
    cells[c].h_max = hm[c];
    cells[c].h_max_active = hma[c];
    cells[c].dx_max_part = h->cells[c].dx_max_part; /* as the last drift left them */
    cells[c].dx_max_sort = h->cells[c].dx_max_sort;
  }
  return 0;
}

#include "abi_time_integration.inl" /* drift, kick, limiter loop: swiftgpu_upload_xparts ... swiftgpu_run_limiter */

extern "C" int swiftgpu_download_counts(swiftgpu_t *h, int32_t *n_density, int32_t *n_gradient,
                                        int32_t *n_force, int64_t nparts) {
  if (!h || nparts != h->n) return 1;
  cudaSetDevice(h->cfg.device);
  if (!h->d_cnt_tmp) CK(cudaMalloc((void **)&h->d_cnt_tmp, sizeof(int32_t) * (size_t)nparts));
  const int32_t *src[3] = {h->nd, h->ng, h->nf};
  int32_t *dst[3] = {n_density, n_gradient, n_force};
  for (int k = 0; k < 3; k++) {
    if (!dst[k]) continue;
    /* device order -> host order */
    k_scatter_i32<<<(unsigned)((nparts + 255) / 256), 256, 0, h->stream>>>(src[k], h->d_d2h, nparts,
                                                                          h->d_cnt_tmp);
    h->stats.n_launches++;
    CK(cudaMemcpyAsync(dst[k], h->d_cnt_tmp, sizeof(int32_t) * nparts, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

extern "C" int swiftgpu_download_timestep(swiftgpu_t *h, float *dt_cfl, int64_t nparts) {
  if (!h || !dt_cfl || nparts != h->n) return 1;
  cudaSetDevice(h->cfg.device);
  if (!(h->phases_done & SWIFTGPU_PHASE_END_FORCE)) return h->fail("download_timestep before run_end_force");
  if (!h->d_cnt_tmp) CK(cudaMalloc((void **)&h->d_cnt_tmp, sizeof(int32_t) * (size_t)nparts));
  /* device order -> host order (floats moved as 32-bit words) */
  k_scatter_i32<<<(unsigned)((nparts + 255) / 256), 256, 0, h->stream>>>((const int32_t *)h->dt_cfl, h->d_d2h,
                                                                        nparts, h->d_cnt_tmp);
  h->stats.n_launches++;
  CK(cudaMemcpyAsync(dt_cfl, h->d_cnt_tmp, sizeof(float) * nparts, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int swiftgpu_download_sort(swiftgpu_t *h, int32_t cell, int32_t sid, int32_t *idx_out,
                                      float *key_min, float *key_max) {
  if (!h || cell < 0 || cell >= h->ncells || sid < 0 || sid > 12) return 1;
  cudaSetDevice(h->cfg.device);
  if (!h->sorted) return h->fail("download_sort before run_sort");
  std::vector<DevCell> one(1);
  CK(cudaMemcpy(one.data(), h->d_cells + cell, sizeof(DevCell), cudaMemcpyDeviceToHost));
  const DevCell &c = one[0];
  if (!((c.sort_mask >> sid) & 1)) return h->fail("(cell, sid) has no sorted array: no item needs it");
  if (!h->full_sorted) {
    if (launch_sort(h, true)) return 1;
    CK(cudaStreamSynchronize(h->stream));
  }
  const int rank = __builtin_popcount((unsigned)c.sort_mask & ((1u << sid) - 1u));
  if (idx_out) {
    CK(cudaMemcpy(idx_out, h->sort_idx + c.sort_base + (int64_t)rank * c.count, sizeof(int32_t) * c.count,
                  cudaMemcpyDeviceToHost));
    /* indices are relative to the cell in DEVICE order: report them in the host's order */
    std::vector<int32_t> d2h(c.count);
    CK(cudaMemcpy(d2h.data(), h->d_d2h + c.first, sizeof(int32_t) * c.count, cudaMemcpyDeviceToHost));
    for (int k = 0; k < c.count; k++) idx_out[k] = d2h[idx_out[k]] - c.first;
  }
  float2 e;
  CK(cudaMemcpy(&e, h->d_ext + c.seg_base + rank, sizeof(e), cudaMemcpyDeviceToHost));
  if (key_min) *key_min = e.x;
  if (key_max) *key_max = e.y;
  return 0;
}

extern "C" int swiftgpu_get_stats(swiftgpu_t *h, swiftgpu_stats *out) {
  if (!h || !out) return 1;
  cudaSetDevice(h->cfg.device);
  if (sync_stats(h)) return 1;
  *out = h->stats;
  return 0;
}

extern "C" int swiftgpu_worklist_stats(const swiftgpu_config *cfg, const swiftgpu_step *step,
                                       const swiftgpu_cell *cells, int32_t ncells,
                                       const int32_t *top, int32_t ntop, int loop,
                                       int64_t out[6]) {
  if (!cfg || !step || !cells || !top || !out || ncells <= 0 || ntop <= 0) return 1;
  Flattener F(cells, ncells, top, ntop, cfg->dim, cfg->periodic, cfg->rank, step->ti_current);
  WorkList W;
  if (loop == 3) {
    std::vector<int32_t> aux;
    F.build_subset(W, aux);
  } else if (loop == 0 || loop == 1 || loop == 2) {
    F.build_loop(loop, W);
  } else {
    return 1;
  }
  for (int k = 0; k < 6; k++) out[k] = 0;
  out[0] = (int64_t)W.items.size();
  out[1] = (int64_t)W.groups.size();
  out[3] = (int64_t)W.sort_requests.size();
  for (const Item &it : W.items) {
    out[2] += (int64_t)cells[it.tcell].count * cells[it.scell].count;
    if (it.mode == MODE_SELF || it.mode == MODE_SUB_SELF) out[4]++;
    if (it.min_depth > 0 || it.max_depth < 127) out[5]++;
  }
  return 0;
}

/* Host-only: an order-sensitive 64-bit digest (FNV-1a) of the flattened list of `loop` - every item
 * and group in list order. The flattening runs on host threads; the digest is how the tests check
 * that the list does not depend on their number. */
extern "C" int swiftgpu_worklist_digest(const swiftgpu_config *cfg, const swiftgpu_step *step,
                                        const swiftgpu_cell *cells, int32_t ncells, const int32_t *top,
                                        int32_t ntop, int loop, uint64_t *digest) {
  if (!cfg || !step || !cells || !top || !digest || ncells <= 0 || ntop <= 0) return 1;
  Flattener F(cells, ncells, top, ntop, cfg->dim, cfg->periodic, cfg->rank, step->ti_current);
  WorkList W;
  if (loop == 3) {
    std::vector<int32_t> aux;
    F.build_subset(W, aux);
  } else if (loop == 0 || loop == 1 || loop == 2) {
    F.build_loop(loop, W);
  } else {
    return 1;
  }
  uint64_t hsh = 1469598103934665603ull;
  auto mix = [&](const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t k = 0; k < n; k++) {
      hsh ^= b[k];
      hsh *= 1099511628211ull;
    }
  };
  for (const Item &it : W.items) {
    mix(&it.tcell, 4); mix(&it.scell, 4); mix(&it.mode, 1); mix(&it.sid, 1);
    mix(&it.min_depth, 1); mix(&it.max_depth, 1); mix(it.shift, 3); mix(&it.flags, 1);
  }
  for (const Group &g : W.groups) mix(&g, sizeof(g));
  *digest = hsh;
  return 0;
}

#include "abi_halo.inl" /* swiftgpu_halo_plan, swiftgpu_nccl_unique_id, swiftgpu_halo_setup, swiftgpu_halo_exchange */
