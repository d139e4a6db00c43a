/*
 * host_tree.cpp - host-side helper that produces the INPUT of libswiftgpu: the
 * flattened cell tree a SWIFT engine would hand over (its own space->cells_top
 * + progeny). It is not on the accelerated path; the benchmark and the tests
 * need it because there is no SWIFT engine around them.
 *
 * It follows the reference's construction rules so that the tree is the one
 * SWIFT would build for the same particles:
 *   - top-level grid: cell index = (int)(x * cdim / dim) per axis, cells stored
 *     in cell_getid order k + cdim[2]*(j + cdim[1]*i)  (src/cell.h:532,
 *     src/space_regrid.c:300-330)
 *   - a cell is split into 8 progeny while count > space_splitsize (=400)
 *     (src/space_split.c:195-199, src/space.h:49)
 *   - progeny k sits at loc + width/2 * ((k&4)?1:0, (k&2)?1:0, (k&1)?1:0) and
 *     receives the particles with x >= pivot in the matching axes
 *     (src/space_split.c:243-245, src/cell_split.c:109-110); empty progeny are
 *     dropped (space_split.c:283-287)
 *   - dmin halves per level; h_min_allowed = dmin/2/kernel_gamma,
 *     h_max_allowed = dmin/kernel_gamma evaluated in double then stored as
 *     float (space_split.c:230-231, space_regrid.c:314-315)
 *   - leaves collect h_max, h_max_active, ti_end_min and set every particle's
 *     depth_h (space_split.c:470-505, cell.h:1787-1815)
 *   - top-level cell -> rank by the uniform grid rule of
 *     partition_uniform_grid (src/partition.c:104-121).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/swiftgpu.h"

namespace {

const float kKernelGamma = 1.825742f; /* kernel_hydro.h:52 (cubic spline 3D) */

struct Builder {
  const double *x;
  const float *h;
  const int8_t *time_bin;
  int64_t ti_current;
  int max_active_bin;
  int splitsize;
  std::vector<swiftgpu_cell> cells;
  int8_t *depth_h; /* per particle, new order */
};

inline int64_t integer_timestep(int bin) {
  /* timeline.h:59 */
  return bin <= 0 ? 0 : (int64_t)1 << (bin + 1);
}
inline int64_t integer_time_end(int64_t ti_current, int bin) {
  /* timeline.h:126 */
  const int64_t dti = integer_timestep(bin);
  if (dti == 0) return 0;
  const int64_t mod = ti_current % dti;
  return mod == 0 ? ti_current : ti_current - mod + dti;
}

void set_h_limits(swiftgpu_cell &c) {
  c.h_min_allowed = (float)((double)c.dmin * 0.5 * (1. / (double)kKernelGamma));
  c.h_max_allowed = (float)((double)c.dmin * (1. / (double)kKernelGamma));
}

/* cell_set_part_h_depth, cell.h:1787 */
int part_h_depth(const std::vector<swiftgpu_cell> &cells, int leaf, float h,
                 int fallback) {
  const swiftgpu_cell *c = &cells[leaf];
  if (h < c->h_min_allowed) return c->depth;
  int ci = leaf;
  while (ci >= 0) {
    c = &cells[ci];
    if (h >= c->h_min_allowed && h < c->h_max_allowed) return c->depth;
    ci = c->parent;
  }
  return fallback;
}

/* Recursive split of the cell `ci` whose particles are ind[first..first+count)
 * (indices into the ORIGINAL arrays), reordering `ind` in place. */
void split_recursive(Builder &b, int ci, int64_t *ind, int64_t *scratch) {
  const int64_t first = b.cells[ci].first_part;
  const int count = b.cells[ci].count;
  if (count > b.splitsize) {
    b.cells[ci].split = 1;
    const swiftgpu_cell parent = b.cells[ci];
    const double pivot[3] = {parent.loc[0] + parent.width[0] / 2,
                             parent.loc[1] + parent.width[1] / 2,
                             parent.loc[2] + parent.width[2] / 2};
    int bucket_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint8_t> bid(count);
    for (int k = 0; k < count; k++) {
      const double *px = &b.x[3 * ind[first + k]];
      const int id = (px[0] >= pivot[0]) * 4 + (px[1] >= pivot[1]) * 2 +
                     (px[2] >= pivot[2]);
      bid[k] = (uint8_t)id;
      bucket_count[id]++;
    }
    int bucket_offset[9];
    bucket_offset[0] = 0;
    for (int k = 0; k < 8; k++)
      bucket_offset[k + 1] = bucket_offset[k] + bucket_count[k];
    int fill[8];
    for (int k = 0; k < 8; k++) fill[k] = bucket_offset[k];
    for (int k = 0; k < count; k++) scratch[first + fill[bid[k]]++] = ind[first + k];
    std::memcpy(ind + first, scratch + first, sizeof(int64_t) * count);

    float h_max = 0.f, h_max_active = 0.f;
    int64_t ti_end_min = INT64_MAX;
    for (int k = 0; k < 8; k++) {
      if (bucket_count[k] == 0) {
        b.cells[ci].progeny[k] = -1;
        continue;
      }
      swiftgpu_cell cp;
      std::memset(&cp, 0, sizeof(cp));
      for (int d = 0; d < 3; d++) {
        cp.loc[d] = parent.loc[d];
        cp.width[d] = parent.width[d] / 2;
      }
      cp.dmin = parent.dmin / 2;
      set_h_limits(cp);
      if (k & 4) cp.loc[0] += cp.width[0];
      if (k & 2) cp.loc[1] += cp.width[1];
      if (k & 1) cp.loc[2] += cp.width[2];
      cp.depth = parent.depth + 1;
      cp.split = 0;
      cp.parent = ci;
      for (int j = 0; j < 8; j++) cp.progeny[j] = -1;
      cp.nodeID = parent.nodeID;
      cp.top = parent.top;
      cp.first_part = first + bucket_offset[k];
      cp.count = bucket_count[k];
      const int cpi = (int)b.cells.size();
      b.cells.push_back(cp);
      b.cells[ci].progeny[k] = cpi;
      split_recursive(b, cpi, ind, scratch);
      h_max = std::max(h_max, b.cells[cpi].h_max);
      h_max_active = std::max(h_max_active, b.cells[cpi].h_max_active);
      ti_end_min = std::min(ti_end_min, b.cells[cpi].ti_end_min);
    }
    b.cells[ci].h_max = h_max;
    b.cells[ci].h_max_active = h_max_active;
    b.cells[ci].ti_end_min = ti_end_min;
  } else {
    swiftgpu_cell &c = b.cells[ci];
    c.split = 0;
    for (int j = 0; j < 8; j++) c.progeny[j] = -1;
    float h_max = 0.f, h_max_active = 0.f;
    int64_t ti_end_min = INT64_MAX;
    for (int k = 0; k < count; k++) {
      const int64_t p = ind[first + k];
      const int bin = b.time_bin[p];
      ti_end_min = std::min(ti_end_min, integer_time_end(b.ti_current, bin));
      h_max = std::max(h_max, b.h[p]);
      if (bin <= b.max_active_bin) h_max_active = std::max(h_max_active, b.h[p]);
    }
    c.h_max = h_max;
    c.h_max_active = h_max_active;
    c.ti_end_min = ti_end_min;
    for (int k = 0; k < count; k++) {
      const int64_t p = ind[first + k];
      b.depth_h[first + k] = (int8_t)part_h_depth(b.cells, ci, b.h[p], 0);
    }
  }
  swiftgpu_cell &c = b.cells[ci];
  c.h_max_old = c.h_max;
  c.dx_max_part = c.dx_max_part_old = 0.f;
  c.dx_max_sort = c.dx_max_sort_old = 0.f;
}

}  // namespace

extern "C" {

struct swifthost_tree {
  std::vector<swiftgpu_cell> cells;
  std::vector<int32_t> top;
};

/*
 * Builds the tree. perm[new] = old index of the particle that lands in slot
 * `new` of the cell-ordered particle array; depth_h[new] its depth_h.
 * rank_grid[3] is the brick grid of ranks (1,1,1 for one GPU).
 */
swifthost_tree *swifthost_build_tree(const double *x, const float *h,
                                     const int8_t *time_bin, int64_t n,
                                     const double dim[3], const int cdim[3],
                                     int splitsize, int max_active_bin,
                                     int64_t ti_current, const int rank_grid[3],
                                     int64_t *perm, int8_t *depth_h) {
  swifthost_tree *t = new swifthost_tree();
  Builder b;
  b.x = x;
  b.h = h;
  b.time_bin = time_bin;
  b.ti_current = ti_current;
  b.max_active_bin = max_active_bin;
  b.splitsize = splitsize;
  b.depth_h = depth_h;

  const int ntop = cdim[0] * cdim[1] * cdim[2];
  const double width[3] = {dim[0] / cdim[0], dim[1] / cdim[1], dim[2] / cdim[2]};
  const double iwidth[3] = {1.0 / width[0], 1.0 / width[1], 1.0 / width[2]};

  /* Counting sort of the particles into top-level cells
   * (space_parts_get_cell_index_mapper, src/space_cell_index.c). */
  std::vector<int32_t> cid(n);
  std::vector<int64_t> count(ntop + 1, 0);
  for (int64_t k = 0; k < n; k++) {
    int idx[3];
    for (int d = 0; d < 3; d++) {
      int v = (int)(x[3 * k + d] * iwidth[d]);
      if (v < 0) v = 0;
      if (v >= cdim[d]) v = cdim[d] - 1;
      idx[d] = v;
    }
    const int c = idx[2] + cdim[2] * (idx[1] + cdim[1] * idx[0]);
    cid[k] = c;
    count[c + 1]++;
  }
  for (int c = 0; c < ntop; c++) count[c + 1] += count[c];
  {
    std::vector<int64_t> fill(count.begin(), count.end() - 1);
    for (int64_t k = 0; k < n; k++) perm[fill[cid[k]]++] = k;
  }
  std::vector<int64_t> scratch(n);

  const float dmin = (float)std::min(width[0], std::min(width[1], width[2]));
  b.cells.reserve((size_t)(ntop * 1.3) + 16);
  t->top.resize(ntop);
  for (int i = 0; i < cdim[0]; i++)
    for (int j = 0; j < cdim[1]; j++)
      for (int k = 0; k < cdim[2]; k++) {
        const int c = k + cdim[2] * (j + cdim[1] * i);
        swiftgpu_cell cell;
        std::memset(&cell, 0, sizeof(cell));
        cell.loc[0] = i * width[0];
        cell.loc[1] = j * width[1];
        cell.loc[2] = k * width[2];
        for (int d = 0; d < 3; d++) cell.width[d] = width[d];
        cell.dmin = dmin;
        set_h_limits(cell);
        cell.depth = 0;
        cell.parent = -1;
        for (int q = 0; q < 8; q++) cell.progeny[q] = -1;
        /* partition_uniform_grid, partition.c:112-120 */
        int ind[3];
        ind[0] = (int)(cell.loc[0] / dim[0] * rank_grid[0]);
        ind[1] = (int)(cell.loc[1] / dim[1] * rank_grid[1]);
        ind[2] = (int)(cell.loc[2] / dim[2] * rank_grid[2]);
        cell.nodeID = ind[0] + rank_grid[0] * (ind[1] + rank_grid[1] * ind[2]);
        cell.top = c;
        cell.first_part = count[c];
        cell.count = (int32_t)(count[c + 1] - count[c]);
        cell.ti_end_min = INT64_MAX;
        b.cells.push_back(cell);
        t->top[c] = c;
      }
  /* Depth-first split, top-level cell by top-level cell (progeny are appended
   * behind the top-level block, so indices stay valid). */
  for (int c = 0; c < ntop; c++) split_recursive(b, c, perm, scratch.data());

  t->cells.swap(b.cells);
  return t;
}

int32_t swifthost_tree_ncells(const swifthost_tree *t) { return (int32_t)t->cells.size(); }
int32_t swifthost_tree_ntop(const swifthost_tree *t) { return (int32_t)t->top.size(); }
void swifthost_tree_copy(const swifthost_tree *t, swiftgpu_cell *cells, int32_t *top) {
  std::memcpy(cells, t->cells.data(), sizeof(swiftgpu_cell) * t->cells.size());
  std::memcpy(top, t->top.data(), sizeof(int32_t) * t->top.size());
}
void swifthost_tree_free(swifthost_tree *t) { delete t; }

/*
 * Pack SoA host columns into the AoS `struct part` records of `layout`
 * (what a SWIFT engine already holds), applying perm (new <- old).
 * Any column pointer may be NULL (field left zero). `cols` order:
 *  0 x(3 f64) 1 v(3 f32) 2 mass 3 h 4 u_or_entropy 5 id(i64) 6 time_bin(i8)
 *  7 depth_h(i8, already in new order) 8 visc_alpha 9 diff_alpha
 *  10 div_v_previous_step 11 rho
 */
void swifthost_pack_parts(const swiftgpu_part_layout *L, int scheme, int64_t n,
                          const int64_t *perm, const double *x, const float *v,
                          const float *mass, const float *h, const float *u,
                          const int64_t *id, const int8_t *time_bin,
                          const int8_t *depth_h, const float *visc_alpha,
                          const float *diff_alpha, const float *div_v_prev,
                          const float *rho, void *out) {
  char *base = (char *)out;
  std::memset(base, 0, (size_t)L->size * n);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < n; k++) {
    const int64_t o = perm ? perm[k] : k;
    char *p = base + (size_t)L->size * k;
    if (x) std::memcpy(p + L->x, &x[3 * o], 3 * sizeof(double));
    if (v) std::memcpy(p + L->v, &v[3 * o], 3 * sizeof(float));
    if (mass) std::memcpy(p + L->mass, &mass[o], sizeof(float));
    if (h) std::memcpy(p + L->h, &h[o], sizeof(float));
    if (u) {
      const int off = scheme == SWIFTGPU_SCHEME_GADGET2 ? L->entropy : L->u;
      std::memcpy(p + off, &u[o], sizeof(float));
    }
    if (id) std::memcpy(p + L->id, &id[o], sizeof(int64_t));
    if (time_bin) std::memcpy(p + L->time_bin, &time_bin[o], 1);
    if (depth_h) std::memcpy(p + L->depth_h, &depth_h[k], 1);
    if (visc_alpha && L->visc_alpha >= 0)
      std::memcpy(p + L->visc_alpha, &visc_alpha[o], sizeof(float));
    if (diff_alpha && L->diff_alpha >= 0)
      std::memcpy(p + L->diff_alpha, &diff_alpha[o], sizeof(float));
    if (div_v_prev && L->div_v_previous_step >= 0)
      std::memcpy(p + L->div_v_previous_step, &div_v_prev[o], sizeof(float));
    if (rho) std::memcpy(p + L->rho, &rho[o], sizeof(float));
  }
}

} /* extern "C" */
