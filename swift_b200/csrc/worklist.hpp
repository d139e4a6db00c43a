/*
 * worklist.hpp - host-side flattening of the reference's recursive task
 * functions into flat lists of leaf-level interaction items.
 *
 * The reference walks the cell tree at task-execution time
 * (runner_dosub_{self,pair}{1,2}_<loop>, runner_doiact_functions_hydro.h:
 * 2933 DOSUB_PAIR1, 3037 DOSUB_SELF1, 3102 DOSUB_PAIR2, 3203 DOSUB_SELF2, and
 * the ghost's redo walkers 3290 DOSUB_PAIR_SUBSET, 3341 DOSUB_SELF_SUBSET).
 * WHERE that recursion stops fixes the float frame in which every pair
 * distance is evaluated (cj->loc + shift of the cells at that level), so the
 * device must evaluate exactly the same (ci, cj, level) items to reproduce the
 * reference's neighbour sets bit for bit. The walk is O(cells) and runs once
 * per tree (density/subset lists) or per change of the h_max-dependent
 * recursion predicates (force list).
 *
 * Every reference leaf call DOPAIR(ci,cj) updates both cells; here it becomes
 * two directed GATHER items (targets in ci / targets in cj). Items are grouped
 * by target cell so that one warp owns a particle's accumulators.
 */
#ifndef SWIFTGPU_WORKLIST_HPP
#define SWIFTGPU_WORKLIST_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/swiftgpu.h"

namespace swiftgpu {

static const float kKernelGammaF = (float)(1.825742); /* kernel_hydro.h:52 */
static const int kRecurseSizeSelf = 100;              /* space.h:64 */
static const int kRecurseSizePair = 100;              /* space.h:65 */

/* sort_part.h:42-58 */
static const double kRunnerShift[13][3] = {
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
/* sortlistID of sort_part.h:63-91: symmetric in the 27 directions. */
inline int sortlist_id(int dir27) { return dir27 <= 13 ? (dir27 == 13 ? 0 : dir27) : 26 - dir27; }

enum ItemMode : int {
  MODE_SELF = 0,        /* DOSELF1 / DOSELF2: (float)(x_t - x_s) on doubles */
  MODE_PAIR_L = 1,      /* DOPAIR1/2, targets in ci (left cell), sources cj ascending */
  MODE_PAIR_R = 2,      /* DOPAIR1/2, targets in cj, sources ci descending */
  MODE_SUB_SELF = 3,    /* DOSELF_SUBSET: floats relative to c->loc */
  MODE_SUB_PAIR = 4,    /* DOPAIR_SUBSET, not flipped: sources ascending */
  MODE_SUB_PAIR_F = 5   /* DOPAIR_SUBSET, flipped: sources descending */
};

/* 24-byte directed item. `tcell` holds the targets, `scell` the sources.
 * For the PAIR modes ci/cj (after space_getsid_and_swap_cells) are
 * (tcell,scell) for MODE_PAIR_L and (scell,tcell) for MODE_PAIR_R. */
struct Item {
  int32_t tcell;
  int32_t scell;
  uint8_t mode;
  uint8_t sid;
  int8_t min_depth;
  int8_t max_depth;
  int8_t shift[3]; /* periodic shift in units of dim[k] (of the ORIENTED pair) */
  uint8_t flags;   /* bit0: limit_max_h (h_max clamp = ci->h_max_allowed) */
  uint32_t sframe; /* offset (float4 units) of the source cell's frame array this item reads (loops_pipe.cuh) */
  uint32_t pad_;
};
static_assert(sizeof(Item) == 24, "Item must be 24 bytes");

/* A group = all items sharing a target cell, contiguous in the item array. */
struct Group {
  int32_t tcell;
  int32_t item_first;
  int32_t item_count;
  int32_t cost; /* sum of source counts, for ordering */
};

struct WorkList {
  std::vector<Item> items;
  std::vector<Group> groups;
  /* (cell, sid) pairs whose sorted index array is needed */
  std::vector<uint64_t> sort_requests; /* cell * 16 + sid */
};

class Flattener {
 public:
  Flattener(const swiftgpu_cell *cells, int ncells, const int32_t *top, int ntop,
            const double dim[3], int periodic, int rank, int64_t ti_current)
      : c_(cells), ncells_(ncells), top_(top), ntop_(ntop), periodic_(periodic),
        rank_(rank), ti_current_(ti_current) {
    for (int k = 0; k < 3; k++) dim_[k] = dim[k];
  }

  /* Host threads of the flattening (SWIFTGPU_HOST_THREADS, default: the hardware's, at most 16). The
   * result does not depend on it: every worker walks a contiguous range of top-level cells with its
   * own emission buffers, which are concatenated in the serial order. */
  static int host_threads() {
    static int v = 0;
    if (v == 0) {
      const char *e = getenv("SWIFTGPU_HOST_THREADS");
      v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      v = std::max(1, std::min(v, 16));
    }
    return v;
  }
  template <class Fn>
  void parallel_ranges(int n, Fn body /* (Flattener &worker, int lo, int hi, int part) */) const {
    const int T = std::max(1, std::min(host_threads(), n / 8));
    std::vector<Flattener> workers;
    workers.reserve(T);
    for (int t = 0; t < T; t++) workers.emplace_back(c_, ncells_, top_, ntop_, dim_, periodic_, rank_, ti_current_);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
      const int lo = (int)((int64_t)n * t / T), hi = (int)((int64_t)n * (t + 1) / T);
      if (T == 1)
        body(workers[t], lo, hi, t);
      else
        th.emplace_back([&, t, lo, hi]() { body(workers[t], lo, hi, t); });
    }
    for (std::thread &x : th) x.join();
  }

  /* loop: 0 density, 1 gradient (same decomposition as density), 2 force */
  void build_loop(int loop, WorkList &out) {
    raw_.clear();
    const int T = std::max(1, std::min(host_threads(), ntop_ / 8));
    std::vector<std::vector<Raw>> selfs(T), pairs(T);
    parallel_ranges(ntop_, [&](Flattener &w, int lo, int hi, int part) {
      w.raw_.clear();
      for (int a = lo; a < hi; a++) {
        const int ca = top_[a];
        if (c_[ca].nodeID == rank_) w.dosub_self(loop, ca, 0);
      }
      selfs[part].swap(w.raw_);
      w.raw_.clear();
      w.for_each_top_pair([&](int ca, int cb) { w.dosub_pair(loop, ca, cb, 0); }, lo, hi);
      pairs[part].swap(w.raw_);
    });
    for (int t = 0; t < T; t++) raw_.insert(raw_.end(), selfs[t].begin(), selfs[t].end());
    for (int t = 0; t < T; t++) raw_.insert(raw_.end(), pairs[t].begin(), pairs[t].end());
    finish(out, /*by_leaf=*/false);
  }

  /* The ghost's redo items of every active local leaf
   * (runner_ghost.c:1548-1572 with the density tasks linked at the top level).
   * tcell of an item is the LEAF (the unit whose redo particles are the
   * targets); the stop-level cell of the recursion, which fixes frame and
   * shift, is stored in `aux`. */
  void build_subset(WorkList &out, std::vector<int32_t> &aux) {
    raw_.clear();
    raw_aux_.clear();
    /* neighbour lists of the top-level cells */
    std::vector<std::vector<int>> ngb(ntop_);
    for_each_top_pair_all([&](int a, int b) { ngb[a].push_back(b); });
    const int T = std::max(1, std::min(host_threads(), ntop_ / 8));
    std::vector<std::vector<Raw>> parts(T);
    parallel_ranges(ntop_, [&](Flattener &w, int lo, int hi, int part) {
      w.raw_.clear();
      std::vector<int> leaves;
      for (int a = lo; a < hi; a++) {
        const int ca = top_[a];
        if (c_[ca].nodeID != rank_) continue;
        leaves.clear();
        w.collect_active_leaves(ca, leaves);
        for (int leaf : leaves) {
          w.cur_leaf_ = leaf;
          w.dosub_self_subset(ca);
          for (int b : ngb[a]) w.dosub_pair_subset(ca, top_[b]);
        }
      }
      parts[part].swap(w.raw_);
    });
    for (int t = 0; t < T; t++) raw_.insert(raw_.end(), parts[t].begin(), parts[t].end());
    finish(out, /*by_leaf=*/true);
    aux.swap(sorted_aux_);
  }

  /* Bits of the h_max-dependent recursion predicates the lists were built with
   * (force: cell.h:966 subpair2, :1007 subself2; density/gradient: :951
   * subpair, :992 subself), for validation after the ghost. Only cells the
   * recursion can descend from (split, at least space_recurse_size particles)
   * ever evaluate them: all other cells report 0. */
  static inline bool recursable(const swiftgpu_cell &c) {
    return c.split && c.count >= kRecurseSizePair;
  }
  static inline bool subpair2(const swiftgpu_cell &c) {
    return recursable(c) && (kKernelGammaF * c.h_max + c.dx_max_part) < 0.5f * c.dmin;
  }
  static inline bool subself2(const swiftgpu_cell &c) {
    return recursable(c) && (kKernelGammaF * c.h_max < 0.5f * c.dmin);
  }
  static inline bool subpair1(const swiftgpu_cell &c) {
    return recursable(c) && (kKernelGammaF * c.h_max_active + c.dx_max_part_old) < 0.5f * c.dmin;
  }
  static inline bool subself1(const swiftgpu_cell &c) {
    return recursable(c) && (kKernelGammaF * c.h_max_active < 0.5f * c.dmin);
  }

 private:
  struct Raw {
    Item it;
    int32_t aux;
  };
  const swiftgpu_cell *c_;
  int ncells_;
  const int32_t *top_;
  int ntop_;
  double dim_[3];
  int periodic_;
  int rank_;
  int64_t ti_current_;
  std::vector<Raw> raw_;
  std::vector<int32_t> raw_aux_, sorted_aux_;
  int cur_leaf_ = -1;

  bool active(int c) const { return c_[c].ti_end_min == ti_current_; }
  bool local(int c) const { return c_[c].nodeID == rank_; }

  /* space_getsid_and_swap_cells, space_getsid.h:47-80 */
  int getsid(int &ci, int &cj, int8_t shift[3]) const {
    int dir = 0;
    int sh[3];
    for (int k = 0; k < 3; k++) {
      double dx = c_[cj].loc[k] - c_[ci].loc[k];
      double s = 0.0;
      sh[k] = 0;
      if (periodic_ && dx < -dim_[k] / 2) {
        s = dim_[k];
        sh[k] = 1;
      } else if (periodic_ && dx > dim_[k] / 2) {
        s = -dim_[k];
        sh[k] = -1;
      }
      dx += s;
      dir = 3 * dir + ((dx < 0.0) ? 0 : ((dx > 0.0) ? 2 : 1));
    }
    if (dir < 13) { /* runner_flip */
      std::swap(ci, cj);
      for (int k = 0; k < 3; k++) sh[k] = -sh[k];
    }
    for (int k = 0; k < 3; k++) shift[k] = (int8_t)sh[k];
    return sortlist_id(dir);
  }

  /* cell_split_pairs (cell.c:63) derived geometrically: progeny bit 4 -> x,
   * 2 -> y, 1 -> z (space_split.c:243-245); octants pid of ci and pjd of cj
   * form a sub-pair iff they touch when cj sits at dir(sid) from ci. */
  static int sub_pairs(int sid, int pid[16], int pjd[16]) {
    /* 13 small tables, computed once */
    struct Table {
      int n[13], a[13][16], b[13][16];
      Table() {
        for (int s = 0; s < 13; s++) n[s] = sub_pairs_compute(s, a[s], b[s]);
      }
    };
    static const Table T;
    for (int k = 0; k < T.n[sid]; k++) {
      pid[k] = T.a[sid][k];
      pjd[k] = T.b[sid][k];
    }
    return T.n[sid];
  }
  static int sub_pairs_compute(int sid, int pid[16], int pjd[16]) {
    int d[3];
    for (int k = 0; k < 3; k++)
      d[k] = kRunnerShift[sid][k] > 0 ? 1 : (kRunnerShift[sid][k] < 0 ? -1 : 0);
    int n = 0;
    for (int a = 0; a < 8; a++)
      for (int b = 0; b < 8; b++) {
        const int ax[3] = {(a >> 2) & 1, (a >> 1) & 1, a & 1};
        const int bx[3] = {(b >> 2) & 1, (b >> 1) & 1, b & 1};
        bool ok = true;
        for (int k = 0; k < 3; k++)
          if (std::abs(bx[k] + 2 * d[k] - ax[k]) > 1) ok = false;
        if (ok) {
          pid[n] = a;
          pjd[n] = b;
          n++;
        }
      }
    return n;
  }

  template <class F>
  void for_each_top_pair_all(F f, int a_lo = 0, int a_hi = -1) const {
    /* every ordered couple (a,b), a != b, of touching top-level cells, a in [a_lo, a_hi) */
    if (a_hi < 0) a_hi = ntop_;
    const swiftgpu_cell &c0 = c_[top_[0]];
    int cdim[3];
    for (int k = 0; k < 3; k++) cdim[k] = (int)std::floor(dim_[k] / c0.width[k] + 0.5);
    std::vector<int> grid((size_t)cdim[0] * cdim[1] * cdim[2], -1);
    auto idx = [&](const swiftgpu_cell &c, int k) {
      return (int)std::floor(c.loc[k] / c.width[k] + 0.5);
    };
    for (int a = 0; a < ntop_; a++) {
      const swiftgpu_cell &c = c_[top_[a]];
      grid[((size_t)idx(c, 0) * cdim[1] + idx(c, 1)) * cdim[2] + idx(c, 2)] = a;
    }
    std::vector<int> seen;
    for (int a = a_lo; a < a_hi; a++) {
      const swiftgpu_cell &c = c_[top_[a]];
      const int ix = idx(c, 0), iy = idx(c, 1), iz = idx(c, 2);
      seen.clear();
      for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dz = -1; dz <= 1; dz++) {
            if (!dx && !dy && !dz) continue;
            int jx = ix + dx, jy = iy + dy, jz = iz + dz;
            if (periodic_) {
              jx = (jx + cdim[0]) % cdim[0];
              jy = (jy + cdim[1]) % cdim[1];
              jz = (jz + cdim[2]) % cdim[2];
            } else if (jx < 0 || jy < 0 || jz < 0 || jx >= cdim[0] || jy >= cdim[1] ||
                       jz >= cdim[2])
              continue;
            const int b = grid[((size_t)jx * cdim[1] + jy) * cdim[2] + jz];
            if (b < 0 || b == a) continue;
            if (std::find(seen.begin(), seen.end(), b) != seen.end()) continue;
            seen.push_back(b);
            f(a, b);
          }
    }
  }
  template <class F>
  void for_each_top_pair(F f, int a_lo = 0, int a_hi = -1) const {
    /* unordered couples with at least one local side
     * (engine_maketasks.c:3562-3569), the smaller index in [a_lo, a_hi) */
    for_each_top_pair_all([&](int a, int b) {
      if (b < a) return;
      if (!local(top_[a]) && !local(top_[b])) return;
      f(top_[a], top_[b]);
    }, a_lo, a_hi);
  }

  void emit(int mode, int t, int s, int sid, const int8_t shift[3], int min_depth,
            int max_depth, int limit_max_h, int aux = -1) {
    Raw r;
    r.it.tcell = t;
    r.it.scell = s;
    r.it.mode = (uint8_t)mode;
    r.it.sid = (uint8_t)sid;
    r.it.min_depth = (int8_t)min_depth;
    r.it.max_depth = (int8_t)max_depth;
    for (int k = 0; k < 3; k++) r.it.shift[k] = shift ? shift[k] : 0;
    r.it.flags = (uint8_t)(limit_max_h ? 1 : 0);
    r.it.sframe = 0;
    r.it.pad_ = 0;
    r.aux = aux;
    raw_.push_back(r);
  }

  /* leaf calls: DOSELF{1,2}_BRANCH / DOPAIR{1,2}_BRANCH */
  void leaf_self(int loop, int c, int limit_min_h, int limit_max_h) {
    if (!active(c)) return;
    const int min_depth = limit_max_h ? c_[c].depth : 0;
    const int max_depth = limit_min_h ? c_[c].depth : 127;
    emit(MODE_SELF, c, c, 0, nullptr, min_depth, max_depth, limit_max_h);
  }
  void leaf_pair(int loop, int ci, int cj, int sid, const int8_t shift[3],
                 int limit_min_h, int limit_max_h) {
    const int min_depth = limit_max_h ? c_[ci].depth : 0;
    const int max_depth = limit_min_h ? c_[ci].depth : 127;
    /* Foreign particles are never updated (functions_hydro.h:1650-1651). */
    if (active(ci) && local(ci))
      emit(MODE_PAIR_L, ci, cj, sid, shift, min_depth, max_depth, limit_max_h);
    if (active(cj) && local(cj))
      emit(MODE_PAIR_R, cj, ci, sid, shift, min_depth, max_depth, limit_max_h);
  }

  /* only called on cells that are split and hold >= space_recurse_size particles */
  bool can_recurse_subpair(int loop, const swiftgpu_cell &c) const {
    if (loop == 2) return (kKernelGammaF * c.h_max + c.dx_max_part) < 0.5f * c.dmin; /* cell.h:966 */
    return (kKernelGammaF * c.h_max_active + c.dx_max_part_old) < 0.5f * c.dmin;     /* :951 */
  }
  bool can_recurse_subself(int loop, const swiftgpu_cell &c) const {
    if (loop == 2) return c.split && (kKernelGammaF * c.h_max < 0.5f * c.dmin); /* cell.h:1007 */
    return (kKernelGammaF * c.h_max_active < 0.5f * c.dmin);                    /* :992 */
  }

  void dosub_pair(int loop, int ci, int cj, int below) {
    if (!active(ci) && !active(cj)) return;
    if (c_[ci].count == 0 || c_[cj].count == 0) return;
    int8_t shift[3];
    const int sid = getsid(ci, cj, shift);
    const swiftgpu_cell &a = c_[ci], &b = c_[cj];
    if (!a.split || a.count < kRecurseSizePair || !b.split || b.count < kRecurseSizePair) {
      leaf_pair(loop, ci, cj, sid, shift, 0, below);
    } else {
      if (!below && (!can_recurse_subpair(loop, a) || !can_recurse_subpair(loop, b))) below = 1;
      if (below) leaf_pair(loop, ci, cj, sid, shift, 1, 1);
      int pid[16], pjd[16];
      const int n = sub_pairs(sid, pid, pjd);
      for (int k = 0; k < n; k++)
        if (a.progeny[pid[k]] >= 0 && b.progeny[pjd[k]] >= 0)
          dosub_pair(loop, a.progeny[pid[k]], b.progeny[pjd[k]], below);
    }
  }

  void dosub_self(int loop, int c, int below) {
    const swiftgpu_cell &a = c_[c];
    if (a.count == 0 || !active(c)) return;
    if (!a.split || a.count < kRecurseSizeSelf) {
      leaf_self(loop, c, 0, below);
    } else {
      if (!below && !can_recurse_subself(loop, a)) below = 1;
      if (below) leaf_self(loop, c, 1, 1);
      for (int k = 0; k < 8; k++)
        if (a.progeny[k] >= 0) {
          dosub_self(loop, a.progeny[k], below);
          for (int j = k + 1; j < 8; j++)
            if (a.progeny[j] >= 0) dosub_pair(loop, a.progeny[k], a.progeny[j], below);
        }
    }
  }

  /* ---- subset walkers ---- */
  void collect_active_leaves(int c, std::vector<int> &out) const {
    if (c_[c].count == 0 || !active(c)) return; /* runner_ghost.c:1147-1148 */
    if (c_[c].split) {
      for (int k = 0; k < 8; k++)
        if (c_[c].progeny[k] >= 0) collect_active_leaves(c_[c].progeny[k], out);
    } else
      out.push_back(c);
  }
  bool contains(int c, int leaf) const {
    return c_[leaf].first_part >= c_[c].first_part &&
           c_[leaf].first_part < c_[c].first_part + c_[c].count;
  }
  int find_sub(int c) const { /* FIND_SUB :3267 */
    for (int k = 0; k < 8; k++)
      if (c_[c].progeny[k] >= 0 && contains(c_[c].progeny[k], cur_leaf_)) return c_[c].progeny[k];
    return -1;
  }
  static bool can_recurse_pair_task(const swiftgpu_cell &c) { /* cell.h:933 */
    return c.split && ((kKernelGammaF * c.h_max_old + c.dx_max_part_old) < 0.5f * c.dmin);
  }
  static bool can_recurse_self_task(const swiftgpu_cell &c) { /* cell.h:978 */
    return c.split && (kKernelGammaF * c.h_max_old < 0.5f * c.dmin);
  }
  void dosub_pair_subset(int ci, int cj) { /* :3290 */
    if (c_[ci].count == 0 || c_[cj].count == 0) return;
    if (!active(ci)) return;
    if (can_recurse_pair_task(c_[ci]) && can_recurse_pair_task(c_[cj])) {
      const int sub = find_sub(ci);
      int a = ci, b = cj;
      int8_t shift[3];
      const int sid = getsid(a, b, shift);
      int pid[16], pjd[16];
      const int n = sub_pairs(sid, pid, pjd);
      for (int k = 0; k < n; k++) {
        const int pa = c_[a].progeny[pid[k]], pb = c_[b].progeny[pjd[k]];
        if (pa == sub && pb >= 0) dosub_pair_subset(pa, pb);
        if (pa >= 0 && pb == sub) dosub_pair_subset(pb, pa);
      }
    } else {
      /* DOPAIR_SUBSET_BRANCH :1036: shift and sid from ci (unswapped) to cj */
      int8_t shift[3];
      int dir = 0;
      for (int k = 0; k < 3; k++) {
        const double d = c_[cj].loc[k] - c_[ci].loc[k];
        double s = 0.0;
        shift[k] = 0;
        if (d < -dim_[k] / 2) {
          s = dim_[k];
          shift[k] = 1;
        } else if (d > dim_[k] / 2) {
          s = -dim_[k];
          shift[k] = -1;
        }
        dir = 3 * dir + ((d + s < 0) ? 0 : (d + s > 0) ? 2 : 1);
      }
      const int flipped = dir < 13;
      emit(flipped ? MODE_SUB_PAIR_F : MODE_SUB_PAIR, cur_leaf_, cj, sortlist_id(dir), shift, 0,
           127, 0, ci);
    }
  }
  void dosub_self_subset(int ci) { /* :3341 */
    if (c_[ci].count == 0 || !active(ci)) return;
    if (c_[ci].split && can_recurse_self_task(c_[ci])) {
      const int sub = find_sub(ci);
      dosub_self_subset(sub);
      for (int j = 0; j < 8; j++)
        if (c_[ci].progeny[j] != sub && c_[ci].progeny[j] >= 0)
          dosub_pair_subset(sub, c_[ci].progeny[j]);
    } else
      emit(MODE_SUB_SELF, cur_leaf_, ci, 0, nullptr, 0, 127, 0, ci);
  }

  /* Order of the items INSIDE a group (free: it only permutes sums). The loops stream the source
   * cells of a group through a ring of stages that the 8 consumer warps of a CTA - each owning one
   * octant of the target leaf - drain in lock-step up to the ring depth. A source cell on the +x
   * side keeps the +x warps busy and the -x warps idle; a run of same-side cells (the sid order
   * the recursion emits) therefore piles up skew that the ring cannot absorb. Greedy re-ordering:
   * next comes the item that keeps the running sum of (direction x source count) smallest, which
   * alternates opposite sides. */
  void balance_directions(std::vector<uint32_t> &order) const {
    /* groups are independent: the threads take contiguous slices of `order` cut at group boundaries */
    const int T = order.size() < 100000 ? 1 : host_threads();
    std::vector<size_t> cut(T + 1, order.size());
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
      size_t p = order.size() * (size_t)t / (size_t)T;
      while (p > 0 && p < order.size() && raw_[order[p]].it.tcell == raw_[order[p - 1]].it.tcell) p++;
      cut[t] = std::max(p, cut[t - 1]);
    }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
      if (T == 1)
        balance_range(order, cut[t], cut[t + 1]);
      else
        th.emplace_back([&, t]() { balance_range(order, cut[t], cut[t + 1]); });
    }
    for (std::thread &x : th) x.join();
  }
  void balance_range(std::vector<uint32_t> &order, size_t begin, size_t end) const {
    std::vector<uint32_t> tmp;
    std::vector<float> d;
    std::vector<char> used;
    for (size_t g0 = begin; g0 < end;) {
      size_t g1 = g0 + 1;
      const int32_t tc = raw_[order[g0]].it.tcell;
      while (g1 < end && raw_[order[g1]].it.tcell == tc) g1++;
      const size_t n = g1 - g0;
      if (n > 2 && n <= 4096) {
        d.assign(3 * n, 0.f);
        const swiftgpu_cell &T = c_[tc];
        for (size_t k = 0; k < n; k++) {
          const swiftgpu_cell &S = c_[raw_[order[g0 + k]].it.scell];
          for (int a = 0; a < 3; a++) {
            double x = (S.loc[a] + 0.5 * S.width[a]) - (T.loc[a] + 0.5 * T.width[a]);
            if (periodic_) x -= dim_[a] * std::floor(x / dim_[a] + 0.5);
            d[3 * k + a] = (float)(x / T.width[a]) * (float)S.count;
          }
        }
        used.assign(n, 0);
        tmp.resize(n);
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (size_t pos = 0; pos < n; pos++) {
          size_t best = 0;
          float bestv = 3.0e38f;
          for (size_t k = 0; k < n; k++) {
            if (used[k]) continue;
            const float ax = sx + d[3 * k], ay = sy + d[3 * k + 1], az = sz + d[3 * k + 2];
            const float v = ax * ax + ay * ay + az * az;
            if (v < bestv) {
              bestv = v;
              best = k;
            }
          }
          used[best] = 1;
          tmp[pos] = order[g0 + best];
          sx += d[3 * best];
          sy += d[3 * best + 1];
          sz += d[3 * best + 2];
        }
        for (size_t k = 0; k < n; k++) order[g0 + k] = tmp[k];
      }
      g0 = g1;
    }
  }

  void finish(WorkList &out, bool keep_aux) {
    /* group by target cell, stable: counting sort on the cell index */
    std::vector<uint32_t> order(raw_.size());
    {
      std::vector<uint32_t> first((size_t)ncells_ + 1, 0);
      for (const Raw &r : raw_) first[(size_t)r.it.tcell + 1]++;
      for (int c = 0; c < ncells_; c++) first[c + 1] += first[c];
      for (uint32_t i = 0; i < (uint32_t)raw_.size(); i++) order[first[raw_[i].it.tcell]++] = i;
    }
    if (!getenv("SWIFTGPU_NO_BALANCE")) balance_directions(order); /* A/B knob */
    out.items.resize(raw_.size());
    out.groups.clear();
    sorted_aux_.clear();
    if (keep_aux) sorted_aux_.resize(raw_.size());
    std::vector<uint64_t> req;
    for (size_t k = 0; k < order.size(); k++) {
      const Raw &r = raw_[order[k]];
      out.items[k] = r.it;
      if (keep_aux) sorted_aux_[k] = r.aux;
      if (out.groups.empty() || out.groups.back().tcell != r.it.tcell) {
        Group g;
        g.tcell = r.it.tcell;
        g.item_first = (int32_t)k;
        g.item_count = 0;
        g.cost = 0;
        out.groups.push_back(g);
      }
      out.groups.back().item_count++;
      out.groups.back().cost += c_[r.it.scell].count;
      if (r.it.mode == MODE_PAIR_L || r.it.mode == MODE_PAIR_R || r.it.mode == MODE_SUB_PAIR ||
          r.it.mode == MODE_SUB_PAIR_F)
        req.push_back((uint64_t)r.it.scell * 16 + r.it.sid);
      if (r.it.mode == MODE_PAIR_L || r.it.mode == MODE_PAIR_R)
        req.push_back((uint64_t)r.it.tcell * 16 + r.it.sid); /* dj_min / di_max */
    }
    std::sort(req.begin(), req.end());
    req.erase(std::unique(req.begin(), req.end()), req.end());
    out.sort_requests.swap(req);
    raw_.clear();
    raw_.shrink_to_fit();
  }
};

}  // namespace swiftgpu
#endif
