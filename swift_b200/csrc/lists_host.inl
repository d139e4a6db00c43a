/*
 * lists_host.inl - host side of the worklists (included by swiftgpu.cu): the frame table
 * (frame_of), upload of a flattened list and its task arrays (upload_list), build_lists (flatten the
 * reference's recursion with worklist.hpp, device cell table, sort segments, octet-box offsets) and
 * ensure_lists.
 */
/* The frame array an item reads: the source cell's particles relative to the
 * origin of the reference's leaf-level call (functions_hydro.h:1327-1338).
 * slot 0 = the cell's own frame (x - loc): sources of DOPAIR when the cell is
 * the right cell cj, DOSELF_SUBSET, and the prefilter frame of the double
 * modes; slot 1 + sid = x - (cj->loc + shift) when the cell is the left cell ci
 * of a pair of orientation sid. Frames are shared by all lists. */
static uint32_t frame_of(H *h, const Item &it) {
  const int slot = it.mode == MODE_PAIR_R ? 1 + it.sid : 0;
  const swiftgpu_cell &sc = h->cells[it.scell];
  double o[3];
  for (int k = 0; k < 3; k++) {
    if (slot == 0)
      o[k] = sc.loc[k];
    else /* the targets' cell is cj: origin cj->loc + shift, as the device derives it */
      o[k] = h->cells[it.tcell].loc[k] + (double)it.shift[k] * h->cfg.dim[k];
  }
  int32_t &idx = h->frame_idx[(size_t)it.scell * 14 + slot];
  if (idx >= 0) {
    const H::FrameRec &F = h->frames_host[idx];
    if (F.o[0] == o[0] && F.o[1] == o[1] && F.o[2] == o[2]) return F.off;
    /* same (cell, orientation) with another origin (tiny periodic grids): look for it, else append */
    for (size_t k = 0; k < h->frames_host.size(); k++) {
      const H::FrameRec &E = h->frames_host[k];
      if (E.first == (int32_t)sc.first_part && E.count == sc.count && E.o[0] == o[0] && E.o[1] == o[1] &&
          E.o[2] == o[2])
        return E.off;
    }
  }
  H::FrameRec F;
  F.first = (int32_t)sc.first_part;
  F.count = sc.count;
  for (int k = 0; k < 3; k++) F.o[k] = o[k];
  F.off = (uint32_t)h->frames_total;
  F.pad = 0;
  h->frames_total += (uint64_t)sc.count;
  if (idx < 0) idx = (int32_t)h->frames_host.size();
  h->frames_host.push_back(F);
  h->frames_valid = false;
  return F.off;
}

static int loop_kind();
/* targets per entry of the host task list: the frame pipeline has a variant with small tasks for
 * sparse target sets and cuts the list for it (k_task_recs of a launch with larger tasks uses its head) */
static int task_list_chunk() { return loop_kind() == 3 ? 8 * PL_SPARSE_CW : TASK_TARGETS; }
static int upload_list(H *h, const WorkList &W, DevList &D, bool subset) {
  D.release();
  D.ngroups = (int)W.groups.size();
  D.nitems = W.items.size();
  std::vector<int32_t> tg, tc, tfirst(std::max<size_t>(W.groups.size(), 1));
  /* Heaviest groups first (LPT) so that the tail of the launch is short. */
  std::vector<int32_t> order(W.groups.size());
  for (size_t g = 0; g < order.size(); g++) order[g] = (int32_t)g;
  std::stable_sort(order.begin(), order.end(),
                   [&](int32_t a, int32_t b) { return W.groups[a].cost > W.groups[b].cost; });
  int64_t tot = 0;
  for (size_t g = 0; g < W.groups.size(); g++) {
    const swiftgpu_cell &c = h->cells[W.groups[g].tcell];
    /* subset lists address the redo list by the leaf's own particle range */
    tfirst[g] = subset ? (int32_t)c.first_part : (int32_t)tot;
    tot += c.count;
  }
  for (int32_t g : order) {
    const swiftgpu_cell &c = h->cells[W.groups[g].tcell];
    const int tt = task_list_chunk();
    const int nch = (c.count + tt - 1) / tt;
    for (int k = 0; k < nch; k++) {
      tg.push_back(g);
      tc.push_back(k);
    }
  }
  D.ntasks = (int)tg.size();
  D.tgt_total = subset ? h->n : tot;
  std::vector<Item> items(W.items);
  for (Item &it : items) it.sframe = frame_of(h, it);
  if (h->frames_total > 0xfffffff0ull) return h->fail("frame arrays exceed 2^32 entries");
  CK(to_device(&D.items, items));
  CK(to_device(&D.groups, W.groups));
  CK(to_device(&D.task_group, tg));
  CK(to_device(&D.task_chunk, tc));
  CK(to_device(&D.tgt_first, tfirst));
  CK(cudaMalloc((void **)&D.tgt_count, std::max(D.ngroups, 1) * sizeof(int32_t)));
  CK(cudaMemset(D.tgt_count, 0, std::max(D.ngroups, 1) * sizeof(int32_t)));
  CK(cudaMalloc((void **)&D.tgt_list, std::max<int64_t>(D.tgt_total, 1) * sizeof(int32_t)));
  CK(cudaMalloc((void **)&D.task_recs, std::max(D.ntasks, 1) * sizeof(TaskRec)));
  return 0;
}

/* Which lists build_lists() (re)builds. LISTS_ALL starts from the uploaded cells
 * (the gradient loop then shares the density list); the other two rebuild one
 * list after the ghost changed a recursion predicate it depends on, with the
 * h_max / h_max_active the device holds now. */
enum { LISTS_ALL = 0, LISTS_FORCE = 1, LISTS_GRADIENT = 2 };

static int pull_cell_hmax(H *h, std::vector<float> &hm, std::vector<float> &hma) {
  hm.resize(h->ncells);
  hma.resize(h->ncells);
  float *d_tmp = h->d_hmax_tmp;
  k_get_cell_hmax<<<(h->ncells + 255) / 256, 256, 0, h->stream>>>(h->d_cells, h->ncells, d_tmp,
                                                                  d_tmp + h->ncells);
  h->stats.n_launches++;
  CK(cudaMemcpyAsync(hm.data(), d_tmp, sizeof(float) * h->ncells, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(hma.data(), d_tmp + h->ncells, sizeof(float) * h->ncells, cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.n_host_syncs++;
  return 0;
}

/* Builds worklists, the device cell table and the sort segments. */
static int build_lists(H *h, int which) {
  if (!h->has_step) return h->fail("swiftgpu_set_step must be called before running a phase");
  if (h->cells.empty()) return h->fail("no cells uploaded");
  if (h->n <= 0) return h->fail("no particles uploaded");
  cudaSetDevice(h->cfg.device);
  /* after the ghost: the recursion sees the h_max / h_max_active the device holds */
  std::vector<float> hm, hma;
  std::vector<swiftgpu_cell> saved;
  if (which != LISTS_ALL) {
    if (!h->d_cells) return h->fail("list rebuild before the first build");
    if (pull_cell_hmax(h, hm, hma)) return 1;
    saved = h->cells;
    for (int c = 0; c < h->ncells; c++) {
      h->cells[c].h_max = hm[c];
      h->cells[c].h_max_active = hma[c];
    }
  }
  Flattener F(h->cells.data(), h->ncells, h->top.data(), (int)h->top.size(), h->cfg.dim,
              h->cfg.periodic, h->cfg.rank, h->step.ti_current);
  WorkList Wd, Ws, Wf, Wg;
  if (which == LISTS_ALL) {
    h->frames_host.clear();
    h->frame_idx.assign((size_t)h->ncells * 14, -1);
    h->frames_total = 0;
    h->frames_valid = false;
    F.build_loop(0, Wd);
    std::vector<int32_t> aux;
    F.build_subset(Ws, aux);
    h->req_density = Wd.sort_requests;
    h->req_subset = Ws.sort_requests;
    h->req_gradient.clear();
    h->loop1_bits.resize(h->ncells);
    for (int c = 0; c < h->ncells; c++)
      h->loop1_bits[c] = (uint8_t)((Flattener::subpair1(h->cells[c]) ? 1 : 0) |
                                   (Flattener::subself1(h->cells[c]) ? 2 : 0));
  }
  if (which == LISTS_ALL || which == LISTS_FORCE) {
    F.build_loop(2, Wf);
    h->req_force = Wf.sort_requests;
    h->force_bits.resize(h->ncells);
    for (int c = 0; c < h->ncells; c++)
      h->force_bits[c] = (uint8_t)((Flattener::subpair2(h->cells[c]) ? 1 : 0) |
                                   (Flattener::subself2(h->cells[c]) ? 2 : 0));
  }
  if (which == LISTS_GRADIENT) {
    F.build_loop(1, Wg);
    h->req_gradient = Wg.sort_requests;
    std::vector<uint8_t> gb(h->ncells);
    for (int c = 0; c < h->ncells; c++)
      gb[c] = (uint8_t)((Flattener::subpair1(h->cells[c]) ? 1 : 0) | (Flattener::subself1(h->cells[c]) ? 2 : 0));
    CK(to_device(&h->d_grad_bits, gb));
  }
  /* the host copy keeps the uploaded (pre-ghost) values for the density and
   * subset recursions of a later re-run */
  if (which != LISTS_ALL)
    for (int c = 0; c < h->ncells; c++) {
      h->cells[c].h_max = saved[c].h_max;
      h->cells[c].h_max_active = saved[c].h_max_active;
    }

  /* sort segments: union of the requests of all lists */
  std::vector<uint64_t> req(h->req_density);
  req.insert(req.end(), h->req_subset.begin(), h->req_subset.end());
  req.insert(req.end(), h->req_gradient.begin(), h->req_gradient.end());
  req.insert(req.end(), h->req_force.begin(), h->req_force.end());
  std::sort(req.begin(), req.end());
  req.erase(std::unique(req.begin(), req.end()), req.end());

  std::vector<DevCell> dc(h->ncells);
  for (int c = 0; c < h->ncells; c++) {
    const swiftgpu_cell &s = h->cells[c];
    DevCell &d = dc[c];
    memset(&d, 0, sizeof(d));
    for (int k = 0; k < 3; k++) d.loc[k] = s.loc[k];
    if (s.first_part + s.count > 0x7fffffffLL) return h->fail("more than 2^31 particles per GPU");
    d.first = (int32_t)s.first_part;
    d.count = s.count;
    d.h_max = s.h_max;
    d.h_max_active = s.h_max_active;
    d.dx_max_sort = s.dx_max_sort;
    d.h_max_allowed = s.h_max_allowed;
    d.h_min_allowed = s.h_min_allowed;
    d.parent = s.parent;
    d.sort_base = -1;
    d.sort_mask = 0;
    d.depth = (int8_t)s.depth;
    d.width = (float)std::max(s.width[0], std::max(s.width[1], s.width[2]));
    d.dx_max_part = s.dx_max_part;
    d.flags = (uint8_t)((s.ti_end_min == h->step.ti_current ? 1 : 0) |
                        (s.nodeID == h->cfg.rank ? 2 : 0) | (s.split ? 4 : 0));
  }
  std::vector<SortSeg> segs;
  std::vector<int32_t> ext_cells;
  segs.reserve(req.size());
  int64_t off = 0;
  int max_seg = 0;
  for (uint64_t r : req) {
    const int c = (int)(r >> 4), sid = (int)(r & 15);
    if (dc[c].sort_base < 0) {
      dc[c].sort_base = off;
      dc[c].seg_base = (int32_t)segs.size();
      ext_cells.push_back(c);
    }
    dc[c].sort_mask |= (uint16_t)(1u << sid);
    SortSeg s;
    s.cell = c;
    s.sid = sid;
    s.off = off;
    segs.push_back(s);
    off += dc[c].count;
    max_seg = std::max(max_seg, dc[c].count);
  }
  /* A list rebuilt after the ghost keeps the h_max the device already holds
   * (it is newer than the host copy). */
  if (which != LISTS_ALL)
    for (int c = 0; c < h->ncells; c++) {
      dc[c].h_max = hm[c];
      dc[c].h_max_active = hma[c];
    }
  if (which == LISTS_ALL) {
    /* pristine table first (uploaded h_max), then the live one */
    std::vector<DevCell> dc0(dc);
    for (int c = 0; c < h->ncells; c++) {
      dc0[c].h_max = h->cells_uploaded_hmax(c);
      dc0[c].h_max_active = h->cells_uploaded_hmax_active(c);
    }
    CK(to_device(&h->d_cells_init, dc0));
    std::vector<float> dmin(h->ncells), dxp(h->ncells), dxpo(h->ncells);
    for (int c = 0; c < h->ncells; c++) {
      dmin[c] = h->cells[c].dmin;
      dxp[c] = h->cells[c].dx_max_part;
      dxpo[c] = h->cells[c].dx_max_part_old;
    }
    CK(to_device(&h->d_dmin, dmin));
    CK(to_device(&h->d_dxp, dxp));
    CK(to_device(&h->d_dxp_old, dxpo));
    cudaFree(h->d_hmax_tmp);
    h->d_hmax_tmp = nullptr;
    CK(cudaMalloc((void **)&h->d_hmax_tmp, 2 * sizeof(float) * std::max(h->ncells, 1)));
  } else {
    /* the pristine table must know the new segments too (run_density copies it over the live one) */
    std::vector<DevCell> dc0(dc);
    for (int c = 0; c < h->ncells; c++) {
      dc0[c].h_max = h->cells_uploaded_hmax(c);
      dc0[c].h_max_active = h->cells_uploaded_hmax_active(c);
    }
    CK(to_device(&h->d_cells_init, dc0));
  }
  CK(to_device(&h->d_cells, dc));
  {
    /* octet boxes of every cell (tile pipeline) */
    std::vector<int32_t> bf(h->ncells);
    int64_t nb = 0;
    for (int c = 0; c < h->ncells; c++) {
      bf[c] = (int32_t)nb;
      nb += (dc[c].count + 7) / 8;
    }
    if (nb > 0x7fffffffLL) return h->fail("too many octet boxes");
    if (which == LISTS_ALL) CK(to_device(&h->d_box_first, bf));
    if (nb != h->nboxes || !h->boxes) {
      cudaFree(h->boxes);
      h->boxes = nullptr;
      CK(cudaMalloc((void **)&h->boxes, std::max<int64_t>(nb, 1) * 2 * sizeof(float4)));
      h->nboxes = nb;
      h->sorted = false;
    }
  }
  CK(to_device(&h->d_segs, segs));
  h->nsegs = (int)segs.size();
  CK(to_device(&h->d_ext_cells, ext_cells));
  h->n_ext_cells = (int)ext_cells.size();
  cudaFree(h->d_ext);
  h->d_ext = nullptr;
  CK(cudaMalloc((void **)&h->d_ext, std::max<size_t>(segs.size(), 1) * sizeof(float2)));
  h->full_sorted = false;
  if (off != h->sort_total || !h->sort_idx) {
    cudaFree(h->sort_idx);
    h->sort_idx = nullptr;
    CK(cudaMalloc((void **)&h->sort_idx, std::max<int64_t>(off, 1) * sizeof(uint32_t)));
    h->sort_total = off;
    cudaFree(h->d_sort_keys);
    h->d_sort_keys = nullptr;
  }
  if (max_seg > SORT_SMEM_MAX && !h->d_sort_keys)
    CK(cudaMalloc((void **)&h->d_sort_keys, std::max<int64_t>(off, 1) * sizeof(float)));
  h->ext_valid = false; /* the key extrema follow the segments */
  if (which == LISTS_ALL) h->sorted = false;

  if (which == LISTS_ALL) {
    if (upload_list(h, Wd, h->L_density, false)) return 1;
    if (upload_list(h, Ws, h->L_subset, true)) return 1;
    h->L_gradient.release();
    h->gradient_own = false;
    CK(to_device(&h->d_loop1_bits, h->loop1_bits));
  }
  if (which == LISTS_ALL || which == LISTS_FORCE) {
    if (upload_list(h, Wf, h->L_force, false)) return 1;
    CK(to_device(&h->d_force_bits, h->force_bits));
  }
  if (which == LISTS_GRADIENT) {
    if (upload_list(h, Wg, h->L_gradient, false)) return 1;
    h->gradient_own = true;
  }
  h->lists_built = true;
  return 0;
}

static int transpose_in(H *h);
static int ensure_lists(H *h) {
  /* new cells after the particles were transposed: redo the device order from
   * the AoS copy (the step restarts from the uploaded particle state) */
  if (h->perm_stale && h->d_aos && h->n > 0 && transpose_in(h)) return 1;
  if (h->lists_built) return 0;
  return build_lists(h, LISTS_ALL);
}

