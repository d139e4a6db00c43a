/*
 * abi_halo.inl - the multi-GPU entry points (included by swiftgpu.cu): halo plans from the cells,
 * NCCL communicator set-up, the three per-step exchanges. Device side and NCCL loader: halo.cuh.
 */
/* ======================================================================== */
/* Multi-GPU halo exchange                                                   */
/* ======================================================================== */
extern "C" int swiftgpu_halo_plan(const swiftgpu_config *cfg, const swiftgpu_cell *cells,
                                  int32_t ncells, const int32_t *top, int32_t ntop, int32_t peer,
                                  int32_t *send_cells, int32_t *nsend_cells, int32_t *recv_cells,
                                  int32_t *nrecv_cells, int64_t *nsend_parts, int64_t *nrecv_parts) {
  if (!cfg || !cells || !top || ncells <= 0 || ntop <= 0) return 1;
  std::map<int, HaloPlan> plans;
  build_halo_plans(cells, top, ntop, cfg->dim, cfg->periodic, cfg->rank, plans);
  HaloPlan P;
  if (plans.count(peer)) P = plans[peer];
  if (nsend_cells) *nsend_cells = (int32_t)P.send_cells.size();
  if (nrecv_cells) *nrecv_cells = (int32_t)P.recv_cells.size();
  if (nsend_parts) *nsend_parts = P.nsend;
  if (nrecv_parts) *nrecv_parts = P.nrecv;
  if (send_cells) std::copy(P.send_cells.begin(), P.send_cells.end(), send_cells);
  if (recv_cells) std::copy(P.recv_cells.begin(), P.recv_cells.end(), recv_cells);
  return 0;
}

extern "C" int swiftgpu_nccl_unique_id(void *id128) {
  if (!id128) return 1;
  NcclApi *N = nccl_api(g_err);
  if (!N) return 2;
  sg_ncclUniqueId id;
  const int rc = N->GetUniqueId(&id);
  if (rc != 0) {
    g_err = std::string("ncclGetUniqueId: ") + N->GetErrorString(rc);
    return 2;
  }
  memcpy(id128, &id, sizeof(id));
  return 0;
}

/* The SoA columns each phase moves (see include/swiftgpu.h). */
static HaloFields halo_fields(H *h, int phase) {
  HaloFields F;
  memset(&F, 0, sizeof(F));
  auto add = [&](void *p, int esz) {
    F.ptr[F.n] = p;
    F.esz[F.n] = esz;
    F.n++;
  };
  const bool sph = h->cfg.scheme == SCH_SPHENIX;
  if (phase == 0) {
    add(h->x, 24); add(h->mv, 16); add(h->hh, 4); add(h->u, 4); add(h->rho, 4);
    add(h->time_bin, 1); add(h->depth_h, 1); add(h->fq1, 16); add(h->fq2, 16);
    if (sph) { add(h->fq3, 16); add(h->alpha, 4); add(h->alpha_diff, 4); }
  } else if (phase == 1) {
    add(h->hh, 4); add(h->rho, 4); add(h->depth_h, 1); add(h->fq1, 16); add(h->fq2, 16);
    if (sph) add(h->fq3, 16);
  } else {
    add(h->fq3, 16); add(h->alpha, 4); add(h->alpha_diff, 4);
  }
  return F;
}

extern "C" int swiftgpu_halo_setup(swiftgpu_t *h, const void *id128) {
  if (!h || !id128) return 1;
  if (h->cfg.nranks <= 1) return h->fail("halo_setup: nranks is 1");
  if (h->cells.empty()) return h->fail("halo_setup: upload the cells first");
  cudaSetDevice(h->cfg.device);
  std::string e;
  NcclApi *N = nccl_api(e);
  if (!N) return h->fail("%s", e.c_str());
  /* new cells: new send / receive lists; the communicator (one per handle, an NCCL id can be used
   * once) is kept */
  halo_release(h, /*keep_comm=*/true);
  if (!h->comm) {
    sg_ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    int rc = N->CommInitRank(&h->comm, h->cfg.nranks, id, h->cfg.rank);
    if (rc != 0) return h->fail("ncclCommInitRank: %s", N->GetErrorString(rc));
  }

  std::map<int, HaloPlan> plans;
  build_halo_plans(h->cells.data(), h->top.data(), (int)h->top.size(), h->cfg.dim, h->cfg.periodic,
                   h->cfg.rank, plans);
  /* widest phase decides the slab size */
  size_t per_part = 0;
  for (int ph = 0; ph < 3; ph++) {
    HaloFields F = halo_fields(h, ph);
    size_t b = 0;
    for (int f = 0; f < F.n; f++) b += F.esz[f];
    per_part = std::max(per_part, b);
  }
  for (auto &kv : plans) {
    HaloPeer P;
    P.peer = kv.first;
    P.nsend = kv.second.nsend;
    P.nrecv = kv.second.nrecv;
    std::vector<int32_t> si, ri;
    si.reserve(P.nsend);
    ri.reserve(P.nrecv);
    /* DEVICE indices: a cell is the same contiguous range on the host and on the device, and its
     * particles travel in the sender's device order (Morton order inside its leaves), which the
     * receiver adopts for its proxy of the cell */
    for (int32_t c : kv.second.send_cells)
      for (int k = 0; k < h->cells[c].count; k++) si.push_back((int32_t)h->cells[c].first_part + k);
    for (int32_t c : kv.second.recv_cells)
      for (int k = 0; k < h->cells[c].count; k++) ri.push_back((int32_t)h->cells[c].first_part + k);
    CK(to_device(&P.d_send_idx, si));
    CK(to_device(&P.d_recv_idx, ri));
    CK(cudaMalloc((void **)&P.d_sendbuf, per_part * std::max<int64_t>(P.nsend, 1) + 32 * HALO_MAX_FIELDS));
    CK(cudaMalloc((void **)&P.d_recvbuf, per_part * std::max<int64_t>(P.nrecv, 1) + 32 * HALO_MAX_FIELDS));
    h->halo.push_back(P);
  }
  h->halo_ready = true;
  return 0;
}

extern "C" int swiftgpu_halo_exchange(swiftgpu_t *h, int phase) {
  if (!h) return 1;
  if (h->cfg.nranks <= 1) return 0;
  if (!h->halo_ready) return h->fail("halo_exchange before halo_setup");
  if (phase < 0 || phase > 2) return h->fail("halo_exchange: bad phase");
  if (phase == 2 && h->cfg.scheme != SCH_SPHENIX) return 0;
  if (!h->x) return h->fail("halo_exchange: no particles uploaded");
  cudaSetDevice(h->cfg.device);
  std::string e;
  NcclApi *N = nccl_api(e);
  if (!N) return h->fail("%s", e.c_str());
  HaloFields F = halo_fields(h, phase);
  int64_t bytes = 0;
  for (HaloPeer &P : h->halo) {
    if (P.nsend > 0) {
      k_halo_pack<<<(unsigned)((P.nsend + 255) / 256), 256, 0, h->stream>>>(F, P.d_send_idx, nullptr, P.nsend,
                                                                          P.d_sendbuf);
      h->stats.n_launches++;
    }
  }
  CK(cudaGetLastError());
  int rc = N->GroupStart();
  if (rc != 0) return h->fail("ncclGroupStart: %s", N->GetErrorString(rc));
  for (HaloPeer &P : h->halo) {
    const size_t sb = halo_field_offset(F, F.n, P.nsend), rb = halo_field_offset(F, F.n, P.nrecv);
    if (P.nsend > 0) {
      rc = N->Send(P.d_sendbuf, sb, /*ncclChar*/ 0, P.peer, h->comm, h->stream);
      if (rc != 0) return h->fail("ncclSend: %s", N->GetErrorString(rc));
      bytes += (int64_t)sb;
    }
    if (P.nrecv > 0) {
      rc = N->Recv(P.d_recvbuf, rb, 0, P.peer, h->comm, h->stream);
      if (rc != 0) return h->fail("ncclRecv: %s", N->GetErrorString(rc));
    }
  }
  rc = N->GroupEnd();
  if (rc != 0) return h->fail("ncclGroupEnd: %s", N->GetErrorString(rc));
  for (HaloPeer &P : h->halo) {
    if (P.nrecv > 0) {
      k_halo_unpack<<<(unsigned)((P.nrecv + 255) / 256), 256, 0, h->stream>>>(F, P.d_recv_idx, nullptr, P.nrecv,
                                                                            P.d_recvbuf);
      h->stats.n_launches++;
    }
  }
  if (phase == 1 && h->d_cells) {
    /* foreign h changed: refresh the foreign cells' h_max like runner_do_recv_part */
    k_foreign_hmax<<<(h->ncells * 32 + 127) / 128, 128, 0, h->stream>>>(
        h->d_cells, h->ncells, h->hh, h->time_bin, h->step.max_active_bin);
    h->stats.n_launches++;
  }
  CK(cudaGetLastError());
  if (phase == 0) h->sorted = false; /* foreign x / h moved: tile records, octet boxes and key extrema are stale */
  h->halo_bytes[phase] = bytes;
  return 0;
}
