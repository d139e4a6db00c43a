/*
 * halo.cuh - multi-GPU halo exchange of libswiftgpu (included by swiftgpu.cu).
 *
 * Replaces the reference's MPI send/recv tasks of whole `struct part` arrays
 * (scheduler.c:977-988,1088-1112; 128-160 B per particle, three times a step)
 * by NCCL point-to-point transfers of only the SoA columns each phase needs,
 * packed and unpacked by device kernels. libnccl.so.2 is resolved at run time
 * with dlopen so that the single-GPU path has no NCCL dependency.
 */
#ifndef SWIFTGPU_HALO_CUH
#define SWIFTGPU_HALO_CUH

#include <dlfcn.h>

#include <map>

/* ---- minimal NCCL prototypes (nccl.h 2.x ABI) ---- */
typedef struct {
  char internal[128];
} sg_ncclUniqueId;
typedef void *sg_ncclComm_t;
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(sg_ncclUniqueId *) = nullptr;
  int (*CommInitRank)(sg_ncclComm_t *, int, sg_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(sg_ncclComm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, sg_ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, sg_ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi *nccl_api(std::string &err) {
  static NcclApi api;
  if (api.lib) return &api;
  const char *env = getenv("SWIFTGPU_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    err = "cannot load libnccl.so.2 (set SWIFTGPU_NCCL_LIB)";
    return nullptr;
  }
#define SG_SYM(field, name)                                  \
  *(void **)(&api.field) = dlsym(lib, name);                 \
  if (!api.field) {                                          \
    err = std::string("libnccl lacks ") + name;              \
    return nullptr;                                          \
  }
  SG_SYM(GetUniqueId, "ncclGetUniqueId");
  SG_SYM(CommInitRank, "ncclCommInitRank");
  SG_SYM(CommDestroy, "ncclCommDestroy");
  SG_SYM(Send, "ncclSend");
  SG_SYM(Recv, "ncclRecv");
  SG_SYM(GroupStart, "ncclGroupStart");
  SG_SYM(GroupEnd, "ncclGroupEnd");
  SG_SYM(GetErrorString, "ncclGetErrorString");
#undef SG_SYM
  api.lib = lib;
  return &api;
}

/* ---- plan: which top-level cells go to / come from each peer ---- */
struct HaloPlan {
  std::vector<int32_t> send_cells, recv_cells; /* indices into the cell array */
  int64_t nsend = 0, nrecv = 0;                /* particles */
};

/* A local top-level cell is sent to peer p iff it touches (26-neighbourhood,
 * periodic) a top-level cell owned by p; a foreign top-level cell owned by p
 * is received from p iff it touches a local one. Both lists are ordered by
 * cell location, so the two sides agree without negotiation. */
static void build_halo_plans(const swiftgpu_cell *cells, const int32_t *top, int ntop,
                             const double dim[3], int periodic, int rank,
                             std::map<int, HaloPlan> &plans) {
  plans.clear();
  if (ntop <= 0) return;
  const swiftgpu_cell &c0 = cells[top[0]];
  int cdim[3];
  for (int k = 0; k < 3; k++) cdim[k] = std::max(1, (int)std::floor(dim[k] / c0.width[k] + 0.5));
  std::vector<int> grid((size_t)cdim[0] * cdim[1] * cdim[2], -1);
  auto idx = [&](const swiftgpu_cell &c, int k) { return (int)std::floor(c.loc[k] / c.width[k] + 0.5); };
  for (int a = 0; a < ntop; a++) {
    const swiftgpu_cell &c = cells[top[a]];
    grid[((size_t)idx(c, 0) * cdim[1] + idx(c, 1)) * cdim[2] + idx(c, 2)] = a;
  }
  std::map<int, std::vector<char>> sent, recvd;
  for (int a = 0; a < ntop; a++) {
    const swiftgpu_cell &c = cells[top[a]];
    if (c.nodeID != rank) continue;
    const int ix = idx(c, 0), iy = idx(c, 1), iz = idx(c, 2);
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
          if (!dx && !dy && !dz) continue;
          int jx = ix + dx, jy = iy + dy, jz = iz + dz;
          if (periodic) {
            jx = (jx + cdim[0]) % cdim[0];
            jy = (jy + cdim[1]) % cdim[1];
            jz = (jz + cdim[2]) % cdim[2];
          } else if (jx < 0 || jy < 0 || jz < 0 || jx >= cdim[0] || jy >= cdim[1] || jz >= cdim[2])
            continue;
          const int b = grid[((size_t)jx * cdim[1] + jy) * cdim[2] + jz];
          if (b < 0 || b == a) continue;
          const int p = cells[top[b]].nodeID;
          if (p == rank) continue;
          std::vector<char> &s = sent[p], &r = recvd[p];
          if (s.empty()) s.assign(ntop, 0);
          if (r.empty()) r.assign(ntop, 0);
          s[a] = 1;
          r[b] = 1;
        }
  }
  auto by_loc = [&](int32_t u, int32_t v) {
    const swiftgpu_cell &a = cells[u], &b = cells[v];
    if (a.loc[0] != b.loc[0]) return a.loc[0] < b.loc[0];
    if (a.loc[1] != b.loc[1]) return a.loc[1] < b.loc[1];
    return a.loc[2] < b.loc[2];
  };
  for (auto &kv : sent) {
    HaloPlan &P = plans[kv.first];
    for (int a = 0; a < ntop; a++)
      if (kv.second[a] && cells[top[a]].count > 0) P.send_cells.push_back(top[a]);
    const std::vector<char> &r = recvd[kv.first];
    for (int a = 0; a < ntop; a++)
      if (r[a] && cells[top[a]].count > 0) P.recv_cells.push_back(top[a]);
    std::sort(P.send_cells.begin(), P.send_cells.end(), by_loc);
    std::sort(P.recv_cells.begin(), P.recv_cells.end(), by_loc);
    for (int32_t c : P.send_cells) P.nsend += cells[c].count;
    for (int32_t c : P.recv_cells) P.nrecv += cells[c].count;
  }
}

/* ---- device side ---- */
#define HALO_MAX_FIELDS 14
struct HaloFields {
  int n;
  void *ptr[HALO_MAX_FIELDS];
  int esz[HALO_MAX_FIELDS];
};

__device__ __forceinline__ void halo_copy(char *dst, const char *src, int esz) {
  switch (esz) {
    case 1: *dst = *src; break;
    case 4: *(float *)dst = *(const float *)src; break;
    case 16: *(float4 *)dst = *(const float4 *)src; break;
    case 24:
      ((double *)dst)[0] = ((const double *)src)[0];
      ((double *)dst)[1] = ((const double *)src)[1];
      ((double *)dst)[2] = ((const double *)src)[2];
      break;
    default:
      for (int b = 0; b < esz; b++) dst[b] = src[b];
  }
}

/* Slab layout: field-major, every field block padded to 32 bytes. */
__host__ __device__ __forceinline__ size_t halo_field_offset(const HaloFields &F, int f, int64_t n) {
  size_t off = 0;
  for (int g = 0; g < f; g++) off += (((size_t)F.esz[g] * (size_t)n + 31) / 32) * 32;
  return off;
}

/* idx holds particle indices; h2d (may be null: idx are device indices already) maps host-order
 * indices to this rank's device order. */
__global__ void k_halo_pack(HaloFields F, const int32_t *idx, const int32_t *h2d, int64_t n, char *buf) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const size_t p = h2d ? (size_t)h2d[idx[s]] : (size_t)idx[s];
  for (int f = 0; f < F.n; f++) {
    const size_t off = halo_field_offset(F, f, n);
    halo_copy(buf + off + (size_t)F.esz[f] * s, (const char *)F.ptr[f] + (size_t)F.esz[f] * p, F.esz[f]);
  }
}
__global__ void k_halo_unpack(HaloFields F, const int32_t *idx, const int32_t *h2d, int64_t n, const char *buf) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const size_t p = h2d ? (size_t)h2d[idx[s]] : (size_t)idx[s];
  for (int f = 0; f < F.n; f++) {
    const size_t off = halo_field_offset(F, f, n);
    halo_copy((char *)F.ptr[f] + (size_t)F.esz[f] * p, buf + off + (size_t)F.esz[f] * s, F.esz[f]);
  }
}

/* runner_do_recv_part (runner_recv.c:95-137): h_max / h_max_active of every
 * FOREIGN cell (all levels) from its particles. One warp per cell. */
__global__ void k_foreign_hmax(DevCell *cells, int ncells, const float *h, const int8_t *time_bin,
                               int max_active_bin) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= ncells) return;
  DevCell &C = cells[c];
  if (C.flags & 2) return; /* local */
  float hm = 0.f, hma = 0.f;
  for (int k = lane; k < C.count; k += 32) {
    const int p = C.first + k;
    const int tb = time_bin[p];
    if (tb == 58) continue; /* time_bin_inhibited */
    const float hh = h[p];
    hm = fmaxf(hm, hh);
    if (tb <= max_active_bin) hma = fmaxf(hma, hh);
  }
  hm = warp_max(hm);
  hma = warp_max(hma);
  if (lane == 0) {
    C.h_max = hm;
    C.h_max_active = hma;
  }
}

struct HaloPeer {
  int peer = -1;
  int64_t nsend = 0, nrecv = 0;
  int32_t *d_send_idx = nullptr, *d_recv_idx = nullptr;
  char *d_sendbuf = nullptr, *d_recvbuf = nullptr;
};

#endif
