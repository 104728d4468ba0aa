// Multi-GPU BBFMM matvec: one process per GPU, the tree partitioned by Morton-contiguous leaf ranges,
// NCCL over NVLink for the two exchange steps.  The reference is single-process rayon (bbfmm.rs:669, 682, 788, 841,
// 1122 are its parallel loops); this file is what replaces "one rayon pool" when the tree spans several B200s.
//
//   ownership   leaves in Morton order, cut into `world` contiguous ranges of nearly equal estimated work
//               (fb_tree_leaf_work); a rank owns the points of its range as sources AND as targets.  Tree topology and
//               point coordinates are replicated (24 B per point), coefficients and results are not.
//   upward      P2M over the OWNED leaves only, M2M over the cells that have owned descendants: every rank holds the
//               exact multipoles of the cells inside its range and a partial sum for the cells that span ranks;
//   exchange 1  ncclAllReduce(sum) of the multipole array completes the spanning cells and hands every rank the halo
//               multipoles its V / W lists need, on a side stream UNDER the near-field pass (P2P needs no multipoles);
//   downward    M2L / P2L / L2L restricted to cells with owned targets, L2P / M2P for owned targets;
//   exchange 2  ncclAllGather of the owned result rows (padded to the largest share) + one scatter kernel: the full
//               result, replicated, in the caller's row order — the next Krylov vector needs exactly that.
// NCCL is loaded at run time (dlopen of the copy already in the process, else libnccl.so.2): the single-GPU drop-in
// has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <numeric>

#include <nccl.h>

#include "fmm.h"

namespace fb {

// ---- NCCL entry points, resolved once ------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl() {
  static NcclApi api = [] {
    NcclApi a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    // the copy already mapped into this process (torch ships its own) wins: two NCCLs in one process is asking for it
    for (const char *nm : names)
      if (!a.handle) a.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (const char *p = std::getenv("FB_NCCL_LIB"))
      if (!a.handle) a.handle = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    for (const char *nm : names)
      if (!a.handle) a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
#define FB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
    FB_SYM(GetUniqueId, "ncclGetUniqueId");
    FB_SYM(CommInitRank, "ncclCommInitRank");
    FB_SYM(CommDestroy, "ncclCommDestroy");
    FB_SYM(AllReduce, "ncclAllReduce");
    FB_SYM(AllGather, "ncclAllGather");
    FB_SYM(GetErrorString, "ncclGetErrorString");
#undef FB_SYM
    return a;
  }();
  return api;
}

static void require_nccl() {
  NcclApi &a = nccl();
  if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather)
    throw Error(FB_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded; set FB_NCCL_LIB to its path");
}

#define FB_NCCL(expr)                                                                                        \
  do {                                                                                                       \
    ncclResult_t r_ = (expr);                                                                                \
    if (r_ != ncclSuccess)                                                                                   \
      throw fb::Error(FB_ERR_CUDA, std::string(#expr) + ": " +                                               \
                                       (fb::nccl().GetErrorString ? fb::nccl().GetErrorString(r_) : "NCCL error")); \
  } while (0)

// contiguous ranges of nearly equal work: boundaries into the leaf sequence (the same cut sharding.py makes)
void partition_by_work(const double *work, size_t n, int parts, uint64_t *bounds) {
  bounds[0] = 0;
  bounds[parts] = n;
  if (parts <= 1) return;
  if (n == 0) {
    for (int k = 1; k < parts; ++k) bounds[k] = 0;
    return;
  }
  std::vector<double> csum(n);
  double run = 0;
  for (size_t i = 0; i < n; ++i) csum[i] = (run += work[i]);
  const double total = csum[n - 1];
  for (int k = 1; k < parts; ++k) {
    const double target = total * (double)k / (double)parts;
    const size_t cut = (size_t)(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin()) + 1;
    bounds[k] = std::max<uint64_t>(std::min(cut, n), bounds[k - 1]);
  }
}

// full result in the caller's row order from the gathered, padded, Morton-ordered shares
__global__ void k_scatter_gathered(const double *gathered, const int *shard_pos, int world, size_t max_rows, int nrhs,
                                   const uint32_t *perm, size_t n, double *full) {
  const size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (pos >= n) return;
  int r = 0;
  while (r + 1 < world && (size_t)shard_pos[r + 1] <= pos) ++r;
  const double *src = gathered + ((size_t)r * max_rows + (pos - (size_t)shard_pos[r])) * nrhs;
  double *dst = full + (size_t)perm[pos] * nrhs;
  for (int k = 0; k < nrhs; ++k) dst[k] = src[k];
}

__global__ void k_iota_u64(const uint32_t *perm, size_t begin, size_t count, unsigned long long *out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < count) out[i] = perm[begin + i];
}

}  // namespace fb

using namespace fb;

struct fb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;  // collectives that run under compute
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  ~fb_comm() {
    if (comm && nccl().CommDestroy) nccl().CommDestroy(comm);
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (ev_done) cudaEventDestroy(ev_done);
    if (stream) cudaStreamDestroy(stream);
  }
};

// per-tree state of the partition (fb_tree::shard)
struct fb_shard {
  fb_comm *comm = nullptr;
  std::vector<uint64_t> leaf_bounds;  // world + 1 boundaries into the Morton leaf sequence
  std::vector<int> pos;               // world + 1 boundaries into the sorted point order
  size_t max_rows = 0;                // largest share
  DBuf<int> d_pos;
  DBuf<int> d_owned_leaves;           // cells of the owned leaves that hold sources (P2M grid)
  int n_owned_leaves = 0;
  DBuf<double> d_gather, d_full;
  TargetBuffers tb;
  TargetSet ts{};
  DBuf<unsigned long long> d_rows;
  std::vector<int> level_lo, level_hi;  // per level: the cells with owned targets are the ids [lo, hi)
  M2LItemTable *m2l_table = nullptr;    // M2L work items of that share
  double last_ms[4] = {0, 0, 0, 0};   // upward, all-reduce wait, downward + leaf, all-gather + scatter
  cudaEvent_t ev[5] = {};
  ~fb_shard() {
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
    if (m2l_table) m2l_stream_table_free(m2l_table);
  }
};

void fb_shard_free(fb_shard *s) { delete s; }

template <class F>
static int guarded(F &&f) {
  try {
    f();
    return FB_OK;
  } catch (const fb::Error &e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return FB_ERR_CUDA;
  }
}

extern "C" {

int fb_partition_by_work(const double *work, size_t n_leaves, int parts, uint64_t *bounds_out) {
  if (!work || !bounds_out || parts < 1) return FB_ERR_INVALID_ARGUMENT;
  partition_by_work(work, n_leaves, parts, bounds_out);
  return FB_OK;
}

int fb_comm_unique_id(uint8_t *id_out128) {
  return guarded([&] {
    FB_REQUIRE(id_out128, "null argument");
    require_nccl();
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    FB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id_out128, &id, sizeof(id));
  });
}

int fb_comm_init(const uint8_t *id128, int rank, int world_size, fb_comm **out) {
  if (!out) return FB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  fb_comm *c = nullptr;
  const int rc = guarded([&] {
    FB_REQUIRE(id128 && world_size >= 1 && rank >= 0 && rank < world_size, "fb_comm_init: bad rank / world size");
    require_nccl();
    c = new fb_comm();
    c->rank = rank;
    c->world = world_size;
    FB_CUDA(cudaGetDevice(&c->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    FB_NCCL(nccl().CommInitRank(&c->comm, world_size, id, rank));
    FB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    FB_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    FB_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  });
  if (rc != FB_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return FB_OK;
}

void fb_comm_free(fb_comm *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  delete c;
}

int fb_comm_rank(const fb_comm *c) { return c ? c->rank : -1; }
int fb_comm_world_size(const fb_comm *c) { return c ? c->world : -1; }

int fb_tree_shard(fb_tree *t, fb_comm *comm) {
  return guarded([&] {
    FB_REQUIRE(t, "null tree");
    FB_CUDA(cudaSetDevice(t->device));
    if (t->shard) {
      fb_shard_free(t->shard);
      t->shard = nullptr;
    }
    if (!comm) return;
    FB_REQUIRE(comm->device == t->device, "the communicator and the tree live on different devices");
    std::unique_ptr<fb_shard> sh(new fb_shard());
    sh->comm = comm;
    const HostTree &ht = t->ht;
    const size_t nl = ht.leaves.size();
    std::vector<uint64_t> leaf_ptr(nl + 1);
    std::vector<double> work(nl);
    FB_REQUIRE(fb_tree_leaf_work(t, leaf_ptr.data(), work.data()) == FB_OK, "leaf work");
    sh->leaf_bounds.resize(comm->world + 1);
    partition_by_work(work.data(), nl, comm->world, sh->leaf_bounds.data());
    sh->pos.resize(comm->world + 1);
    for (int r = 0; r <= comm->world; ++r) sh->pos[r] = (int)leaf_ptr[sh->leaf_bounds[r]];
    for (int r = 0; r < comm->world; ++r) sh->max_rows = std::max(sh->max_rows, (size_t)(sh->pos[r + 1] - sh->pos[r]));
    sh->d_pos.upload(sh->pos, t->stream);
    std::vector<int> owned;
    for (uint64_t l = sh->leaf_bounds[comm->rank]; l < sh->leaf_bounds[comm->rank + 1]; ++l) {
      const int c = ht.leaves[l];
      if (ht.pt_end[c] > ht.pt_begin[c]) owned.push_back(c);
    }
    // ancestors of a contiguous leaf range are one contiguous id range per level
    sh->level_lo.assign(ht.depth + 1, INT32_MAX);
    sh->level_hi.assign(ht.depth + 1, 0);
    for (int c0 : owned)
      for (int c = c0; c >= 0; c = ht.parent[c]) {
        const int lv = ht.level[c];
        if (c >= sh->level_lo[lv] && c < sh->level_hi[lv]) break;  // the rest of the chain is already inside
        sh->level_lo[lv] = std::min(sh->level_lo[lv], c);
        sh->level_hi[lv] = std::max(sh->level_hi[lv], c + 1);
      }
    sh->n_owned_leaves = (int)owned.size();
    sh->d_owned_leaves.upload(owned, t->stream);
    // owned rows (Morton order) as the target subset: cell flags, leaf tiles, fused W/X row map
    const size_t p0 = (size_t)sh->pos[comm->rank], cnt = (size_t)sh->pos[comm->rank + 1] - p0;
    FB_REQUIRE(cnt > 0, "a rank owns no points: more ranks than leaves with work");
    sh->d_rows.reserve(cnt);
    FB_LAUNCH(k_iota_u64, (unsigned)((cnt + 255) / 256), 256, 0, t->stream, t->d_perm.p, p0, cnt, sh->d_rows.p);
    sh->ts = t->subset_target_set_dev(sh->d_rows.p, cnt, sh->tb);
    for (auto &e : sh->ev) FB_CUDA(cudaEventCreate(&e));
    FB_CUDA(cudaStreamSynchronize(t->stream));
    t->shard = sh.release();
  });
}

int fb_tree_shard_rows(const fb_tree *t, int rank, uint64_t *begin_pos, uint64_t *end_pos) {
  if (!t || !t->shard || rank < 0 || rank >= t->shard->comm->world) return FB_ERR_INVALID_ARGUMENT;
  if (begin_pos) *begin_pos = (uint64_t)t->shard->pos[rank];
  if (end_pos) *end_pos = (uint64_t)t->shard->pos[rank + 1];
  return FB_OK;
}

// one partitioned matvec; weights (all N rows, replicated) were uploaded with fb_tree_upload_weights
int fb_tree_matvec_sharded(fb_tree *t) {
  return guarded([&] {
    FB_REQUIRE(t && t->shard, "fb_tree_shard must be called first");
    FB_CUDA(cudaSetDevice(t->device));
    FB_REQUIRE(t->d_w_user.cap >= t->n * (size_t)t->nrhs, "fb_tree_upload_weights must be called first");
    fb_shard &sh = *t->shard;
    fb_comm &cm = *sh.comm;
    cudaStream_t s = t->stream;
    const size_t nc = t->ht.ncells();
    FB_CUDA(cudaEventRecord(sh.ev[0], s));
    t->sort_weights();
    t->upward(sh.d_owned_leaves.p, sh.n_owned_leaves, sh.ts.cell_flag);
    FB_CUDA(cudaEventRecord(sh.ev[1], s));
    // exchange 1 under the near-field pass
    const bool fuse = t->ht.adaptive && t->n_x_cells > 0 && sh.ts.row_of_pos != nullptr;
    const bool p2p_first = fuse || t->n_w_entries == 0;  // the near-field kernel then needs no multipoles
    const size_t mult_count = nc * (size_t)t->nrhs * coef_stride(t->P);
    if (cm.world > 1) {
      FB_CUDA(cudaEventRecord(cm.ev_ready, s));
      FB_CUDA(cudaStreamWaitEvent(cm.stream, cm.ev_ready, 0));
      FB_NCCL(nccl().AllReduce(t->d_mult.p, t->d_mult.p, mult_count, ncclDouble, ncclSum, cm.comm, cm.stream));
      FB_CUDA(cudaEventRecord(cm.ev_done, cm.stream));
      g_launches.fetch_add(1);
    }
    if (p2p_first) {
      t->d_out.zero(std::max(sh.ts.m, sh.max_rows) * (size_t)t->nrhs, s);
      if (t->timing) FB_CUDA(cudaEventRecord(t->ev[10], s));
      t->launch_p2p(sh.ts, false, fuse, s, false);
      if (t->timing) FB_CUDA(cudaEventRecord(t->ev[11], s));
    }
    if (cm.world > 1) FB_CUDA(cudaStreamWaitEvent(s, cm.ev_done, 0));
    FB_CUDA(cudaEventRecord(sh.ev[2], s));
    if (t->m2l_plan && (!sh.m2l_table || m2l_stream_table_nrhs(sh.m2l_table) != t->nrhs)) {
      if (sh.m2l_table) m2l_stream_table_free(sh.m2l_table);
      sh.m2l_table = nullptr;
      sh.m2l_table = m2l_stream_table_new(t->m2l_plan, t->nrhs, sh.level_lo.data(), sh.level_hi.data(), s);
    }
    if (p2p_first) {
      t->downward(sh.ts.cell_flag, fuse ? &sh.ts : nullptr, true, false, sh.m2l_table);
      if (t->timing) FB_CUDA(cudaEventRecord(t->ev[7], s));
      t->launch_l2p(sh.ts, false);
      if (t->timing) FB_CUDA(cudaEventRecord(t->ev[8], s));
    } else {
      t->d_out.reserve(sh.max_rows * (size_t)t->nrhs);
      t->evaluate_sources_fused(sh.ts);
    }
    FB_CUDA(cudaEventRecord(sh.ev[3], s));
    // exchange 2: owned rows -> full result on every rank
    const size_t share = sh.max_rows * (size_t)t->nrhs;
    sh.d_gather.reserve(share * cm.world);
    sh.d_full.reserve(t->n * (size_t)t->nrhs);
    if (cm.world > 1) {
      FB_NCCL(nccl().AllGather(t->d_out.p, sh.d_gather.p, share, ncclDouble, cm.comm, s));
      g_launches.fetch_add(1);
    } else {
      FB_CUDA(cudaMemcpyAsync(sh.d_gather.p, t->d_out.p, share * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    FB_LAUNCH(k_scatter_gathered, (unsigned)((t->n + 255) / 256), 256, 0, s, sh.d_gather.p, sh.d_pos.p, cm.world,
              sh.max_rows, t->nrhs, t->d_perm.p, t->n, sh.d_full.p);
    FB_CUDA(cudaEventRecord(sh.ev[4], s));
    FB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 4; ++k) {
      float ms = 0;
      FB_CUDA(cudaEventElapsedTime(&ms, sh.ev[k], sh.ev[k + 1]));
      sh.last_ms[k] = ms;
    }
    t->last_out_rows = t->n;
    if (t->timing && p2p_first) {  // per-kernel times of this rank's share, same slots as fb_tree_last_timing
      auto ms = [&](int a, int b) {
        float v = 0;
        cudaEventElapsedTime(&v, t->ev[a], t->ev[b]);
        return (double)v;
      };
      t->last_ms[0] = ms(0, 1);
      t->last_ms[1] = ms(1, 2);
      t->last_ms[2] = ms(3, 4);
      t->last_ms[3] = ms(4, 5);
      t->last_ms[4] = ms(5, 6);
      t->last_ms[5] = ms(7, 8);
      t->last_ms[6] = ms(10, 11);
      t->last_ms[7] = sh.last_ms[0] + sh.last_ms[1] + sh.last_ms[2] + sh.last_ms[3];
    }
  });
}

int fb_tree_sharded_timing(const fb_tree *t, double *ms_out4) {
  if (!t || !t->shard || !ms_out4) return FB_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < 4; ++k) ms_out4[k] = t->shard->last_ms[k];
  return FB_OK;
}

int fb_tree_sharded_result_device(const fb_tree *t, const double **dev_ptr) {
  if (!t || !t->shard || !dev_ptr) return FB_ERR_INVALID_ARGUMENT;
  *dev_ptr = t->shard->d_full.p;
  return FB_OK;
}

int fb_tree_sharded_download(fb_tree *t, double *out_vals) {
  return guarded([&] {
    FB_REQUIRE(t && t->shard && out_vals, "null argument");
    FB_CUDA(cudaSetDevice(t->device));
    FB_CUDA(cudaMemcpyAsync(out_vals, t->shard->d_full.p, t->n * (size_t)t->nrhs * sizeof(double), cudaMemcpyDeviceToHost,
                            t->stream));
    FB_CUDA(cudaStreamSynchronize(t->stream));
  });
}

}  // extern "C"
