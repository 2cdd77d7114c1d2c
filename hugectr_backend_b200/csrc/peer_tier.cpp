// NVLink tier (include/hpsx.h hpsx_cache_peer_tier_*): the rows of the page-locked host tables of a model,
// sharded over the HBM of the GPUs of one box, read one-sidedly by the direct-pull kernels of every replica.
//
// The reference deploys one full embedding cache per device and sends every cache miss of every device to the ONE host
// parameter server (hps_backend/src/model_state.cpp:395-419, include/backend.hpp:70-74).  Measured on this pool
// (profiles/pcie_conc_r02.txt): the host fabric gives one GPU 49-55 GB/s but four GPUs 29 GB/s and eight 21-37 GB/s each —
// plain pinned cudaMemcpyAsync drops just like the zero-copy gather, so no host-side layout fixes it.  The tier takes
// the misses of a replica off that fabric: rank r keeps the rows with owner_of(key, world) == r in its HBM, every rank
// maps every shard, and the HBM index that resolves a missed key to a row address (kernels.cu index_find) holds
// addresses inside the shards.  pull_binned_kernel / pull_misses_kernel are unchanged — the same loads now travel over
// NVLink (or stay in local HBM).  The host table stays the source of truth: keys past a shard's capacity, and
// everything after hpsx_ps_update_database_per_model until the tier is rebuilt, are pulled from host memory as before.
#include <cmath>
#include <cstring>

#include "engine_internal.hpp"

using namespace hpsx;
using namespace hpsx::eng;

namespace {

uint64_t shard_capacity(size_t rows, uint32_t world) {
  if (world <= 1) return std::max<size_t>(rows, 1);
  const double per = static_cast<double>(rows) / world;
  // owner_of() hashes uniformly: a shard deviates from rows/world by a few sqrt(rows/world)
  return static_cast<uint64_t>(per + 8.0 * std::sqrt(per) + 1024.0);
}

struct TierLocks {
  std::unique_lock<std::mutex> ws;
  std::unique_lock<std::shared_mutex> pull, rw;
  explicit TierLocks(hpsx_cache* c) : ws(c->async_mu), pull(c->pull_rw), rw(c->rw) {}
};

}  // namespace

namespace hpsx {
namespace eng {

void tier_release(hpsx_cache* c) {
  PeerTier& tr = c->tier;
  for (size_t t = 0; t < tr.peers.size(); ++t)
    for (size_t p = 0; p < tr.peers[t].size(); ++p)
      if (tr.peers[t][p].ipc && tr.peers[t][p].base) cudaIpcCloseMemHandle(tr.peers[t][p].base);
  for (PeerTier::Shard& s : tr.own)
    if (s.base) cudaFree(s.base);
  tr = PeerTier();
}

int tier_build_locked(hpsx_cache* c, uint32_t rank, uint32_t world) {
  if (!c->direct_pull)
    return fail(HPSX_ERR_UNSUPPORTED, "the NVLink tier needs enable_pagelock (it extends the direct pull of cache misses)");
  if (world == 0 || world > static_cast<uint32_t>(kMaxPeers) || rank >= world)
    return fail(HPSX_ERR_INVALID_ARG, "peer tier: rank/world out of range");
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  Model* model = c->model;
  const size_t T = model->tables.size();
  PeerTier& tr = c->tier;
  if (tr.world != 0 && (tr.world != world || tr.rank != rank || tr.own.size() != T)) {
    if (tr.committed) {
      const int rc = sync_direct_pull_index(c, c->async_stream);
      if (rc != HPSX_OK) return rc;
    }
    tier_release(c);
  }
  if (tr.committed) {  // the index must not point into shards that are about to be rewritten
    const int rc = sync_direct_pull_index(c, c->async_stream);
    if (rc != HPSX_OK) return rc;
    tr.committed = false;
    tr.repointed = 0;
  }
  tr.rank = rank;
  tr.world = world;
  tr.own.resize(T);
  // mappings of other ranks' shards are dropped: their owners may have rebuilt them too
  for (size_t t = 0; t < tr.peers.size(); ++t)
    for (size_t p = 0; p < tr.peers[t].size(); ++p)
      if (tr.peers[t][p].ipc && tr.peers[t][p].base) cudaIpcCloseMemHandle(tr.peers[t][p].base);
  tr.peers.assign(T, std::vector<PeerTier::Shard>(world));

  constexpr size_t kChunk = 1 << 20;
  int64_t* d_keys = nullptr;
  uint64_t* d_addrs = nullptr;
  unsigned long long* d_count = nullptr;
  HPSX_CU(cudaMalloc(&d_keys, kChunk * sizeof(int64_t)));
  HPSX_CU(cudaMalloc(&d_addrs, kChunk * sizeof(uint64_t)));
  HPSX_CU(cudaMalloc(&d_count, sizeof(unsigned long long)));
  cudaStream_t stream = c->async_stream;
  int rc = HPSX_OK;
  for (size_t t = 0; t < T && rc == HPSX_OK; ++t) {
    const HostTable& ht = *model->tables[t];
    const size_t dim = ht.dim();
    const unsigned long long device_rows = model->device_rows[t];  // != 0: generated here, no host copy
    const uint64_t cap = shard_capacity(device_rows != 0 ? static_cast<size_t>(device_rows) : ht.rows(), world);
    PeerTier::Shard& own = tr.own[t];
    cudaError_t ce = cudaSuccess;
    if (own.base == nullptr || own.cap < cap) {
      if (own.base) cudaFree(own.base);
      own = PeerTier::Shard();
      ce = cudaMalloc(&own.base, PeerTier::Shard::bytes(cap, dim));
      if (ce == cudaSuccess) own.cap = cap;
    }
    int64_t* shard_keys = reinterpret_cast<int64_t*>(own.base);
    float* shard_rows = reinterpret_cast<float*>(own.base + PeerTier::Shard::rows_offset(own.cap));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream);
    if (ce == cudaSuccess && device_rows != 0) {
      ce = launch_tier_fill_procedural(device_rows, model->device_seed[t], rank, world, dim, shard_keys, shard_rows, own.cap,
                                       d_count, stream);
    } else if (ce == cudaSuccess) {
      ht.export_rows(kChunk, [&](const int64_t* k, const uint64_t* a, size_t n) {
        if (ce != cudaSuccess) return;
        ce = cudaMemcpyAsync(d_keys, k, n * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_addrs, a, n * sizeof(uint64_t), cudaMemcpyHostToDevice, stream);
        if (ce == cudaSuccess)
          ce = launch_tier_fill(d_keys, d_addrs, n, rank, world, dim, shard_keys, shard_rows, own.cap, d_count, stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);  // the pageable vectors are reused
      });
    }
    unsigned long long count = 0;
    // (the stream is non-blocking: a plain cudaMemcpy would not wait for the fill kernel)
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
    if (ce == cudaSuccess) ce = cudaMemcpy(&count, d_count, sizeof(count), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) {
      rc = fail(HPSX_ERR_CUDA, std::string("building the NVLink tier shard: ") + cudaGetErrorString(ce));
      break;
    }
    own.rows = std::min<uint64_t>(count, own.cap);
    tr.peers[t][rank] = own;
    tr.peers[t][rank].ipc = false;
  }
  cudaFree(d_keys);
  cudaFree(d_addrs);
  cudaFree(d_count);
  if (rc != HPSX_OK) tier_release(c);
  return rc;
}

int tier_attach_local_locked(hpsx_cache* c, size_t table, uint32_t peer, hpsx_cache* pc) {
  PeerTier& tr = c->tier;
  if (tr.world == 0) return fail(HPSX_ERR_INVALID_ARG, "peer tier: build it first");
  if (!pc || table >= tr.peers.size() || peer >= tr.world || table >= pc->tier.own.size() || pc->tier.own[table].base == nullptr)
    return fail(HPSX_ERR_INVALID_ARG, "peer tier: the peer has no shard of that table");
  if (pc->tier.world != tr.world || pc->tier.rank != peer)
    return fail(HPSX_ERR_INVALID_ARG, "peer tier: the peer cache was built as another rank / world");
  if (pc->device != c->device) {
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
    int can = 0;
    HPSX_CU(cudaDeviceCanAccessPeer(&can, c->device, pc->device));
    if (!can)
      return fail(HPSX_ERR_UNSUPPORTED, "peer tier: device " + std::to_string(c->device) + " cannot map the memory of device " +
                                            std::to_string(pc->device));
    const cudaError_t e = cudaDeviceEnablePeerAccess(pc->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled)
      cudaGetLastError();
    else if (e != cudaSuccess)
      return fail(HPSX_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
  }
  PeerTier::Shard s = pc->tier.own[table];
  s.ipc = false;
  tr.peers[table][peer] = s;
  return HPSX_OK;
}

int tier_commit_locked(hpsx_cache* c) {
  PeerTier& tr = c->tier;
  if (tr.world == 0) return fail(HPSX_ERR_INVALID_ARG, "peer tier: build it first");
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  for (size_t t = 0; t < tr.peers.size(); ++t)
    for (uint32_t p = 0; p < tr.world; ++p)
      if (tr.peers[t][p].base == nullptr)
        return fail(HPSX_ERR_INVALID_ARG, "peer tier: the shard of rank " + std::to_string(p) + " (table " + std::to_string(t) +
                                              ") is not attached");
  unsigned long long* d_n = nullptr;
  HPSX_CU(cudaMalloc(&d_n, sizeof(unsigned long long)));
  cudaStream_t stream = c->async_stream;
  cudaError_t ce = cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), stream);
  unsigned long long inserted = 0;
  for (size_t t = 0; t < tr.peers.size() && ce == cudaSuccess; ++t) {
    const size_t dim = c->tables[t].dim;
    const bool device_table = c->model->device_rows[t] != 0;
    if (device_table) {
      // no host table behind it: the index is built from the shards themselves
      unsigned long long total = 0;
      for (uint32_t p = 0; p < tr.world; ++p) total += tr.peers[t][p].rows;
      uint64_t cap = 1024;
      while (cap < 2 * total) cap <<= 1;
      if (c->indexes.size() <= t) c->indexes.resize(t + 1, nullptr);
      if (c->indexes[t] == nullptr || c->tables[t].index_mask + 1 != cap) {
        if (c->indexes[t] != nullptr) cudaFree(c->indexes[t]);
        c->indexes[t] = nullptr;
        c->tables[t].index = nullptr;
        ce = cudaMalloc(&c->indexes[t], cap * sizeof(IndexSlot));
        if (ce != cudaSuccess) break;
      }
      c->tables[t].index = c->indexes[t];
      c->tables[t].index_mask = cap - 1;
      ce = launch_index_clear(c->indexes[t], cap, stream);
      inserted += total;
    }
    for (uint32_t p = 0; p < tr.world && ce == cudaSuccess; ++p) {
      const PeerTier::Shard& s = tr.peers[t][p];
      const int64_t* keys = reinterpret_cast<const int64_t*>(s.base);
      const float* rows = reinterpret_cast<const float*>(s.base + PeerTier::Shard::rows_offset(s.cap));
      ce = device_table ? launch_index_insert_shard(c->indexes[t], c->tables[t].index_mask, keys, rows, s.rows, dim, stream)
                        : launch_index_repoint(c->indexes[t], c->tables[t].index_mask, keys, rows, s.rows, dim, d_n, stream);
    }
  }
  unsigned long long n = 0;
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
  if (ce == cudaSuccess) ce = cudaMemcpy(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost);
  cudaFree(d_n);
  if (ce != cudaSuccess) return fail(HPSX_ERR_CUDA, std::string("peer tier commit: ") + cudaGetErrorString(ce));
  tr.repointed = n + inserted;
  tr.committed = true;
  return HPSX_OK;
}

int tier_connect_local_locked(Model* m, const std::vector<hpsx_cache*>& caches) {
  (void)m;
  const uint32_t world = static_cast<uint32_t>(caches.size());
  if (world == 0) return HPSX_OK;
  for (uint32_t r = 0; r < world; ++r) {
    const int rc = tier_build_locked(caches[r], r, world);
    if (rc != HPSX_OK) return rc;
  }
  for (uint32_t r = 0; r < world; ++r) {
    for (uint32_t p = 0; p < world; ++p) {
      if (p == r) continue;
      for (size_t t = 0; t < caches[r]->tables.size(); ++t) {
        const int rc = tier_attach_local_locked(caches[r], t, p, caches[p]);
        if (rc != HPSX_OK) return rc;
      }
    }
    const int rc = tier_commit_locked(caches[r]);
    if (rc != HPSX_OK) return rc;
  }
  return HPSX_OK;
}

}  // namespace eng
}  // namespace hpsx

extern "C" {

int hpsx_cache_peer_tier_build(hpsx_cache* c, uint32_t rank, uint32_t world) {
  HPSX_GUARD_BEGIN
  if (!c) return fail(HPSX_ERR_INVALID_ARG, "null cache");
  TierLocks lk(c);
  return tier_build_locked(c, rank, world);
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_export(hpsx_cache* c, size_t table, void* handle64, uint64_t* rows, uint64_t* cap) {
  HPSX_GUARD_BEGIN
  if (!c || !handle64 || !rows || !cap) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  std::shared_lock<std::shared_mutex> lk(c->pull_rw);
  if (table >= c->tier.own.size() || c->tier.own[table].base == nullptr)
    return fail(HPSX_ERR_INVALID_ARG, "peer tier: no shard of that table (build first)");
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == HPSX_SHARD_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  HPSX_CU(cudaIpcGetMemHandle(&h, c->tier.own[table].base));
  std::memcpy(handle64, &h, sizeof(h));
  *rows = c->tier.own[table].rows;
  *cap = c->tier.own[table].cap;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_attach_ipc(hpsx_cache* c, size_t table, uint32_t peer, const void* handle64, uint64_t rows,
                                    uint64_t cap) {
  HPSX_GUARD_BEGIN
  if (!c || !handle64) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  TierLocks lk(c);
  PeerTier& tr = c->tier;
  if (tr.world == 0 || table >= tr.peers.size() || peer >= tr.world || peer == tr.rank || rows > cap)
    return fail(HPSX_ERR_INVALID_ARG, "peer tier: bad table / peer (build first; a rank does not attach itself)");
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  if (tr.committed) {
    const int rc = sync_direct_pull_index(c, c->async_stream);
    if (rc != HPSX_OK) return rc;
    tr.committed = false;
    tr.repointed = 0;
  }
  PeerTier::Shard& s = tr.peers[table][peer];
  if (s.ipc && s.base) cudaIpcCloseMemHandle(s.base);
  s = PeerTier::Shard();
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  HPSX_CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s.base = static_cast<unsigned char*>(p);
  s.rows = rows;
  s.cap = cap;
  s.ipc = true;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_attach_local(hpsx_cache* c, uint32_t peer, hpsx_cache* peer_cache) {
  HPSX_GUARD_BEGIN
  if (!c || !peer_cache || c == peer_cache) return fail(HPSX_ERR_INVALID_ARG, "peer tier: bad peer cache");
  TierLocks lk(c);
  std::shared_lock<std::shared_mutex> plk(peer_cache->pull_rw);
  if (c->tier.committed) {
    const int rc = sync_direct_pull_index(c, c->async_stream);
    if (rc != HPSX_OK) return rc;
    c->tier.committed = false;
    c->tier.repointed = 0;
  }
  for (size_t t = 0; t < c->tables.size(); ++t) {
    const int rc = tier_attach_local_locked(c, t, peer, peer_cache);
    if (rc != HPSX_OK) return rc;
  }
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_commit(hpsx_cache* c) {
  HPSX_GUARD_BEGIN
  if (!c) return fail(HPSX_ERR_INVALID_ARG, "null cache");
  TierLocks lk(c);
  return tier_commit_locked(c);
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_detach(hpsx_cache* c) {
  HPSX_GUARD_BEGIN
  if (!c) return fail(HPSX_ERR_INVALID_ARG, "null cache");
  TierLocks lk(c);
  if (c->tier.world == 0) return HPSX_OK;
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  int rc = HPSX_OK;
  if (c->tier.committed) rc = sync_direct_pull_index(c, c->async_stream);
  tier_release(c);
  return rc;
  HPSX_GUARD_END
}

int hpsx_cache_peer_tier_info(hpsx_cache* c, hpsx_peer_tier_info* out) {
  HPSX_GUARD_BEGIN
  if (!c || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  std::shared_lock<std::shared_mutex> lk(c->pull_rw);
  std::memset(out, 0, sizeof(*out));
  const PeerTier& tr = c->tier;
  out->rank = tr.rank;
  out->world = tr.world;
  out->committed = tr.committed ? 1 : 0;
  out->index_entries_in_tier = tr.repointed;
  for (size_t t = 0; t < tr.own.size(); ++t) {
    out->own_rows += tr.own[t].rows;
    if (tr.own[t].base) out->own_bytes += PeerTier::Shard::bytes(tr.own[t].cap, c->tables[t].dim);
  }
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_peer_tier_connect_local(hpsx_ps* ps, const char* model) {
  HPSX_GUARD_BEGIN
  if (!ps || !model) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  Model* m = nullptr;
  {
    std::lock_guard<std::mutex> lk(ps->mu);
    auto it = ps->models.find(model);
    if (it != ps->models.end()) m = it->second.get();
  }
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + model + "'");
  std::vector<hpsx_cache*> caches;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    for (auto& kv : m->caches) caches.push_back(kv.second.get());  // std::map: ascending device order
  }
  if (caches.empty()) return fail(HPSX_ERR_NOT_FOUND, std::string("model '") + model + "' has no embedding cache");
  // lock order everywhere: async_mu, then pull_rw, then rw; caches in ascending device order
  std::vector<std::unique_lock<std::mutex>> ws;
  std::vector<std::unique_lock<std::shared_mutex>> held;
  for (hpsx_cache* c : caches) ws.emplace_back(c->async_mu);
  for (hpsx_cache* c : caches) held.emplace_back(c->pull_rw);
  for (hpsx_cache* c : caches) held.emplace_back(c->rw);
  return tier_connect_local_locked(m, caches);
  HPSX_GUARD_END
}

}  // extern "C"
