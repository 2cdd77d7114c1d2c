// Model-parallel groups (include/hpsx.h: hpsx_shard_group_*): the fused exchange over
// NVLink peer memory.  Kernels: kernels.cu (shard_dispatch, shard_signal_wait, probe_gather_inbox,
// shard_scatter_stage); the miss path reuses the session's machinery (hpsx.cpp).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine_internal.hpp"

using namespace hpsx;
using namespace hpsx::eng;

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void shard_fill_peers(hpsx_shard_group* g) {
  using G = hpsx_shard_group;
  for (uint32_t p = 0; p < g->world; ++p) {
    unsigned char* base = g->peer_arena[p];
    uint32_t* ctrl = reinterpret_cast<uint32_t*>(base);
    // the slot / cells of rank p that belong to THIS rank
    g->peers.inbox_keys[p] = reinterpret_cast<int64_t*>(base + g->off_keys) + static_cast<size_t>(g->rank) * g->slot_cap;
    g->peers.inbox_pos[p] = reinterpret_cast<uint32_t*>(base + g->off_pos) + static_cast<size_t>(g->rank) * g->slot_cap;
    g->peers.inbox_cnt[p] = ctrl + G::kCnt + g->rank;
    g->peers.flag_dispatch[p] = ctrl + G::kFlagDispatch + g->rank;
    g->peers.flag_return[p] = ctrl + G::kFlagReturn + g->rank;
    g->peers.out[p] = reinterpret_cast<float*>(base + g->off_out);
  }
  g->connected = true;
}

int shard_lookup(hpsx_shard_group* g, const int64_t* d_keys, size_t n, float** d_out) {
  using G = hpsx_shard_group;
  NvtxRange range("hpsx_shard_group_lookup");
  hpsx_session* s = g->s;
  hpsx_cache* c = s->cache;
  const size_t t = g->table;
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  const uint32_t seq = ++g->seq;
  const uint32_t epoch = c->epoch.fetch_add(1, std::memory_order_relaxed);
  const double tr0 = now_ms();
  uint32_t* ctrl = g->ctrl();
  uint32_t* d_status = ctrl + G::kStatus;
  uint32_t* d_miss_count = ctrl + G::kMissCount;
  const DeviceTable& dt = c->tables[t];
  ++s->stats.lookups;
  s->stats.keys += n;

  HPSX_CU(cudaMemsetAsync(ctrl + G::kCursor, 0, (G::kSeen + kMaxPeers - G::kCursor) * sizeof(uint32_t), s->stream));
  uint32_t m = 0, status = 0;
  bool returned = false;  // the speculative return wave ran on the device
  // the miss path below may pull rows from the page-locked host table: keep a database reload out for the whole call
  std::shared_lock<std::shared_mutex> pull_lock(c->pull_rw, std::defer_lock);
  if (c->direct_pull) pull_lock.lock();
  {
    std::shared_lock<std::shared_mutex> rlock(c->rw);
    HPSX_CU(launch_shard_dispatch(d_keys, n, g->world, g->peers, ctrl + G::kCursor, s->stream));
    HPSX_CU(launch_shard_signal_wait(g->peers, g->world, seq, 0, ctrl + G::kCursor, ctrl + G::kCnt,
                                     ctrl + G::kFlagDispatch, static_cast<uint32_t>(std::min<size_t>(g->miss_cap, 0xFFFFFFFFu)),
                                     d_status, g->timeout_ns, s->stream, nullptr, nullptr, ctrl + G::kSeen));
    HPSX_CU(cudaEventRecord(s->ev[2 * t], s->stream));
    HPSX_CU(launch_probe_gather_inbox(dt, g->peers, g->world, g->rank, g->slot_cap,
                                      reinterpret_cast<const int64_t*>(g->arena + g->off_keys),
                                      reinterpret_cast<const uint32_t*>(g->arena + g->off_pos), ctrl + G::kCnt, d_status,
                                      epoch, !c->is_static, d_miss_count, g->d_miss_pos, g->d_miss_keys, g->hd_miss_keys,
                                      n, s->stream));
    HPSX_CU(cudaEventRecord(s->ev[2 * t + 1], s->stream));
    // return wave, speculatively: it runs only if the gather recorded no miss (else the host resolves them below)
    HPSX_CU(launch_shard_signal_wait(g->peers, g->world, seq, 1, ctrl + G::kCursor, ctrl + G::kCnt, ctrl + G::kFlagReturn, 0,
                                     d_status, g->timeout_ns, s->stream, d_miss_count, ctrl + G::kDone, ctrl + G::kSeen));
    s->stats.kernel_launches += 4;
    HPSX_CU(cudaMemcpyAsync(g->h_ctrl, ctrl, G::kWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    status = g->h_ctrl[G::kStatus];
    returned = g->h_ctrl[G::kDone] != 0u;
    m = (status || returned) ? 0u : g->h_ctrl[G::kMissCount];
  }
  const double tr1 = now_ms();
  s->stats.d2h_bytes += G::kWords * sizeof(uint32_t);
  hpsx_shard_stats& st = g->last;
  st = hpsx_shard_stats{};
  for (uint32_t p = 0; p < g->world; ++p) {
    st.sent[p] = g->h_ctrl[G::kCursor + p];
    st.received[p] = status ? 0u : g->h_ctrl[G::kCnt + p];
    st.keys_received += st.received[p];
    if (p != g->rank) {
      st.keys_sent_remote += st.sent[p];
      st.keys_received_remote += st.received[p];
    }
  }
  st.misses = m;
  if (status == 0) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev[2 * t], s->ev[2 * t + 1]) == cudaSuccess) {
      s->stats.probe_kernel_ms += ms;
      ++s->stats.probe_kernel_launches;
      s->stats.probe_kernel_keys += st.keys_received;
    }
    s->stats.hits += st.keys_received - m;
  }

  int rc = HPSX_OK;
  if (m > 0) {
    // resolve the misses into a local stage, forward the rows to their requesters, insert them here
    rc = ensure_pool_stage(s, m);
    std::unique_lock<std::shared_mutex> wlock(c->rw, std::defer_lock);
    if (rc == HPSX_OK && c->direct_pull) {
      if (!c->is_static) wlock.lock();
      const bool use_sorted = m >= pull_sort_min() && m <= s->cap_keys;
      cudaError_t e = cudaMemsetAsync(s->d_counters + 2 * s->vt + t, 0, sizeof(uint32_t), s->stream);
      if (e == cudaSuccess && use_sorted) {
        rc = ensure_sort_workspace(s);
        if (rc == HPSX_OK)
          e = launch_resolve_and_sort_misses(dt, g->d_miss_keys, m, s->d_addr[0], s->d_sidx[0], s->d_addr[1],
                                             s->d_sidx[1], s->d_sort_temp, s->sort_temp_bytes, s->stream);
      }
      if (rc == HPSX_OK && e == cudaSuccess)
        e = launch_pull_misses(dt, g->d_miss_keys, g->d_miss_pos, d_miss_count, st.keys_received, nullptr,
                               s->d_pool_stage, !c->is_static, 1, 0.f, epoch, s->d_counters + s->vt + t,
                               s->d_counters + 2 * s->vt + t, use_sorted ? s->d_addr[1] : nullptr,
                               use_sorted ? s->d_sidx[1] : nullptr, m, s->stream);
      if (rc == HPSX_OK && e != cudaSuccess) rc = fail(HPSX_ERR_CUDA, std::string("direct pull: ") + cudaGetErrorString(e));
      s->stats.kernel_launches += use_sorted ? 2 : 1;
      s->stats.misses += m;
      s->stats.h2d_bytes += static_cast<uint64_t>(m) * g->dim * sizeof(float);
    } else if (rc == HPSX_OK) {
      const MissBufs mb{g->h_miss_keys, g->d_miss_keys, g->d_miss_pos};
      s->stats.d2h_bytes += static_cast<uint64_t>(m) * sizeof(int64_t);
      rc = stream_miss_rows(s, t, 0, m, nullptr, false, epoch, s->d_pool_stage, nullptr, &mb);
      if (rc == HPSX_OK && !c->is_static) {
        wlock.lock();
        const cudaError_t e = launch_insert_merge(dt, g->d_miss_keys, nullptr, s->d_pool_stage, m, nullptr, true, epoch,
                                                  s->d_counters + s->vt + t, s->stream);
        if (e != cudaSuccess) rc = fail(HPSX_ERR_CUDA, std::string("insert: ") + cudaGetErrorString(e));
        ++s->stats.kernel_launches;
      }
    }
    if (rc == HPSX_OK) {
      const cudaError_t e = launch_shard_scatter_stage(s->d_pool_stage, g->d_miss_pos, m, g->dim, g->peers, g->world, s->stream);
      if (e != cudaSuccess) rc = fail(HPSX_ERR_CUDA, std::string("scatter: ") + cudaGetErrorString(e));
      ++s->stats.kernel_launches;
    }
    if (rc != HPSX_OK) {
      // the peers still wait for this rank's return flag: raise the error bit and fall through
      const uint32_t one = 1;
      cudaMemcpyAsync(d_status, &one, sizeof(one), cudaMemcpyHostToDevice, s->stream);
    }
    if (wlock.owns_lock()) cudaStreamSynchronize(s->stream);  // slots are rewritten under the exclusive lock only
  }
  const double tr2 = now_ms();
  const std::string keep = g_err;
  if (!returned) {
    // after a timeout nobody is listening any more: publish, do not wait again
    const unsigned long long wait_ns = (status & 2u) ? 0ull : g->timeout_ns;
    HPSX_CU(launch_shard_signal_wait(g->peers, g->world, seq, 1, ctrl + G::kCursor, ctrl + G::kCnt, ctrl + G::kFlagReturn, 0,
                                     d_status, wait_ns, s->stream));
    ++s->stats.kernel_launches;
    HPSX_CU(cudaMemcpyAsync(g->h_ctrl + G::kStatus, d_status, (G::kSeen + kMaxPeers - G::kStatus) * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    status |= g->h_ctrl[G::kStatus];
  }
  st.status = status;
  if (trace_on())
    std::fprintf(stderr, "[hpsx] shard lookup rank %u n=%zu: dispatch+wait+gather+sync %.3f ms | misses %u: %.3f ms | return wave %.3f ms\n",
                 g->rank, n, tr1 - tr0, m, tr2 - tr1, now_ms() - tr2);
  if (rc != HPSX_OK) return fail(rc, keep);
  if (status != 0) {
    std::string seen;
    if (status & 2u) {
      // which peer was late, and by how much: its flag cell still held an older sequence number
      seen = " (this rank is at sequence " + std::to_string(seq) + "; last flag seen per peer:";
      for (uint32_t p = 0; p < g->world; ++p) seen += " " + std::to_string(g->h_ctrl[G::kSeen + p] >> 1);
      seen += ")";
    }
    std::string why = (status & 2u) ? "a rank did not arrive before the timeout" + seen
                      : (status & 4u) ? "this rank received more keys than its miss list can hold"
                                      : "another rank of the group reported a failure";
    return fail(HPSX_ERR_INTERNAL, "model-parallel lookup failed: " + why);
  }
  if (d_out) *d_out = reinterpret_cast<float*>(g->arena + g->off_out);
  return HPSX_OK;
}

}  // namespace

hpsx_shard_group::~hpsx_shard_group() {
  if (!s || s->device < 0) return;
  DeviceGuard guard(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (uint32_t p = 0; p < peer_arena.size(); ++p)
    if (p != rank && peer_arena[p] != nullptr && peer_ipc[p]) cudaIpcCloseMemHandle(peer_arena[p]);
  cudaFree(arena);
  cudaFree(d_miss_pos);
  cudaFree(d_miss_keys);
  if (h_miss_keys) cudaFreeHost(h_miss_keys);
  if (h_ctrl) cudaFreeHost(h_ctrl);
}

extern "C" {

// ------------------------------------------------------------------------------------------------
// model-parallel group
// ------------------------------------------------------------------------------------------------
int hpsx_shard_group_create(hpsx_session* s, size_t table, uint32_t rank, uint32_t world, hpsx_shard_group** out,
                            void* handle64) {
  HPSX_GUARD_BEGIN
  if (!s || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (!s->cache) return fail(HPSX_ERR_UNSUPPORTED, "a model-parallel group needs a GPU session (gpucache = true)");
  if (table >= s->model->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (world == 0 || world > static_cast<uint32_t>(kMaxPeers) || rank >= world)
    return fail(HPSX_ERR_INVALID_ARG, "need rank < world <= " + std::to_string(kMaxPeers));
  const size_t cap = s->cap_per_table[table];
  if (cap == 0 || cap >= (1ull << kShardPosBits))
    return fail(HPSX_ERR_UNSUPPORTED, "keys per request of the sharded table must be in [1, 2^26)");
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  std::unique_ptr<hpsx_shard_group> g(new hpsx_shard_group());
  g->s = s;
  g->table = table;
  g->rank = rank;
  g->world = world;
  g->slot_cap = static_cast<uint32_t>(cap);
  g->dim = s->model->tables[table]->dim();
  g->off_keys = 4096;
  g->off_pos = align_up(g->off_keys + static_cast<size_t>(world) * cap * sizeof(int64_t), 512);
  g->off_out = align_up(g->off_pos + static_cast<size_t>(world) * cap * sizeof(uint32_t), 512);
  g->arena_bytes = g->off_out + cap * g->dim * sizeof(float);
  HPSX_CU(cudaMalloc(&g->arena, g->arena_bytes));
  HPSX_CU(cudaMemset(g->arena, 0, 4096));
  g->miss_cap = static_cast<size_t>(world) * cap;
  HPSX_CU(cudaMalloc(&g->d_miss_pos, g->miss_cap * sizeof(uint32_t)));
  HPSX_CU(cudaMalloc(&g->d_miss_keys, g->miss_cap * sizeof(int64_t)));
  if (!s->cache->direct_pull) {
    HPSX_CU(cudaHostAlloc(&g->h_miss_keys, g->miss_cap * sizeof(int64_t), cudaHostAllocMapped | cudaHostAllocPortable));
    HPSX_CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g->hd_miss_keys), g->h_miss_keys, 0));
  }
  HPSX_CU(cudaMallocHost(&g->h_ctrl, hpsx_shard_group::kWords * sizeof(uint32_t)));
  // Nothing a lookup launches may be loaded lazily while a peer's flag-wait kernel spins (kernels.h): load this
  // library's kernels now, and run one tiny address sort so that the CUB kernels of the sorted pull are loaded too.
  HPSX_CU(cudaMemsetAsync(g->ctrl() + hpsx_shard_group::kCursor, 0,
                          (hpsx_shard_group::kSeen + kMaxPeers - hpsx_shard_group::kCursor) * sizeof(uint32_t), s->stream));
  HPSX_CU(cudaStreamSynchronize(s->stream));  // the driver's own memset kernel of that size is loaded now, too
  {
    // the stream-ordered allocator behind the miss stage creates its pool at first use
    const int prc = ensure_pool_stage(s, std::min<size_t>(cap, 4096));
    if (prc != HPSX_OK) return prc;
  }
  HPSX_CU(preload_shard_kernels());
  HPSX_CU(preload_miss_path_kernels());
  if (s->cache->direct_pull) {
    const int wrc = ensure_sort_workspace(s);
    if (wrc != HPSX_OK) return wrc;
    // CUB picks its kernels by problem size (one tile, or histogram + onesweep passes): sort once at each size
    for (size_t m : {static_cast<size_t>(1), std::min<size_t>(cap, 1u << 16)}) {
      HPSX_CU(cudaMemsetAsync(g->d_miss_keys, 0, m * sizeof(int64_t), s->stream));
      HPSX_CU(launch_resolve_and_sort_misses(s->cache->tables[table], g->d_miss_keys, m, s->d_addr[0], s->d_sidx[0],
                                             s->d_addr[1], s->d_sidx[1], s->d_sort_temp, s->sort_temp_bytes, s->stream));
    }
    HPSX_CU(cudaStreamSynchronize(s->stream));
  }
  g->peer_arena.assign(world, nullptr);
  g->peer_ipc.assign(world, false);
  g->peer_arena[rank] = g->arena;
  if (handle64) {
    cudaIpcMemHandle_t h;
    HPSX_CU(cudaIpcGetMemHandle(&h, g->arena));
    std::memcpy(handle64, &h, sizeof(h));
  }
  if (world == 1) shard_fill_peers(g.get());
  *out = g.release();
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_shard_group_connect_ipc(hpsx_shard_group* g, const void* all_handles) {
  HPSX_GUARD_BEGIN
  if (!g || !all_handles) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(g->s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  const unsigned char* hs = static_cast<const unsigned char*>(all_handles);
  for (uint32_t p = 0; p < g->world; ++p) {
    if (p == g->rank || g->peer_arena[p] != nullptr) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hs + static_cast<size_t>(p) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    HPSX_CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer_arena[p] = static_cast<unsigned char*>(ptr);
    g->peer_ipc[p] = true;
  }
  shard_fill_peers(g);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_shard_group_connect_local(hpsx_shard_group* g, hpsx_shard_group* const* groups) {
  HPSX_GUARD_BEGIN
  if (!g || !groups) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(g->s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  for (uint32_t p = 0; p < g->world; ++p) {
    if (p == g->rank) continue;
    const hpsx_shard_group* o = groups[p];
    if (!o || o->world != g->world || o->rank != p || o->slot_cap != g->slot_cap || o->dim != g->dim)
      return fail(HPSX_ERR_INVALID_ARG, "group " + std::to_string(p) + " does not match (world, rank, capacity, dim)");
    if (o->s->device != g->s->device) {
      int can = 0;
      HPSX_CU(cudaDeviceCanAccessPeer(&can, g->s->device, o->s->device));
      if (!can) return fail(HPSX_ERR_UNSUPPORTED, "no peer access between the devices of the group");
      const cudaError_t e = cudaDeviceEnablePeerAccess(o->s->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) HPSX_CU(e);
      cudaGetLastError();
    }
    g->peer_arena[p] = o->arena;
  }
  shard_fill_peers(g);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_shard_group_lookup(hpsx_shard_group* g, const int64_t* d_keys, size_t n, float** d_out) {
  HPSX_GUARD_BEGIN
  if (!g) return fail(HPSX_ERR_INVALID_ARG, "null group");
  if (!g->connected) return fail(HPSX_ERR_INVALID_ARG, "the group is not connected to its peers yet");
  if (n > g->slot_cap)
    return fail(HPSX_ERR_INVALID_ARG, std::to_string(n) + " keys exceed max_batch_size * maxnum_catfeature_query_per_table_per_sample = " +
                                          std::to_string(g->slot_cap));
  if (n > 0 && !d_keys) return fail(HPSX_ERR_INVALID_ARG, "null keys");
  std::lock_guard<std::mutex> lk(g->s->mu);
  return shard_lookup(g, d_keys, n, d_out);
  HPSX_GUARD_END
}

int hpsx_shard_group_get_stats(const hpsx_shard_group* g, hpsx_shard_stats* out) {
  if (!g || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = g->last;
  return HPSX_OK;
}

int hpsx_shard_group_capacity(const hpsx_shard_group* g, size_t* rows) {
  if (!g || !rows) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *rows = g->slot_cap;
  return HPSX_OK;
}

int hpsx_shard_group_set_timeout_ms(hpsx_shard_group* g, uint64_t ms) {
  if (!g || ms == 0) return fail(HPSX_ERR_INVALID_ARG, "null group / zero timeout");
  g->timeout_ns = ms * 1000000ull;
  return HPSX_OK;
}

int hpsx_shard_group_destroy(hpsx_shard_group* g) {
  HPSX_GUARD_BEGIN
  delete g;
  return HPSX_OK;
  HPSX_GUARD_END
}

}  // extern "C"
