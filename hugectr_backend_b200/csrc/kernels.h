// Launchers of the sm_100a kernels of the HPS lookup path.  Host C++ only sees this header.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "hpsx_common.h"

namespace hpsx {

// One table of the HBM embedding cache as the kernels see it.
struct DeviceTable {
  Bucket* buckets = nullptr;   // [num_buckets]
  float* values = nullptr;     // [num_buckets * kWays, dim] row-major, 512-B aligned base
  uint32_t num_buckets = 0;
  uint32_t dim = 0;            // floats per row
  float default_value = 0.f;
  // direct-pull mode (enable_pagelock): index of the whole host table, rows read over PCIe by kernels
  const IndexSlot* index = nullptr;  // [index_mask + 1]
  uint64_t index_mask = 0;
  const float* sentinel_row = nullptr;  // row of the key that doubles as the empty marker, if loaded
  uint32_t host_parts = 1;              // partitions of the host table = bins of the miss lists
};

// Binned miss list of ONE probe launch sequence against one table (DESIGN.md §3 "Miss path").  Bin b < num_bins holds
// the misses whose row lives in partition b of the host table — a window of <= 256 MiB of host memory, inside which
// zero-copy PCIe reads run at the link's full rate; bin num_bins is the spill list for entries that found their bin
// full.  Appended to by the probe kernels, consumed by launch_pull_binned / launch_insert_binned without any sort
// and without the host learning a count in between.
struct MissBins {
  uint32_t* count = nullptr;  // [num_bins + 1], zeroed by the caller; count[b] may run past bin_cap (consumers clamp)
  int64_t* keys = nullptr;    // [num_bins * bin_cap + spill_cap]
  uint32_t* pos = nullptr;    // same layout: destination row (| request << kShardPosBits when requests are merged)
  uint32_t num_bins = 0;      // partitions of the host table
  uint32_t bin_cap = 0;
  uint32_t spill_cap = 0;
};

constexpr uint32_t kMissSlot = 0xFFFFFFFFu;
constexpr uint32_t kSrcMissBit = 0x80000000u;  // pooled path: src index refers to the miss stage

enum ProbeVariant : int {
  kProbeLdg = 0,  // warp-per-32-keys, LDG.128 row copies through registers (any row size)
  kProbeTma = 1,  // cp.async.bulk row staging through a shared-memory ring (UBLKCP), dim*4 % 16 == 0
  kProbeV8 = 4,   // default: 256-bit row vectors with L2 evict_first, bucket keys kept in L2 (evict_last)
};

// K2+K3+K6 fused: probe the cache for keys[0..n), copy hit rows to out[i*dim..),
// write the default vector for misses, touch LRU stamps, and append (position,key) of every miss to
// the miss list (warp ballot + prefix, one atomic per warp tile).
// `miss_count` must be zeroed by the caller (stream-ordered).  `touch`=false for static caches.
// `hd_miss_keys` (nullable) is the device-visible address of a mapped pinned host buffer that
// receives a mirror of the miss keys, so the host needs no separate D2H copy of them.
cudaError_t launch_probe_gather(const DeviceTable& t, const int64_t* d_keys, size_t n, float* d_out,
                                uint32_t epoch, bool touch, uint32_t* d_miss_count,
                                uint32_t* d_miss_pos, int64_t* d_miss_keys, int64_t* hd_miss_keys,
                                int variant, cudaStream_t stream, const uint32_t* d_pos = nullptr,
                                uint32_t pos_base = 0, void* d_out_bf16 = nullptr, const MissBins* bins = nullptr,
                                bool skip_miss_rows = false);
// `bins` (nullable): misses are appended to the binned lists (d_miss_count / d_miss_pos / d_miss_keys are unused);
// with `skip_miss_rows` the kernel stores nothing for a missed key — launch_pull_binned writes that row.
// `d_out_bf16` (nullable, 16-B aligned): a bf16 mirror of d_out in the same row order, written by the same kernel
// (and by the miss kernels below), so the dense head reads bf16 activations without a conversion pass.
// `pos_base`: this launch covers keys [pos_base, pos_base + n) of a larger request whose chunks share one miss
// list (d_out / d_keys already point at the chunk; recorded miss positions are request-relative).
// With `d_pos`, key i is delivered to row d_pos[i] of d_out instead of row i, and d_out may be another
// GPU's buffer (NVLink peer mapping): the model-parallel return leg fused into the gather.

// Probe only (pooled path): d_src[i] = slot index on hit, kSrcMissBit | miss-list index on miss.
cudaError_t launch_probe_index(const DeviceTable& t, const int64_t* d_keys, size_t n, uint32_t epoch,
                               bool touch, uint32_t* d_src, uint32_t* d_miss_count,
                               uint32_t* d_miss_pos, int64_t* d_miss_keys, int64_t* hd_miss_keys,
                               cudaStream_t stream);

// K4+K5 fused: for every miss i in [0,m): row = stage[i*dim..); if d_out: out[pos[i]*dim..) = row
// (merge); with d_stage == nullptr the row is read back from out[pos[i]*dim..) and only inserted; if `insert`: put (key,row) into the cache — skip when present, else first empty way, else
// the way with the oldest stamp that was not touched in this epoch.  One warp per miss, per-bucket lock.
cudaError_t launch_insert_merge(const DeviceTable& t, const int64_t* d_miss_keys,
                                const uint32_t* d_miss_pos, const float* d_stage, size_t m,
                                float* d_out, bool insert, uint32_t epoch, uint32_t* d_inserted,
                                cudaStream_t stream, void* d_out_bf16 = nullptr);

// K9 (cache refresh): for every key that is still resident overwrite its cached row with d_stage[i*dim..);
// *d_updated (nullable) counts the rows rewritten.  Caller holds the cache's host lock exclusively.
cudaError_t launch_update_values(const DeviceTable& t, const int64_t* d_keys, const float* d_stage, size_t n,
                                 uint32_t* d_updated, cudaStream_t stream);

// Direct pull (K4+K5 without the CPU): for every miss i in [0, *d_miss_count): find the key in the
// HBM-resident index of the page-locked host table, read the row straight from mapped pinned host
// memory (zero-copy over PCIe), then merge it into out[pos[i]] (when `d_out` and the lookup runs in
// synchronous-insertion mode), keep it in stage[i] (when `d_stage`), and insert it into the cache
// (when `insert`).  Keys absent from the host table get the default vector and are counted in
// *d_absent.  The miss count stays on the device: no host round trip between probe and pull.
// insert_mode: 1 synchronous, 0 asynchronous semantics (misses keep the default vector in `out`),
// -1 decide on the device: synchronous iff 1 - m/n < hit_rate_threshold.
cudaError_t launch_pull_misses(const DeviceTable& t, const int64_t* d_miss_keys, const uint32_t* d_miss_pos,
                               const uint32_t* d_miss_count, size_t n_keys, float* d_out, float* d_stage,
                               bool insert, int insert_mode, float hit_rate_threshold, uint32_t epoch,
                               uint32_t* d_inserted, uint32_t* d_absent, const unsigned long long* d_sorted_addr,
                               const uint32_t* d_sorted_idx, size_t m_hint, cudaStream_t stream,
                               int max_ctas_per_sm = 0, void* d_out_bf16 = nullptr, int64_t* d_mark_absent = nullptr,
                               float* const* batch_outs = nullptr, int batch_count = 0);
// batch_outs/batch_count: the miss list merges the misses of `batch_count` requests of one table; entry i goes to
// row (pos & (2^26 - 1)) of batch_outs[pos >> 26] (d_out is ignored).
// d_mark_absent (= d_miss_keys, writable): keys missing from the host table are replaced by the empty marker, so
// that launch_insert_merge(d_stage = nullptr), which inserts the pulled rows from the output buffer, skips them.
// max_ctas_per_sm > 0 caps the persistent grid so that other kernels (the probes of later request chunks) keep
// SM resources while the pull waits on PCIe.

// Binned pull (the default miss path with enable_pagelock and synchronous insertion): for every entry of `bins`,
// walking the bins in ascending order so that the PCIe reads in flight stay inside one or two host-memory windows:
// find the key in the HBM index of the page-locked host table, read its row over PCIe and store it to row `pos` of
// `d_out` (of batch_outs[pos >> kShardPosBits] when requests are merged); a key that is not in the host table gets
// the default vector, is counted in *d_absent and is replaced by kEmptyKey in the list.  The cache is NOT touched:
// probes of later chunks of the request run beside this kernel.  Grid: `grid_ctas` CTAs, persistent.
cudaError_t launch_pull_binned(const DeviceTable& t, const MissBins& bins, float* d_out, void* d_out_bf16,
                               float* const* batch_outs, int batch_count, uint32_t* d_absent, int grid_ctas,
                               cudaStream_t stream, int insert = 0, uint32_t epoch = 0, uint32_t* d_inserted = nullptr,
                               int rows_in_flight = 1);
// rows_in_flight >= 4 (NVLink tier; rows of <= 32 vectors): every warp keeps four rows in flight instead of one.
// insert: the pulling warp also inserts the row (fused; HBM work hidden behind the PCIe reads).  Only when nothing
// probes the cache meanwhile: the caller holds the cache exclusively and every probe of the request has completed.
// Inserts the rows the binned pull delivered, reading them back from the output buffer (HBM to HBM); entries whose
// key was replaced by kEmptyKey are skipped.  Caller excludes probes (they would copy rows while slots are rewritten).
cudaError_t launch_insert_binned(const DeviceTable& t, const MissBins& bins, const float* d_out,
                                 float* const* batch_outs, int batch_count, uint32_t epoch, uint32_t* d_inserted,
                                 cudaStream_t stream);

// Locality for the host link: random 512-B reads over a multi-GB pinned table run at ~32-42 GB/s, the same
// reads in ascending address order at ~51 GB/s (tools/pcie_probe.cu: page-table / IOTLB reach).  So when the
// host knows the miss count `m`, the misses are first resolved to host addresses and radix-sorted by page
// number; launch_pull_misses then walks them in that order (d_sorted_addr / d_sorted_idx, m_hint = m).
size_t sort_misses_temp_bytes(size_t max_items);
cudaError_t launch_resolve_and_sort_misses(const DeviceTable& t, const int64_t* d_miss_keys, size_t m,
                                           unsigned long long* d_addr_tmp, uint32_t* d_idx_tmp,
                                           unsigned long long* d_addr_sorted, uint32_t* d_idx_sorted, void* d_temp,
                                           size_t temp_bytes, cudaStream_t stream);

// Adds (key -> host row address) pairs to a direct-pull index.  Slots must have been cleared with
// launch_index_clear.  Duplicated keys keep the last address written.
cudaError_t launch_index_clear(IndexSlot* slots, uint64_t capacity, cudaStream_t stream);
cudaError_t launch_index_build(IndexSlot* slots, uint64_t mask, const int64_t* d_keys,
                               const uint64_t* d_row_addrs, size_t n, cudaStream_t stream);

// NVLink tier: with several GPUs in the box the rows of a page-locked host table are also kept
// SHARDED over the GPUs' HBM — rank r holds the rows with owner_of(key, world) == r as [keys | rows] in one
// allocation that every rank maps (peer access / CUDA IPC) — and the direct-pull index of every rank points at
// those copies instead of at host memory.  The pull kernels are unchanged: a missed row is read with the same
// loads, over NVLink from the owner's HBM (or from local HBM) instead of over PCIe from host DRAM.  One-sided: the
// owner's SMs, streams and locks are not involved, replicas stay independent (no collective, no flags).
// launch_tier_fill: of n (key, device-visible host row address) pairs append those owned by `rank` to the shard
// (*d_count = entries so far, may run past `cap`: the excess stays host-only).
cudaError_t launch_tier_fill(const int64_t* d_keys, const uint64_t* d_row_addrs, size_t n, uint32_t rank, uint32_t world,
                             size_t dim, int64_t* shard_keys, float* shard_rows, unsigned long long cap,
                             unsigned long long* d_count, cudaStream_t stream);
// launch_index_repoint: index[shard_keys[i]].row = shard_rows + i * dim for the keys the index already holds;
// shard_keys / shard_rows may be a peer's memory.  *d_repointed (nullable) counts the entries changed.
cudaError_t launch_index_repoint(IndexSlot* slots, uint64_t mask, const int64_t* shard_keys, const float* shard_rows,
                                 unsigned long long n, size_t dim, unsigned long long* d_repointed, cudaStream_t stream);

// Tables that exist ONLY in the tier (model-parallel rows: no host copy, no local cache — BASELINE configs[3]):
// launch_tier_fill_procedural generates the shard on the device (keys [0, num_rows) owned by `rank`, synth_value rows),
// launch_index_insert_shard enters a shard's keys into the index (insert or overwrite), and launch_tier_gather serves a
// request: d_out[i] = row(d_keys[i]) read from the owner's shard (NVLink or local HBM), default vector + *d_absent
// (nullable) for keys in no shard.
cudaError_t launch_tier_fill_procedural(unsigned long long num_rows, unsigned long long seed, uint32_t rank, uint32_t world,
                                        size_t dim, int64_t* shard_keys, float* shard_rows, unsigned long long cap,
                                        unsigned long long* d_count, cudaStream_t stream);
cudaError_t launch_index_insert_shard(IndexSlot* slots, uint64_t mask, const int64_t* shard_keys, const float* shard_rows,
                                      unsigned long long n, size_t dim, cudaStream_t stream);
cudaError_t launch_tier_gather(const DeviceTable& t, const int64_t* d_keys, size_t n, float* d_out, uint32_t* d_absent,
                               cudaStream_t stream);

// a8: pooled[b*dim..) = sum_{j<hotness} row(src[b*hotness+j]) (mean: / hotness), ascending j, fp32.
// Rows come from the cache slab or (kSrcMissBit) from the staged miss rows.
cudaError_t launch_pooled_gather(const DeviceTable& t, const uint32_t* d_src, const float* d_stage,
                                 size_t num_bags, size_t hotness, bool mean, float* d_pooled,
                                 cudaStream_t stream);

cudaError_t launch_table_clear(const DeviceTable& t, cudaStream_t stream);
cudaError_t launch_count_resident(const DeviceTable& t, unsigned long long* d_count,
                                  cudaStream_t stream);
cudaError_t launch_dump_keys(const DeviceTable& t, int64_t* d_keys, unsigned long long cap,
                             unsigned long long* d_count, cudaStream_t stream);

// K1: dedup.  Workspace: ws_keys[cap] (int64), ws_ids[cap] (u32), cap = power of two >= 2n.
// d_counter[0] receives the number of unique keys; d_counter[1] is scratch for the sentinel key.
cudaError_t launch_unique(const int64_t* d_keys, size_t n, int64_t* ws_keys, uint32_t* ws_ids,
                          size_t cap, int64_t* d_unique, uint32_t* d_inverse, uint32_t* d_counter,
                          cudaStream_t stream);

// Multi-GPU routing: bucket keys by owner_of(key, num_shards).  d_counts[num_shards] zeroed inside;
// d_cursor[num_shards] is scratch.
cudaError_t launch_route_keys(const int64_t* d_keys, size_t n, uint32_t num_shards,
                              int64_t* d_routed_keys, uint32_t* d_perm, uint32_t* d_counts,
                              uint32_t* d_cursor, cudaStream_t stream);
cudaError_t launch_scatter_rows(const float* d_rows, const uint32_t* d_perm, size_t n, size_t dim,
                                float* d_out, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Fused model-parallel exchange over NVLink peer memory (hpsx_shard_group in hpsx.h).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;            // GPUs of one box
constexpr int kMaxBatchOuts = 16;        // requests whose misses one pull launch may serve (= kMaxBatchRequests)
constexpr uint32_t kShardPosBits = 26;   // miss destination = (requester << 26) | row in its output

// What rank r knows about every rank p of the group (entry r = its own arena).  Passed to kernels by value.
struct ShardPeers {
  int64_t* inbox_keys[kMaxPeers];   // p's inbox slot for keys sent by r: [slot_cap]
  uint32_t* inbox_pos[kMaxPeers];   // matching request positions
  uint32_t* inbox_cnt[kMaxPeers];   // cell of p's count array that r writes
  uint32_t* flag_dispatch[kMaxPeers];  // cell of p's dispatch-flag array that r writes
  uint32_t* flag_return[kMaxPeers];    // cell of p's return-flag array that r writes
  float* out[kMaxPeers];            // p's output buffer [slot_cap, dim]
};

// Step 1: bucket keys by owner_of(key, world) and store (key, position) into the owners' inboxes.
// d_cursor[kMaxPeers] (local, zeroed by the caller) receives the per-owner counts.
cudaError_t launch_shard_dispatch(const int64_t* d_keys, size_t n, uint32_t world, const ShardPeers& peers,
                                  uint32_t* d_cursor, cudaStream_t stream);
// Steps 2 and 4: publish (phase 0: counts + dispatch flag, phase 1: return flag with this rank's error bit)
// to every peer, then spin until every peer's flag for sequence `seq` has arrived in our own control block.
// *d_status: bit 0 error (ours or a peer's), bit 1 timeout, bit 2 more keys received than `capacity`.
cudaError_t launch_shard_signal_wait(const ShardPeers& peers, uint32_t world, uint32_t seq, int phase,
                                     const uint32_t* d_cursor, const uint32_t* d_my_cnt, const uint32_t* d_my_flags,
                                     uint32_t capacity, uint32_t* d_status, unsigned long long timeout_ns,
                                     cudaStream_t stream, const uint32_t* d_skip_if_nonzero = nullptr,
                                     uint32_t* d_done = nullptr, uint32_t* d_seen = nullptr);
// d_seen (nullable, [kMaxPeers]): after a timeout, the last value read from every peer's flag cell.
// d_skip_if_nonzero: the kernel does nothing when that word is non-zero (speculative return wave behind a gather
// that may have recorded misses); d_done is set to 1 when the wave ran.
// Step 3: probe the cache for every key in the local inbox and store the rows (or the default vector) into
// row pos of the SENDER's output buffer; misses are appended to the miss list with kShardPosBits-encoded
// destinations.  Does nothing when *d_status != 0.
cudaError_t launch_probe_gather_inbox(const DeviceTable& t, const ShardPeers& peers, uint32_t world, uint32_t rank,
                                      uint32_t slot_cap,
                                      const int64_t* d_inbox_keys, const uint32_t* d_inbox_pos,
                                      const uint32_t* d_inbox_cnt, const uint32_t* d_status, uint32_t epoch, bool touch,
                                      uint32_t* d_miss_count, uint32_t* d_miss_pos, int64_t* d_miss_keys,
                                      int64_t* hd_miss_keys, size_t expected_keys, cudaStream_t stream);
// Rows of resolved misses (d_stage[i]) to their requesters' outputs.
cudaError_t launch_shard_scatter_stage(const float* d_stage, const uint32_t* d_miss_pos, size_t m, size_t dim,
                                       const ShardPeers& peers, uint32_t world, cudaStream_t stream);

// CUDA loads a kernel lazily, at its first launch, and that load can stall until the kernels already running on the
// device have finished.  A flag-wait kernel of the model-parallel exchange spins until ANOTHER stream's kernels have
// published; if one of those is launched for the first time at that moment (two ranks of one process on one device:
// tests/test_shard_group_gpu.py), the load waits for the spinner and the spinner for the load — until the timeout.
// (Measured with tools/stream_alias_probe.cu: of 552 stream pairs only the very first, the one that loads the
// setter kernel, is serialised; more hardware queues change nothing.)  These load every kernel a lookup of a
// shard group can launch, on the current device; hpsx_shard_group_create calls them.
cudaError_t preload_miss_path_kernels();  // kernels.cu
cudaError_t preload_shard_kernels();      // shard_kernels.cu

// Measurement primitive: out[i] = table[idx[i]] for 128-float rows (the random-gather ceiling the
// probe+gather kernel is compared with in bench.py).
cudaError_t launch_gather_rows(const float* d_table, const uint32_t* d_idx, size_t n, size_t dim,
                               float* d_out, cudaStream_t stream);

// Synthetic rows generated on the device (model-parallel shards too large for host memory).
cudaError_t launch_synth_rows(const int64_t* d_keys, size_t n, size_t dim, uint64_t seed,
                              float* d_rows, cudaStream_t stream);

}  // namespace hpsx
