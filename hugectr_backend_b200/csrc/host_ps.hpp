// Host-DRAM parameter server: the volatile database of the HPS (hash_map / parallel_hash_map
// backend) behind the HBM cache.  Serves cache misses and the gpucache=false CPU path.
//
// Behaviour restated from the reference documentation (docs/hierarchical_parameter_server.md:67-78,
// 244-246, 400-416): tables are hash-partitioned into `num_partitions` flat hash maps, values live in
// `allocation_rate`-sized slabs, a key that is in no database gets default_value_for_each_table.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "hpsx_common.h"

namespace hpsx {

// Fixed pool; size follows the reference (HCTR_DEFAULT_CONCURRENCY, else hardware_concurrency:
// src/thread_pool.cpp:25-41).  parallel_for is re-entrant from several caller threads.
class ThreadPool {
 public:
  explicit ThreadPool(size_t num_threads);
  ~ThreadPool();
  size_t size() const { return workers_.size() + 1; }  // workers + the calling thread
  // Runs fn(task) for task in [0, num_tasks); the caller participates; returns when all are done.
  void parallel_for(size_t num_tasks, const std::function<void(size_t)>& fn);
  // Fire-and-forget job (asynchronous cache insertion).
  void post(std::function<void()> job);
  static size_t default_concurrency();

 private:
  struct Batch {
    const std::function<void(size_t)>* fn;
    size_t num_tasks;
    std::atomic<size_t> next{0};
    std::atomic<size_t> done{0};
    std::mutex mu;
    std::condition_variable cv;
  };
  void worker_loop();
  void job_loop();
  static void run_batch(const std::shared_ptr<Batch>& b);

  std::vector<std::thread> workers_;
  std::thread job_thread_;  // only when workers_ is empty: serves post()ed jobs
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Batch>> batches_;
  std::deque<std::function<void()>> jobs_;
  bool stop_ = false;
};

// One embedding table in host DRAM.
class HostTable {
 public:
  HostTable(size_t dim, float default_value, size_t num_partitions, size_t allocation_rate, size_t pull_window_bytes = 0);
  ~HostTable();
  HostTable(const HostTable&) = delete;
  HostTable& operator=(const HostTable&) = delete;

  size_t dim() const { return dim_; }
  float default_value() const { return default_value_; }
  size_t rows() const { return rows_.load(std::memory_order_relaxed); }
  size_t num_partitions() const { return parts_.size(); }

  // Insert or overwrite `n` rows (key file + vector file contents).
  void insert(const int64_t* keys, const float* vectors, size_t n, ThreadPool& pool);
  // keys [0,n) with synthetic rows, see hpsx_common.h synth_value().
  // With num_shards > 1 only the keys whose owner_of(key, num_shards) == shard are loaded
  // (model-parallel row sharding).
  void fill_procedural(size_t n, uint64_t seed, ThreadPool& pool, uint32_t shard = 0, uint32_t num_shards = 1);

  // out[i*stride .. +dim) = row(keys[i]) or default.  Returns the number of absent keys.
  // Multi-threaded over key ranges; software-prefetched probe + row gather.
  size_t fetch(const int64_t* keys, size_t n, float* out, size_t stride, ThreadPool& pool) const;

  // The first min(count, rows) keys in load order (cache warm-up, a9).
  void warm_keys(size_t count, std::vector<int64_t>& out) const;
  // Pre-size the partition maps for `rows` more rows (avoids rehashing during bulk loads).  On a still EMPTY table
  // this also fixes the partition count: at least the configured `num_partitions`, and enough of them that one
  // partition's rows fill at most kPullWindowBytes of host memory (see below).
  void reserve(size_t rows);

  // Direct pull reads rows over PCIe faster the closer together the rows in flight lie in host memory
  // (tools/pcie_probe2.cu; profiles/pcie_probe2_r02.txt, pcie_probe2_48g_r02.txt): a 5 GiB table gives 39 GB/s in
  // random order and 51 GB/s inside windows of <= 256 MiB; a 48 GiB table gives 25 GB/s random, 41 GB/s inside
  // 64-256 MiB windows, 45.5 GB/s inside 8-16 MiB windows and 46 GB/s fully sorted.  A partition therefore doubles
  // as a locality bin: its rows live in slabs of its own (one window), partition_of(key) needs no memory access, and
  // the probe kernel appends a missed key straight to its partition's miss list.  Default window: 16 MiB
  // (volatile_db "hpsx_pull_window_mb" / hpsx_volatile_params.pull_window_bytes); at most kMaxPartitions partitions.
  static constexpr size_t kDefaultPullWindowBytes = 16ull << 20;
  static constexpr size_t kMaxPartitions = 4096;

  // enable_pagelock (reference key: src/backend.cpp:506-511): page-lock the used part of every value
  // slab and map it into the CUDA address space, so kernels can read rows straight from host DRAM
  // ("direct pull").  Idempotent; call again after the table grew.  False + *err on failure.
  bool pagelock(std::string* err);
  bool pagelocked() const { return pagelocked_; }
  // Visits every (key, device-visible row address) pair of a page-locked table in chunks of at most
  // `chunk` rows (the key that doubles as the empty marker is reported by sentinel_row_device()).
  void export_rows(size_t chunk,
                   const std::function<void(const int64_t*, const uint64_t*, size_t)>& fn) const;
  const float* sentinel_row_device() const;

 private:
  struct Slot {
    int64_t key;
    uint64_t row;  // index into this partition's slabs
  };
  struct Partition {
    std::vector<Slot> slots;  // open addressing, power-of-two capacity
    size_t count = 0;
    std::vector<float*> slabs;
    std::vector<size_t> locked_bytes;   // page-locked prefix of every slab (0: not registered)
    std::vector<char*> slab_device;     // device-visible base address of every registered slab
    size_t rows_used = 0;
    bool has_sentinel = false;  // row of the key that doubles as the empty marker
    uint64_t sentinel_row = 0;
    std::mutex mu;
  };

  static constexpr int64_t kEmpty = INT64_MIN;
  size_t partition_of(uint64_t h) const { return host_partition_of_hash(h, static_cast<uint32_t>(parts_.size())); }
  // [begin, end) of `keys`, single-threaded; the caller holds rw_ (shared)
  size_t fetch_range(const int64_t* keys, size_t begin, size_t end, float* out, size_t stride) const;
  // Only while the table is empty: partition count for a table of `expected_rows` rows.
  void repartition_for(size_t expected_rows);
  // order[offsets[p] .. offsets[p+1]) = the indices i in [0, n) with keep(i) whose key_at(i) lives in partition p
  template <typename KeyAt, typename Keep>
  void group_by_partition(size_t n, const KeyAt& key_at, const Keep& keep, ThreadPool& pool, std::vector<uint32_t>& order,
                          std::vector<size_t>& offsets) const;
  float* row_ptr(const Partition& p, uint64_t row) const {
    return p.slabs[row >> slab_shift_] + (row & slab_mask_) * dim_;
  }
  uint64_t row_device_addr(const Partition& p, uint64_t row) const {
    return reinterpret_cast<uint64_t>(p.slab_device[row >> slab_shift_]) +
           (row & slab_mask_) * dim_ * sizeof(float);
  }
  // returns the row pointer of `key` in partition `p`, inserting a fresh row if absent
  float* upsert(Partition& p, int64_t key, uint64_t h);
  const float* find(const Partition& p, int64_t key, uint64_t h) const;
  void grow(Partition& p);
  uint64_t alloc_row(Partition& p);
  size_t slab_bytes() const {  // whole 4-KiB pages, so that a page-locked prefix never leaves the allocation
    return ((static_cast<size_t>(1) << slab_shift_) * dim_ * sizeof(float) + 4095) & ~static_cast<size_t>(4095);
  }
  void note_loaded(const int64_t* keys, size_t n);

  size_t dim_;
  float default_value_;
  size_t requested_partitions_ = 1;
  size_t allocation_rate_ = 256ull << 20;
  size_t pull_window_bytes_ = kDefaultPullWindowBytes;
  bool partitions_sized_ = false;
  std::vector<std::unique_ptr<Partition>> parts_;
  size_t slab_shift_;  // rows per slab = 1 << slab_shift_
  uint64_t slab_mask_;
  std::atomic<size_t> rows_{0};
  mutable std::shared_mutex rw_;  // fetch: shared; insert/fill: exclusive
  std::vector<int64_t> load_order_;   // keys in the order insert() first saw them
  size_t procedural_rows_ = 0;        // fill_procedural(n): keys [0,n) precede load_order_
  bool pagelocked_ = false;
};

}  // namespace hpsx
