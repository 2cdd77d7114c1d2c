#include "host_ps.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "hpsx_common.h"

namespace hpsx {

// ------------------------------------------------------------------------------------------------
// ThreadPool
// ------------------------------------------------------------------------------------------------
size_t ThreadPool::default_concurrency() {
  if (const char* env = std::getenv("HCTR_DEFAULT_CONCURRENCY")) {
    const long v = std::atol(env);
    if (v > 0) return static_cast<size_t>(v);
  }
  const unsigned hc = std::thread::hardware_concurrency();
  return hc > 0 ? hc : 4;
}

ThreadPool::ThreadPool(size_t num_threads) {
  if (num_threads == 0) num_threads = default_concurrency();
  // the caller of parallel_for is one of the lanes, so spawn one fewer
  for (size_t i = 1; i < num_threads; ++i) workers_.emplace_back([this] { worker_loop(); });
  // A pool of one lane has no worker; post() must still not run its job on the poster's thread (the poster may
  // hold locks the job needs: an asynchronous cache insertion posted from inside a lookup), so such a pool keeps
  // one thread that serves jobs only.
  if (workers_.empty()) job_thread_ = std::thread([this] { job_loop(); });
}

ThreadPool::~ThreadPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& t : workers_) t.join();
  if (job_thread_.joinable()) job_thread_.join();
}

void ThreadPool::job_loop() {
  while (true) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
      if (jobs_.empty()) return;  // stop_ and drained
      job = std::move(jobs_.front());
      jobs_.pop_front();
    }
    job();
  }
}

void ThreadPool::run_batch(const std::shared_ptr<Batch>& b) {
  while (true) {
    const size_t t = b->next.fetch_add(1, std::memory_order_relaxed);
    if (t >= b->num_tasks) break;
    (*b->fn)(t);
    if (b->done.fetch_add(1, std::memory_order_acq_rel) + 1 == b->num_tasks) {
      std::lock_guard<std::mutex> lk(b->mu);
      b->cv.notify_all();
    }
  }
}

void ThreadPool::worker_loop() {
  while (true) {
    std::shared_ptr<Batch> batch;
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return stop_ || !batches_.empty() || !jobs_.empty(); });
      if (stop_ && batches_.empty() && jobs_.empty()) return;
      // drop exhausted batches from the front
      while (!batches_.empty() &&
             batches_.front()->next.load(std::memory_order_relaxed) >= batches_.front()->num_tasks)
        batches_.pop_front();
      if (!batches_.empty()) {
        batch = batches_.front();
      } else if (!jobs_.empty()) {
        job = std::move(jobs_.front());
        jobs_.pop_front();
      } else {
        continue;
      }
    }
    if (batch)
      run_batch(batch);
    else if (job)
      job();
  }
}

void ThreadPool::parallel_for(size_t num_tasks, const std::function<void(size_t)>& fn) {
  if (num_tasks == 0) return;
  if (num_tasks == 1 || workers_.empty()) {
    for (size_t t = 0; t < num_tasks; ++t) fn(t);
    return;
  }
  auto b = std::make_shared<Batch>();
  b->fn = &fn;
  b->num_tasks = num_tasks;
  {
    std::lock_guard<std::mutex> lk(mu_);
    batches_.push_back(b);
  }
  cv_.notify_all();
  run_batch(b);
  std::unique_lock<std::mutex> lk(b->mu);
  b->cv.wait(lk, [&] { return b->done.load(std::memory_order_acquire) == b->num_tasks; });
}

void ThreadPool::post(std::function<void()> job) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    jobs_.push_back(std::move(job));
  }
  cv_.notify_all();  // workers and the job thread share the condition variable
}

// ------------------------------------------------------------------------------------------------
// HostTable
// ------------------------------------------------------------------------------------------------
HostTable::HostTable(size_t dim, float default_value, size_t num_partitions, size_t allocation_rate, size_t pull_window_bytes)
    : dim_(dim), default_value_(default_value) {
  if (dim == 0) throw std::invalid_argument("embedding vector size must be > 0");
  if (num_partitions == 0) num_partitions = 1;
  requested_partitions_ = num_partitions;
  if (allocation_rate == 0) allocation_rate = 256ull << 20;
  allocation_rate_ = allocation_rate;
  if (pull_window_bytes != 0) pull_window_bytes_ = std::max<size_t>(pull_window_bytes, 1u << 20);
  const size_t row_bytes = dim * sizeof(float);
  size_t rows_per_slab = std::max<size_t>(1, allocation_rate / row_bytes);
  slab_shift_ = 0;
  while ((2ull << slab_shift_) <= rows_per_slab) ++slab_shift_;
  slab_mask_ = (1ull << slab_shift_) - 1;
  for (size_t p = 0; p < num_partitions; ++p) {
    parts_.emplace_back(new Partition());
    parts_.back()->slots.assign(1024, Slot{kEmpty, 0});
  }
}

HostTable::~HostTable() {
  for (auto& p : parts_) {
    for (size_t i = 0; i < p->slabs.size(); ++i) {
      if (i < p->locked_bytes.size() && p->locked_bytes[i] != 0) cudaHostUnregister(p->slabs[i]);
      std::free(p->slabs[i]);
    }
  }
}

bool HostTable::pagelock(std::string* err) {
  std::unique_lock<std::shared_mutex> lk(rw_);
  const size_t row_bytes = dim_ * sizeof(float);
  const size_t rows_per_slab = static_cast<size_t>(1) << slab_shift_;
  for (auto& pp : parts_) {
    Partition& p = *pp;
    p.locked_bytes.resize(p.slabs.size(), 0);
    p.slab_device.resize(p.slabs.size(), nullptr);
    for (size_t i = 0; i < p.slabs.size(); ++i) {
      const size_t used_rows = std::min(rows_per_slab, p.rows_used - std::min(p.rows_used, i * rows_per_slab));
      // slabs are allocated in whole 4-KiB pages (alloc_row), so the rounded length never leaves the allocation
      const size_t want = std::min(slab_bytes(), (used_rows * row_bytes + 4095) & ~static_cast<size_t>(4095));
      if (want == 0 || want <= p.locked_bytes[i]) continue;
      if (p.locked_bytes[i] != 0) {
        cudaHostUnregister(p.slabs[i]);
        p.locked_bytes[i] = 0;
      }
      cudaError_t e = cudaHostRegister(p.slabs[i], want, cudaHostRegisterMapped | cudaHostRegisterPortable);
      void* dev = nullptr;
      if (e == cudaSuccess) e = cudaHostGetDevicePointer(&dev, p.slabs[i], 0);
      if (e != cudaSuccess) {
        cudaGetLastError();
        if (err) *err = std::string("page-locking a host table slab failed: ") + cudaGetErrorString(e);
        return false;
      }
      p.locked_bytes[i] = want;
      p.slab_device[i] = static_cast<char*>(dev);
    }
  }
  pagelocked_ = true;
  return true;
}

void HostTable::export_rows(size_t chunk,
                            const std::function<void(const int64_t*, const uint64_t*, size_t)>& fn) const {
  std::shared_lock<std::shared_mutex> lk(rw_);
  std::vector<int64_t> keys;
  std::vector<uint64_t> addrs;
  keys.reserve(chunk);
  addrs.reserve(chunk);
  for (const auto& pp : parts_) {
    const Partition& p = *pp;
    for (const Slot& s : p.slots) {
      if (s.key == kEmpty) continue;
      keys.push_back(s.key);
      addrs.push_back(row_device_addr(p, s.row));
      if (keys.size() == chunk) {
        fn(keys.data(), addrs.data(), keys.size());
        keys.clear();
        addrs.clear();
      }
    }
  }
  if (!keys.empty()) fn(keys.data(), addrs.data(), keys.size());
}

const float* HostTable::sentinel_row_device() const {
  std::shared_lock<std::shared_mutex> lk(rw_);
  for (const auto& pp : parts_)
    if (pp->has_sentinel && !pp->slab_device.empty())
      return reinterpret_cast<const float*>(row_device_addr(*pp, pp->sentinel_row));
  return nullptr;
}

uint64_t HostTable::alloc_row(Partition& p) {
  const uint64_t row = p.rows_used++;
  if ((row >> slab_shift_) >= p.slabs.size()) {
    void* mem = nullptr;
    if (posix_memalign(&mem, 4096, slab_bytes()) != 0 || mem == nullptr) throw std::bad_alloc();
    p.slabs.push_back(static_cast<float*>(mem));
  }
  return row;
}

void HostTable::grow(Partition& p) {
  std::vector<Slot> old;
  old.swap(p.slots);
  p.slots.assign(old.size() * 2, Slot{kEmpty, 0});
  const size_t mask = p.slots.size() - 1;
  for (const Slot& s : old) {
    if (s.key == kEmpty) continue;
    size_t i = mix64(static_cast<uint64_t>(s.key)) & mask;
    while (p.slots[i].key != kEmpty) i = (i + 1) & mask;
    p.slots[i] = s;
  }
}

float* HostTable::upsert(Partition& p, int64_t key, uint64_t h) {
  if (key == kEmpty) {
    if (!p.has_sentinel) {
      p.sentinel_row = alloc_row(p);
      p.has_sentinel = true;
      rows_.fetch_add(1, std::memory_order_relaxed);
    }
    return row_ptr(p, p.sentinel_row);
  }
  if ((p.count + 1) * 2 > p.slots.size()) grow(p);
  const size_t mask = p.slots.size() - 1;
  size_t i = h & mask;
  while (true) {
    Slot& s = p.slots[i];
    if (s.key == key) return row_ptr(p, s.row);
    if (s.key == kEmpty) {
      s.key = key;
      s.row = alloc_row(p);
      ++p.count;
      rows_.fetch_add(1, std::memory_order_relaxed);
      return row_ptr(p, s.row);
    }
    i = (i + 1) & mask;
  }
}

const float* HostTable::find(const Partition& p, int64_t key, uint64_t h) const {
  if (key == kEmpty) return p.has_sentinel ? row_ptr(p, p.sentinel_row) : nullptr;
  const size_t mask = p.slots.size() - 1;
  size_t i = h & mask;
  while (true) {
    const Slot& s = p.slots[i];
    if (s.key == key) return row_ptr(p, s.row);
    if (s.key == kEmpty) return nullptr;
    i = (i + 1) & mask;
  }
}

void HostTable::note_loaded(const int64_t* keys, size_t n) {
  load_order_.insert(load_order_.end(), keys, keys + n);
}

void HostTable::warm_keys(size_t count, std::vector<int64_t>& out) const {
  std::shared_lock<std::shared_mutex> lk(rw_);
  out.clear();
  out.reserve(count);
  for (size_t k = 0; k < procedural_rows_ && out.size() < count; ++k)
    out.push_back(static_cast<int64_t>(k));
  for (size_t i = 0; i < load_order_.size() && out.size() < count; ++i) out.push_back(load_order_[i]);
}

void HostTable::repartition_for(size_t expected_rows) {
  // caller holds rw_ exclusively
  if (partitions_sized_ || rows_.load(std::memory_order_relaxed) != 0) return;
  partitions_sized_ = true;  // the first bulk load (or reserve) decides
  const size_t row_bytes = dim_ * sizeof(float);
  const size_t bytes = expected_rows * row_bytes;
  // partitions of ~80 % of a window, so that the usual +-10 % spread of a hash partition still fits ONE slab
  const size_t fill = pull_window_bytes_ - pull_window_bytes_ / 5;
  size_t want = std::max(requested_partitions_, (bytes + fill - 1) / fill);
  want = std::min<size_t>(std::max<size_t>(want, 1), kMaxPartitions);
  // slabs of one window (a power of two of rows), never larger than allocation_rate: a partition allocates a second
  // slab only when it outgrows the first
  {
    const size_t per_part = (bytes + want - 1) / want;
    const size_t target = std::min(allocation_rate_, std::max<size_t>(pull_window_bytes_, per_part + per_part / 4));
    const size_t rows_per_slab = std::max<size_t>(1, target / row_bytes);
    slab_shift_ = 0;
    while ((1ull << slab_shift_) < rows_per_slab) ++slab_shift_;
    while (slab_shift_ > 0 && (1ull << slab_shift_) * row_bytes > allocation_rate_) --slab_shift_;
    slab_mask_ = (1ull << slab_shift_) - 1;
  }
  if (want == parts_.size()) return;
  parts_.clear();
  for (size_t p = 0; p < want; ++p) {
    parts_.emplace_back(new Partition());
    parts_.back()->slots.assign(1024, Slot{kEmpty, 0});
  }
}

void HostTable::reserve(size_t rows) {
  std::unique_lock<std::shared_mutex> lk(rw_);
  repartition_for(rows);
  const size_t per_part = rows / parts_.size() + rows / (parts_.size() * 8) + 64;
  for (auto& pp : parts_) {
    Partition& p = *pp;
    while ((p.count + per_part) * 2 > p.slots.size()) grow(p);
  }
}

template <typename KeyAt, typename Keep>
void HostTable::group_by_partition(size_t n, const KeyAt& key_at, const Keep& keep, ThreadPool& pool,
                                   std::vector<uint32_t>& order, std::vector<size_t>& offsets) const {
  const size_t P = parts_.size();
  const size_t tasks = std::max<size_t>(1, std::min(pool.size() * 2, (n + 65535) / 65536));
  const size_t per = (n + tasks - 1) / tasks;
  std::vector<size_t> counts(tasks * P, 0);
  pool.parallel_for(tasks, [&](size_t t) {
    size_t* c = counts.data() + t * P;
    const size_t e = std::min(n, (t + 1) * per);
    for (size_t i = t * per; i < e; ++i)
      if (keep(i)) ++c[partition_of(mix64(static_cast<uint64_t>(key_at(i))))];
  });
  offsets.assign(P + 1, 0);
  for (size_t p = 0; p < P; ++p) {
    size_t run = offsets[p];
    for (size_t t = 0; t < tasks; ++t) {
      const size_t c = counts[t * P + p];
      counts[t * P + p] = run;  // becomes the task's write cursor for this partition
      run += c;
    }
    offsets[p + 1] = run;
  }
  order.resize(offsets[P]);
  pool.parallel_for(tasks, [&](size_t t) {
    size_t* c = counts.data() + t * P;
    const size_t e = std::min(n, (t + 1) * per);
    for (size_t i = t * per; i < e; ++i)
      if (keep(i)) order[c[partition_of(mix64(static_cast<uint64_t>(key_at(i))))]++] = static_cast<uint32_t>(i);
  });
}

// Bulk loads run in chunks of kLoadChunk keys: the chunk is grouped by partition in parallel, then one task per
// partition (the maps are not concurrent) inserts its group.  Cost O(n) whatever the partition count.
static constexpr size_t kLoadChunk = static_cast<size_t>(1) << 24;

void HostTable::insert(const int64_t* keys, const float* vectors, size_t n, ThreadPool& pool) {
  std::unique_lock<std::shared_mutex> lk(rw_);
  repartition_for(n);
  note_loaded(keys, n);
  std::vector<uint32_t> order;
  std::vector<size_t> offsets;
  for (size_t base = 0; base < n; base += kLoadChunk) {
    const size_t m = std::min(kLoadChunk, n - base);
    const int64_t* k = keys + base;
    group_by_partition(m, [&](size_t i) { return k[i]; }, [](size_t) { return true; }, pool, order, offsets);
    pool.parallel_for(parts_.size(), [&](size_t p) {
      Partition& part = *parts_[p];
      for (size_t j = offsets[p]; j < offsets[p + 1]; ++j) {
        const size_t i = order[j];
        float* dst = upsert(part, k[i], mix64(static_cast<uint64_t>(k[i])));
        std::memcpy(dst, vectors + (base + i) * dim_, dim_ * sizeof(float));
      }
    });
  }
}

void HostTable::fill_procedural(size_t n, uint64_t seed, ThreadPool& pool, uint32_t shard,
                                uint32_t num_shards) {
  std::unique_lock<std::shared_mutex> lk(rw_);
  const bool sharded = num_shards > 1;
  const size_t mine = sharded ? n / num_shards + n / (8 * num_shards) : n;
  repartition_for(mine);
  if (!sharded) procedural_rows_ = std::max(procedural_rows_, n);
  const size_t P = parts_.size();
  {
    const size_t per_part = mine / P + mine / (P * 8) + 64;
    for (auto& pp : parts_)
      while ((pp->count + per_part) * 2 > pp->slots.size()) grow(*pp);
  }
  std::vector<uint32_t> order;
  std::vector<size_t> offsets;
  for (size_t base = 0; base < n; base += kLoadChunk) {
    const size_t m = std::min(kLoadChunk, n - base);
    auto key_at = [&](size_t i) { return static_cast<int64_t>(base + i); };
    auto keep = [&](size_t i) { return !sharded || owner_of(key_at(i), num_shards) == shard; };
    group_by_partition(m, key_at, keep, pool, order, offsets);
    // one task per partition claims the rows of its keys and generates them in place
    pool.parallel_for(P, [&](size_t p) {
      Partition& part = *parts_[p];
      for (size_t j = offsets[p]; j < offsets[p + 1]; ++j) {
        const int64_t key = key_at(order[j]);
        float* dst = upsert(part, key, mix64(static_cast<uint64_t>(key)));
        for (size_t d = 0; d < dim_; ++d) dst[d] = synth_value(key, static_cast<uint32_t>(d), seed);
      }
    });
    if (sharded) {
      // model-parallel shard: only the keys this shard owns exist here; they are also the warm-up order
      std::vector<uint32_t> kept(order);
      std::sort(kept.begin(), kept.end());
      for (uint32_t i : kept) load_order_.push_back(key_at(i));
    }
  }
}

size_t HostTable::fetch_range(const int64_t* keys, size_t begin, size_t end, float* out,
                              size_t stride) const {
  // Rolling three-stage software pipeline over the keys: key i is hashed and its home slot prefetched, key i - kLag is
  // resolved (the slot line has arrived) and its row prefetched, key i - 2 kLag is copied (the row has arrived).  Unlike a
  // batch-at-a-time pipeline there is no drain between batches: ~2 kLag cache misses stay in flight throughout.
  // kLag = 8 measured best on the build container's cores (1 M x 32 table, requests of 1024 keys: ~50 us against 57-76 us
  // for the batch-of-16 form and 54-90 us for kLag = 16: a core has about a dozen line-fill buffers, a key needs three lines).
  constexpr size_t kLag = 8, kRing = 32;
  static_assert(kRing >= 2 * kLag + 1 && (kRing & (kRing - 1)) == 0, "ring must hold the whole pipeline");
  const size_t row_bytes = dim_ * sizeof(float);
  size_t absent = 0;
  uint64_t hs[kRing];
  const float* rows[kRing];
  const size_t n = end - begin;
  for (size_t i = 0; i < n + 2 * kLag; ++i) {
    if (i < n) {
      const uint64_t h = mix64(static_cast<uint64_t>(keys[begin + i]));
      hs[i & (kRing - 1)] = h;
      const Partition& p = *parts_[partition_of(h)];
      __builtin_prefetch(&p.slots[h & (p.slots.size() - 1)], 0, 0);
    }
    if (i >= kLag && i - kLag < n) {
      const size_t k = i - kLag;
      const uint64_t h = hs[k & (kRing - 1)];
      const float* r = find(*parts_[partition_of(h)], keys[begin + k], h);
      rows[k & (kRing - 1)] = r;
      if (r) {
        const char* c = reinterpret_cast<const char*>(r);
        for (size_t off = 0; off < row_bytes; off += 64) __builtin_prefetch(c + off, 0, 0);
      }
    }
    if (i >= 2 * kLag) {
      const size_t k = i - 2 * kLag;
      float* dst = out + (begin + k) * stride;
      const float* r = rows[k & (kRing - 1)];
      if (r) {
        std::memcpy(dst, r, row_bytes);
      } else {
        for (size_t d = 0; d < dim_; ++d) dst[d] = default_value_;
        ++absent;
      }
    }
  }
  return absent;
}

size_t HostTable::fetch(const int64_t* keys, size_t n, float* out, size_t stride,
                        ThreadPool& pool) const {
  std::shared_lock<std::shared_mutex> lk(rw_);
  if (n == 0) return 0;
  constexpr size_t kMinChunk = 2048;
  const size_t max_tasks = pool.size() * 4;
  const size_t tasks = std::max<size_t>(1, std::min(max_tasks, (n + kMinChunk - 1) / kMinChunk));
  if (tasks == 1) return fetch_range(keys, 0, n, out, stride);
  const size_t chunk = (n + tasks - 1) / tasks;
  std::atomic<size_t> absent{0};
  pool.parallel_for(tasks, [&](size_t t) {
    const size_t b = t * chunk;
    const size_t e = std::min(n, b + chunk);
    if (b < e) absent.fetch_add(fetch_range(keys, b, e, out, stride), std::memory_order_relaxed);
  });
  return absent.load();
}

}  // namespace hpsx
