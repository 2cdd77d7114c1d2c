// libhpsx.so — C ABI of the HPS engine (include/hpsx.h): host orchestration of the lookup hot path.
//
//   keys (host) --H2D--> probe+gather kernel --> out (device, possibly the Triton output buffer)
//                              | miss list (warp-compacted)
//                              v
//        D2H miss keys -> host parameter server gather (pinned, chunked, double-buffered)
//                      -> H2D rows -> merge+insert kernel (writes `out` and the cache slot)
//
// The reference performs the same stages inside libhuge_ctr_hps.so behind
// LookupSessionBase::lookup (call site hps_backend/src/model_instance_state.cpp:194-195).
#include <cuda.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "engine.hpp"
#include "engine_internal.hpp"

using namespace hpsx;

namespace hpsx {
namespace eng {

thread_local std::string g_err;

int fail(int code, std::string msg) {
  g_err = std::move(msg);
  return code;
}

bool primary_context_active(int dev) {
  typedef CUresult (*Fn)(CUdevice, unsigned int*, int*);
  static Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<Fn>(p);
  }();
  if (fn == nullptr) return true;
  unsigned int flags = 0;
  int active = 0;
  return fn(static_cast<CUdevice>(dev), &flags, &active) != CUDA_SUCCESS || active != 0;
}

}  // namespace eng
}  // namespace hpsx

namespace hpsx {
namespace eng {

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// HPSX_TRACE=1: one stderr line per lookup with the host-side timeline (debugging aid only).
bool trace_on() {
  static const bool on = [] {
    const char* e = std::getenv("HPSX_TRACE");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

// ------------------------------------------------------------------------------------------------
// sparse model files: <dir>/key (int64 | uint32) + <dir>/emb_vector (fp32 row-major)
// (docs/architecture.md:185-218; writer samples/hps-triton-ensemble/01_model_training.ipynb:498-505)
// ------------------------------------------------------------------------------------------------

}  // namespace eng
}  // namespace hpsx

using namespace hpsx::eng;

namespace {

Model* find_model(hpsx_ps* ps, const char* name) {
  if (!ps || !name) return nullptr;
  std::lock_guard<std::mutex> lk(ps->mu);
  auto it = ps->models.find(name);
  return it == ps->models.end() ? nullptr : it->second.get();
}

size_t resolve_partitions(size_t requested) {
  if (requested > 0) return requested;
  // docs/hierarchical_parameter_server.md:410-412: min(number of cores, 16)
  return std::max<size_t>(1, std::min<size_t>(ThreadPool::default_concurrency(), 16));
}

int load_sparse_dir(hpsx_ps* ps, HostTable* table, const std::string& dir) {
  // engine extension for benchmarks: "synthetic:rows=<N>,seed=<S>" names a procedural table
  // (keys [0,N), rows from synth_value()) instead of a directory, so that a 10 M-row table needs no 5 GB file
  if (dir.rfind("synthetic:", 0) == 0) {
    unsigned long long rows = 0, seed = 0;
    if (std::sscanf(dir.c_str(), "synthetic:rows=%llu,seed=%llu", &rows, &seed) != 2)
      return fail(HPSX_ERR_INVALID_ARG, "malformed synthetic table spec '" + dir + "' (want synthetic:rows=<N>,seed=<S>)");
    table->fill_procedural(static_cast<size_t>(rows), static_cast<uint64_t>(seed), *ps->pool);
    return HPSX_OK;
  }
  const std::string key_path = dir + "/key";
  const std::string vec_path = dir + "/emb_vector";
  struct stat ks {}, vs {};
  if (stat(key_path.c_str(), &ks) != 0)
    return fail(HPSX_ERR_IO, "cannot stat sparse key file '" + key_path + "'");
  if (stat(vec_path.c_str(), &vs) != 0)
    return fail(HPSX_ERR_IO, "cannot stat sparse vector file '" + vec_path + "'");
  const size_t row_bytes = table->dim() * sizeof(float);
  if (static_cast<size_t>(vs.st_size) % row_bytes != 0)
    return fail(HPSX_ERR_IO, "'" + vec_path + "' is not a whole number of " +
                                 std::to_string(table->dim()) + "-float rows");
  const size_t rows = static_cast<size_t>(vs.st_size) / row_bytes;
  if (rows == 0) {
    if (ks.st_size != 0) return fail(HPSX_ERR_IO, "'" + key_path + "' has keys but no vectors");
    return HPSX_OK;
  }
  size_t key_bytes = 0;
  if (static_cast<size_t>(ks.st_size) == rows * 8)
    key_bytes = 8;
  else if (static_cast<size_t>(ks.st_size) == rows * 4)
    key_bytes = 4;
  else
    return fail(HPSX_ERR_IO, "'" + key_path + "' (" + std::to_string(ks.st_size) +
                                 " B) does not match " + std::to_string(rows) + " rows of '" +
                                 vec_path + "'");
  FILE* kf = std::fopen(key_path.c_str(), "rb");
  FILE* vf = std::fopen(vec_path.c_str(), "rb");
  if (!kf || !vf) {
    if (kf) std::fclose(kf);
    if (vf) std::fclose(vf);
    return fail(HPSX_ERR_IO, "cannot open sparse files under '" + dir + "'");
  }
  table->reserve(rows);
  constexpr size_t kChunk = 1 << 18;
  std::vector<int64_t> keys(std::min(rows, kChunk));
  std::vector<uint32_t> keys32(key_bytes == 4 ? keys.size() : 0);
  std::vector<float> vecs(keys.size() * table->dim());
  int rc = HPSX_OK;
  for (size_t done = 0; done < rows && rc == HPSX_OK;) {
    const size_t n = std::min(kChunk, rows - done);
    size_t got;
    if (key_bytes == 8) {
      got = std::fread(keys.data(), 8, n, kf);
    } else {
      got = std::fread(keys32.data(), 4, n, kf);
      for (size_t i = 0; i < got; ++i) keys[i] = static_cast<int64_t>(keys32[i]);
    }
    if (got != n || std::fread(vecs.data(), row_bytes, n, vf) != n) {
      rc = fail(HPSX_ERR_IO, "short read in sparse files under '" + dir + "'");
      break;
    }
    table->insert(keys.data(), vecs.data(), n, *ps->pool);
    done += n;
  }
  std::fclose(kf);
  std::fclose(vf);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// cache construction + warm-up
// ------------------------------------------------------------------------------------------------
int insert_rows_from_host(hpsx_ps* ps, hpsx_cache* c, size_t t, const int64_t* keys, size_t n,
                          cudaStream_t stream, int64_t* d_keys, float* h_stage, float* d_stage,
                          size_t chunk_rows, uint32_t* d_inserted) {
  // caller holds c->rw exclusively and has the device selected
  const HostTable& ht = *c->model->tables[t];
  const size_t dim = ht.dim();
  for (size_t off = 0; off < n; off += chunk_rows) {
    const size_t m = std::min(chunk_rows, n - off);
    ht.fetch(keys + off, m, h_stage, dim, *ps->pool);
    HPSX_CU(cudaMemcpyAsync(d_keys, keys + off, m * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    HPSX_CU(cudaMemcpyAsync(d_stage, h_stage, m * dim * sizeof(float), cudaMemcpyHostToDevice,
                            stream));
    const uint32_t epoch = c->epoch.fetch_add(1, std::memory_order_relaxed);
    HPSX_CU(launch_insert_merge(c->tables[t], d_keys, nullptr, d_stage, m, nullptr, true, epoch,
                                d_inserted, stream));
    HPSX_CU(cudaStreamSynchronize(stream));  // h_stage / d_keys are reused by the next chunk
  }
  return HPSX_OK;
}

// enable_pagelock: page-lock the host tables and mirror their key -> row-address index in HBM, so that misses
// are pulled by kernels straight from host DRAM (no CPU gather, no staging copy).  Called at cache creation
// and again after the host tables changed (update_database): registers the new slab ranges, re-allocates an
// index that became too small and rebuilds it.  The caller excludes lookups (c->rw exclusive, or the cache is
// not published yet) and has the device selected.
int sync_direct_pull_index_impl(hpsx_cache* c, cudaStream_t stream) {
  Model* model = c->model;
  const size_t T = model->tables.size();
  int64_t* d_ikeys = nullptr;
  uint64_t* d_iaddrs = nullptr;
  constexpr size_t kIndexChunk = 1 << 20;
  HPSX_CU(cudaMalloc(&d_ikeys, kIndexChunk * sizeof(int64_t)));
  HPSX_CU(cudaMalloc(&d_iaddrs, kIndexChunk * sizeof(uint64_t)));
  int rc = HPSX_OK;
  if (c->indexes.size() < T) c->indexes.resize(T, nullptr);
  for (size_t t = 0; t < T && rc == HPSX_OK; ++t) {
    HostTable& ht = *model->tables[t];
    std::string err;
    if (!ht.pagelock(&err)) {
      rc = fail(HPSX_ERR_CUDA, err);
      break;
    }
    uint64_t cap = 1024;
    while (cap < 2 * ht.rows()) cap <<= 1;
    IndexSlot* slots = c->indexes[t];
    cudaError_t ce = cudaSuccess;
    if (slots == nullptr || c->tables[t].index_mask + 1 < cap) {
      if (slots != nullptr) cudaFree(slots);
      c->indexes[t] = slots = nullptr;
      c->tables[t].index = nullptr;
      ce = cudaMalloc(&slots, cap * sizeof(IndexSlot));
      if (ce == cudaSuccess) c->indexes[t] = slots;
    } else {
      cap = c->tables[t].index_mask + 1;
    }
    if (ce == cudaSuccess) ce = launch_index_clear(slots, cap, stream);
    if (ce == cudaSuccess) {
      ht.export_rows(kIndexChunk, [&](const int64_t* k, const uint64_t* a, size_t n) {
        if (ce != cudaSuccess) return;
        ce = cudaMemcpyAsync(d_ikeys, k, n * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_iaddrs, a, n * sizeof(uint64_t), cudaMemcpyHostToDevice, stream);
        if (ce == cudaSuccess) ce = launch_index_build(slots, cap - 1, d_ikeys, d_iaddrs, n, stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);  // the pageable vectors are reused
      });
    }
    if (ce != cudaSuccess) {
      rc = fail(HPSX_ERR_CUDA, std::string("building the direct-pull index: ") + cudaGetErrorString(ce));
      break;
    }
    c->tables[t].index = slots;
    c->tables[t].index_mask = cap - 1;
    c->tables[t].sentinel_row = ht.sentinel_row_device();
    c->tables[t].host_parts = static_cast<uint32_t>(ht.num_partitions());
  }
  cudaFree(d_ikeys);
  cudaFree(d_iaddrs);
  return rc;
}

int build_cache(hpsx_ps* ps, Model* model, int device, std::unique_ptr<hpsx_cache>* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(HPSX_ERR_CUDA, "model '" + model->cfg.model_name + "' is deployed on device " +
                                   std::to_string(device) + " but only " + std::to_string(ndev) +
                                   " CUDA device(s) are usable");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice(" + std::to_string(device) + ") failed");

  std::unique_ptr<hpsx_cache> c(new hpsx_cache());
  c->model = model;
  c->device = device;
  c->is_static = model->cfg.embedding_cache_type == CacheType::Static;
  const size_t T = model->tables.size();
  c->tables.resize(T);
  c->slots.resize(T);
  for (size_t t = 0; t < T; ++t) c->max_dim = std::max(c->max_dim, model->tables[t]->dim());

  cudaStream_t stream = nullptr;
  HPSX_CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  HPSX_CU(cudaStreamCreateWithFlags(&c->async_stream, cudaStreamNonBlocking));
  const size_t chunk = kStageChunkRows;
  int64_t* d_keys = nullptr;
  float *h_stage = nullptr, *d_stage = nullptr;
  uint32_t* d_inserted = nullptr;
  HPSX_CU(cudaMalloc(&d_keys, chunk * sizeof(int64_t)));
  HPSX_CU(cudaMalloc(&d_stage, chunk * c->max_dim * sizeof(float)));
  HPSX_CU(cudaMallocHost(&h_stage, chunk * c->max_dim * sizeof(float)));
  HPSX_CU(cudaMalloc(&d_inserted, sizeof(uint32_t)));
  // the asynchronous-insert workspace reuses these buffers after warm-up
  c->async_d_keys = d_keys;
  c->async_d_stage = d_stage;
  c->async_h_stage = h_stage;
  c->async_rows = chunk;

  int rc = HPSX_OK;
  std::vector<int64_t> warm;
  for (size_t t = 0; t < T && rc == HPSX_OK; ++t) {
    const HostTable& ht = *model->tables[t];
    const size_t rows = ht.rows();
    double pct = c->is_static ? 1.0 : static_cast<double>(model->cfg.cache_size_percentage);
    pct = std::min(1.0, std::max(0.0, pct));
    const size_t warm_rows = std::min(rows, static_cast<size_t>(std::ceil(pct * static_cast<double>(rows))));
    const double lf = model->load_factor > 0.f ? model->load_factor : 0.5;
    const size_t want_slots = static_cast<size_t>(std::ceil(static_cast<double>(warm_rows) / lf));
    const size_t buckets = std::max<size_t>(1, (want_slots + kWays - 1) / kWays);
    if (buckets > 0xFFFFFFFFull / kWays)
      return fail(HPSX_ERR_UNSUPPORTED, "embedding cache of table " + std::to_string(t) +
                                            " needs more than 2^32 slots");
    DeviceTable& dt = c->tables[t];
    dt.num_buckets = static_cast<uint32_t>(buckets);
    dt.dim = static_cast<uint32_t>(ht.dim());
    dt.default_value = ht.default_value();
    c->slots[t] = buckets * kWays;
    HPSX_CU(cudaMalloc(&dt.buckets, buckets * sizeof(Bucket)));
    HPSX_CU(cudaMalloc(&dt.values, buckets * kWays * ht.dim() * sizeof(float)));
    HPSX_CU(launch_table_clear(dt, stream));
    HPSX_CU(cudaStreamSynchronize(stream));
    if (warm_rows > 0) {
      ht.warm_keys(warm_rows, warm);
      std::unique_lock<std::shared_mutex> lk(c->rw);
      rc = insert_rows_from_host(ps, c.get(), t, warm.data(), warm.size(), stream, d_keys, h_stage,
                                 d_stage, chunk, nullptr);
    }
  }
  if (rc == HPSX_OK && model->direct_pull) {
    rc = sync_direct_pull_index_impl(c.get(), stream);
    c->direct_pull = rc == HPSX_OK;
  }
  cudaFree(d_inserted);
  cudaStreamDestroy(stream);
  if (rc != HPSX_OK) return rc;
  *out = std::move(c);
  return HPSX_OK;
}

// ------------------------------------------------------------------------------------------------
// session helpers
// ------------------------------------------------------------------------------------------------
int check_tables(hpsx_session* s, const size_t* n_per_table, size_t num_tables) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  const size_t T = s->model->tables.size();
  if (num_tables > T)
    return fail(HPSX_ERR_INVALID_ARG, "lookup names " + std::to_string(num_tables) +
                                          " tables but model '" + s->model->cfg.model_name +
                                          "' has " + std::to_string(T));
  if (num_tables > 0 && !n_per_table) return fail(HPSX_ERR_INVALID_ARG, "null num_keys_per_table");
  for (size_t t = 0; t < num_tables; ++t) {
    if (n_per_table[t] > s->cap_per_table[t])
      return fail(HPSX_ERR_INVALID_ARG,
                  "table " + std::to_string(t) + ": " + std::to_string(n_per_table[t]) +
                      " keys exceed max_batch_size * maxnum_catfeature_query_per_table_per_sample = " +
                      std::to_string(s->cap_per_table[t]));
  }
  return HPSX_OK;
}

struct AsyncJob {
  hpsx_ps* ps;
  hpsx_cache* cache;
  size_t table;
  std::vector<int64_t> keys;
};

void run_async_insert(AsyncJob job) {
  hpsx_cache* c = job.cache;
  {
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> ws(c->async_mu);  // one job owns the workspace
    std::unique_lock<std::shared_mutex> lk(c->rw);
    // dedup on the host: the same missing key often repeats inside one request
    std::sort(job.keys.begin(), job.keys.end());
    job.keys.erase(std::unique(job.keys.begin(), job.keys.end()), job.keys.end());
    uint32_t* d_ins = nullptr;
    if (cudaMalloc(&d_ins, sizeof(uint32_t)) == cudaSuccess) {
      cudaMemsetAsync(d_ins, 0, sizeof(uint32_t), c->async_stream);
      insert_rows_from_host(job.ps, c, job.table, job.keys.data(), job.keys.size(), c->async_stream,
                            c->async_d_keys, c->async_h_stage, c->async_d_stage, c->async_rows,
                            d_ins);
      uint32_t h = 0;
      cudaMemcpy(&h, d_ins, sizeof(h), cudaMemcpyDeviceToHost);
      c->async_inserted.fetch_add(h, std::memory_order_relaxed);
      cudaFree(d_ins);
    }
  }
  std::lock_guard<std::mutex> lk(c->async_mu);
  --c->async_pending;
  c->async_cv.notify_all();
}

// Stream the rows of the `m` missing keys of table `t` host -> device.  The keys are already on the
// host: the probe kernel mirrored them into the mapped pinned buffer s->h_miss_keys + key_off.
// Per chunk: thread-pool gather from the host database into a pinned stage, cudaMemcpyAsync H2D, then
// (when `d_out` or `insert`) the fused merge+insert kernel.  Three stages rotate, so the gather of
// chunk i+1 overlaps the copy and kernel of chunk i; nothing here waits for the stream except to
// reuse a stage.  With `d_all_stage` the rows land at d_all_stage + i*dim and stay there for the
// caller (pooled path).  The first kernel that rewrites cache slots takes `wlock` (exclusive) and
// keeps it; the caller synchronises the stream before releasing it.
// bf16 mirror destination of rows [row_off, ...) of virtual table t, or nullptr (no mirror requested / other table)
void* bf16_dst(const hpsx_session* s, size_t t, size_t row_off, size_t dim) {
  if (s->bf16_out == nullptr || t != s->bf16_table) return nullptr;
  return static_cast<unsigned char*>(s->bf16_out) + row_off * dim * 2u;
}

}  // namespace

namespace hpsx {
namespace eng {

int sync_direct_pull_index(hpsx_cache* c, cudaStream_t stream) { return sync_direct_pull_index_impl(c, stream); }

int stream_miss_rows(hpsx_session* s, size_t t, size_t key_off, uint32_t m, float* d_out, bool insert,
                     uint32_t epoch, float* d_all_stage, std::unique_lock<std::shared_mutex>* wlock,
                     const MissBufs* mb) {
  hpsx_cache* c = s->cache;
  const size_t rt = t % s->model->tables.size();  // t may be a virtual table (request * T + table)
  const HostTable& ht = *s->model->tables[rt];
  const size_t dim = ht.dim();
  const int64_t* h_keys = mb ? mb->h_keys : s->h_miss_keys + key_off;
  const int64_t* d_mkeys = mb ? mb->d_keys : s->d_miss_keys + key_off;
  const uint32_t* d_mpos = mb ? mb->d_pos : s->d_miss_pos + key_off;
  uint32_t* d_inserted = s->d_counters + s->vt + t;
  s->stats.misses += m;
  size_t chunk = (static_cast<size_t>(m) + 3) / 4;
  chunk = std::min<size_t>(kStageChunkRows, std::max<size_t>(kMinStageChunkRows, chunk));
  size_t ci = 0;
  for (size_t off = 0; off < m; off += chunk, ++ci) {
    const size_t mc = std::min<size_t>(chunk, m - off);
    const int b = static_cast<int>(ci % kNumStages);
    // the copy (+ kernel) that last used this stage must be done before the host refills it
    HPSX_CU(cudaEventSynchronize(s->stage_free[b]));
    const double t0 = now_ms();
    const size_t absent = ht.fetch(h_keys + off, mc, s->h_stage[b], dim, *s->ps->pool);
    s->stats.host_gather_ms += now_ms() - t0;
    s->stats.default_filled += absent;
    float* d_rows = d_all_stage ? d_all_stage + off * dim : s->d_stage[b];
    HPSX_CU(cudaMemcpyAsync(d_rows, s->h_stage[b], mc * dim * sizeof(float), cudaMemcpyHostToDevice,
                            s->stream));
    s->stats.h2d_bytes += mc * dim * sizeof(float);
    if (d_out != nullptr || insert) {
      if (insert && wlock != nullptr && !wlock->owns_lock()) wlock->lock();
      HPSX_CU(launch_insert_merge(c->tables[rt], d_mkeys + off, d_mpos + off, d_rows, mc, d_out, insert, epoch,
                                  d_inserted, s->stream, d_out ? bf16_dst(s, t, 0, dim) : nullptr));
      ++s->stats.kernel_launches;
    }
    HPSX_CU(cudaEventRecord(s->stage_free[b], s->stream));
  }
  return HPSX_OK;
}

}  // namespace eng
}  // namespace hpsx

namespace {

// Asynchronous insertion: this response already carries the default vector for the misses (written
// by the probe kernel); a pool worker fetches + inserts later (docs/architecture.md:65-67).
void post_async_insert(hpsx_session* s, size_t t, size_t key_off, uint32_t m) {
  hpsx_cache* c = s->cache;
  s->stats.misses += m;
  s->stats.default_filled += m;
  const int64_t* h_keys = s->h_miss_keys + key_off;
  AsyncJob job{s->ps, c, t % s->model->tables.size(), std::vector<int64_t>(h_keys, h_keys + m)};
  {
    std::lock_guard<std::mutex> lk(c->async_mu);
    ++c->async_pending;
  }
  s->ps->pool->post([job]() mutable { run_async_insert(std::move(job)); });
}

bool decide_sync(const hpsx_session* s, size_t n, uint32_t m) {
  if (s->insert_mode == 0) return false;
  if (s->insert_mode > 0) return true;
  // [UPSTREAM] hit_rate < hit_rate_threshold -> synchronous insertion
  const double hit_rate = 1.0 - static_cast<double>(m) / static_cast<double>(n);
  return hit_rate < static_cast<double>(s->model->cfg.hit_rate_threshold);
}

void account_probe_time(hpsx_session* s, size_t t, size_t n) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, s->ev[2 * t], s->ev[2 * t + 1]) == cudaSuccess) {
    s->stats.probe_kernel_ms += ms;
    ++s->stats.probe_kernel_launches;
    s->stats.probe_kernel_keys += n;
  }
}

}  // namespace

namespace hpsx {
namespace eng {

// Miss lists shorter than this are pulled in miss-list order: the resolve + radix sort costs ~70 us of launches,
// more than the sorted order saves on a few thousand rows.
size_t pull_sort_min() { return 16384; }

}  // namespace eng
}  // namespace hpsx

namespace hpsx {
namespace eng {

// Workspace of the address-sorted pull, allocated on first use.
int ensure_sort_workspace(hpsx_session* s) {
  if (s->d_sort_temp != nullptr) return HPSX_OK;
  const size_t cap = std::max<size_t>(s->cap_keys, 1);
  for (int i = 0; i < 2; ++i) {
    HPSX_CU(cudaMalloc(&s->d_addr[i], cap * sizeof(unsigned long long)));
    HPSX_CU(cudaMalloc(&s->d_sidx[i], cap * sizeof(uint32_t)));
  }
  s->sort_temp_bytes = sort_misses_temp_bytes(cap);
  HPSX_CU(cudaMalloc(&s->d_sort_temp, std::max<size_t>(s->sort_temp_bytes, 16)));
  return HPSX_OK;
}

}  // namespace eng
}  // namespace hpsx

namespace {

// ------------------------------------------------------------------------------------------------
// Binned direct pull: the miss path of enable_pagelock models with synchronous insertion.
//
//   stream A   [keys c0 ready] probe c0 | probe c1 | probe c2 | probe c3 | pull + insert (fused) | counters D2H
//   stream C   H2D keys c0 | c1 | c2 | c3          (host keys only; copy engine)
//
// A probe appends every miss to the list of its host-table partition (MissBins) and stores nothing for it; the
// pull kernel walks those lists bin by bin — the PCIe reads in flight stay inside one <= 256-MiB window of host
// memory, which is what the host link rewards (51 vs 39 GB/s, profiles/pcie_probe2_r02.txt) — writes the rows into
// the output and, holding the cache exclusively, inserts them in the same pass.  Nothing is sorted and no count
// crosses to the host in between: the whole request is enqueued at once and the host waits exactly once.
// A "group" = the probe launches that share one set of bins and one pull: all chunks of one table of a plain
// request, or table t of ALL requests of a batch (positions then carry the request index in their high bits).
// Measured and NOT done (profiles/r02_pipeline_timeline.txt): running the pull of chunk c beside the probe of
// chunk c+1.  Both kernels collapse when they share the GPU (probe 5x slower, pull at 20 GB/s), although a plain
// zero-copy gather and a plain HBM gather coexist (tools/pcie_probe3.cu); the serial order is faster.  Only
// requests whose vectors go to HOST memory keep one group per chunk: pull c then runs on stream B beside probe
// c+1, because the D2H copy of chunk c (stream D, 15 ms for a Criteo request) is what has to start early there.
// ------------------------------------------------------------------------------------------------
struct BinGroup {
  size_t table = 0;  // real table
  size_t n = 0;      // keys probing into the group
  MissBins bins;
  size_t count_off = 0;       // first word of the group's counts in s->d_bin_count
  float* out = nullptr;       // plain request: base of the table's output rows
  size_t row_off = 0;         // plain request: the chunk covers rows [row_off, row_off + n) of virtual table v
  size_t v = 0;
  std::vector<float*> outs;   // batch: one base per request
  void* out_bf16 = nullptr;
};

struct ProbeLaunch {
  size_t group = 0, v = 0;  // v = virtual table (request * T + table): timing events
  const int64_t* keys = nullptr;  // this slice, host or device
  size_t n = 0, stage_off = 0;
  float* out = nullptr;  // where row 0 of this slice goes
  uint32_t pos_base = 0;
  void* out_bf16 = nullptr;
  bool first_of_v = false, last_of_v = false;
};

int ensure_binned_workspace(hpsx_session* s, size_t count_words, size_t entries, size_t events) {
  if (!s->stream_b) {
    // the pull kernels are few CTAs that must get onto the SMs AHEAD of the thousands of queued probe CTAs of the next
    // chunk (the link, not the SMs, is the scarce resource): highest stream priority
    int lo = 0, hi = 0;
    HPSX_CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HPSX_CU(cudaStreamCreateWithPriority(&s->stream_b, cudaStreamNonBlocking, hi));
  }
  if (!s->stream_c) HPSX_CU(cudaStreamCreateWithFlags(&s->stream_c, cudaStreamNonBlocking));
  if (!s->stream_d) HPSX_CU(cudaStreamCreateWithFlags(&s->stream_d, cudaStreamNonBlocking));
  if (count_words > s->bin_count_cap) {
    if (s->d_bin_count) HPSX_CU(cudaFree(s->d_bin_count));
    if (s->h_bin_count) HPSX_CU(cudaFreeHost(s->h_bin_count));
    s->d_bin_count = nullptr;
    s->h_bin_count = nullptr;
    s->bin_count_cap = 0;
    const size_t cap = count_words + count_words / 2 + 64;
    HPSX_CU(cudaMalloc(&s->d_bin_count, cap * sizeof(uint32_t)));
    HPSX_CU(cudaMallocHost(&s->h_bin_count, cap * sizeof(uint32_t)));
    s->bin_count_cap = cap;
  }
  if (entries > s->bin_entry_cap) {
    if (s->d_bin_keys) HPSX_CU(cudaFree(s->d_bin_keys));
    if (s->d_bin_pos) HPSX_CU(cudaFree(s->d_bin_pos));
    s->d_bin_keys = nullptr;
    s->d_bin_pos = nullptr;
    s->bin_entry_cap = 0;
    const size_t cap = entries + entries / 8 + 1024;
    HPSX_CU(cudaMalloc(&s->d_bin_keys, cap * sizeof(int64_t)));
    HPSX_CU(cudaMalloc(&s->d_bin_pos, cap * sizeof(uint32_t)));
    s->bin_entry_cap = cap;
  }
  while (s->ev_chunk.size() < events) {
    cudaEvent_t e;
    HPSX_CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->ev_chunk.push_back(e);
  }
  return HPSX_OK;
}

// Is this request one the binned path serves?  (Everything else takes gpu_lookup_direct's general form.)
bool binned_path_applies(const hpsx_session* s, const size_t* n_v, size_t num_v, const uint32_t* const* pos_per_table) {
  const hpsx_cache* c = s->cache;
  const size_t T = s->model->tables.size();
  const bool always_sync = s->insert_mode > 0 || (s->insert_mode < 0 && s->model->cfg.hit_rate_threshold >= 1.0f);
  if (!always_sync || pos_per_table != nullptr) return false;
  for (size_t t = 0; t < T; ++t)
    if (c->tables[t].host_parts == 0 || c->tables[t].host_parts > 4096) return false;
  if (num_v > T) {
    if (num_v % T != 0 || num_v / T > static_cast<size_t>(kMaxBatchOuts) || s->bf16_out != nullptr) return false;
    for (size_t v = 0; v < num_v; ++v)
      if (n_v[v] >= (1ull << kShardPosBits)) return false;  // request-relative rows must fit below the request bits
  }
  for (size_t v = 0; v < num_v; ++v)
    if (n_v[v] > 0xFFFFFFFFull) return false;
  return true;
}

// Caller holds c->rw (exclusive, or shared when `split`) and c->pull_rw (shared).
int gpu_lookup_direct_binned(hpsx_session* s, const void* const* keys_v, bool keys_on_device, float* const* out_v,
                             const size_t* n_v, size_t num_v, uint32_t epoch, bool split,
                             std::shared_lock<std::shared_mutex>& rlock, std::unique_lock<std::shared_mutex>& wlock) {
  hpsx_cache* c = s->cache;
  const size_t T = s->model->tables.size();
  const bool batch = num_v > T;
  const size_t R = batch ? num_v / T : 1;
  const double tr0 = now_ms();

  // ---- plan: groups and probe launches
  std::vector<BinGroup> groups;
  std::vector<ProbeLaunch> launches;
  size_t stage_off = 0;
  if (batch) {
    for (size_t t = 0; t < T; ++t) {
      BinGroup g;
      g.table = t;
      g.outs.assign(R, nullptr);
      for (size_t r = 0; r < R; ++r) {
        const size_t v = r * T + t, n = n_v[v];
        g.outs[r] = out_v[v];
        if (n == 0) continue;
        ProbeLaunch L;
        L.group = groups.size();
        L.v = v;
        L.keys = static_cast<const int64_t*>(keys_v[v]);
        L.n = n;
        L.stage_off = stage_off;
        L.out = out_v[v];
        L.pos_base = static_cast<uint32_t>(r) << kShardPosBits;
        L.first_of_v = L.last_of_v = true;
        launches.push_back(L);
        stage_off += n;
        g.n += n;
      }
      if (g.n != 0) groups.push_back(std::move(g));
    }
  } else {
    for (size_t v = 0; v < num_v; ++v) {
      const size_t n = n_v[v];
      if (n == 0) continue;
      const size_t dim = c->tables[v].dim;
      // chunks: host keys are copied chunk by chunk so that the copy of chunk c+1 overlaps the probe of chunk c
      const bool chunked = n >= kPipelineMinKeys && (!keys_on_device || s->host_out != nullptr);
      const size_t K = chunked ? static_cast<size_t>(std::max(1, s->request_chunks)) : 1;
      const size_t csz = ((n + K - 1) / K + 31) / 32 * 32;
      const bool group_per_chunk = s->host_out != nullptr;
      for (size_t o = 0; o < n; o += csz) {
        const size_t nc = std::min(csz, n - o);
        if (o == 0 || group_per_chunk) {
          BinGroup g;
          g.table = v;
          g.out = out_v[v];
          g.row_off = o;
          g.v = v;
          g.out_bf16 = bf16_dst(s, v, 0, dim);
          groups.push_back(std::move(g));
        }
        groups.back().n += nc;
        ProbeLaunch L;
        L.group = groups.size() - 1;
        L.v = v;
        L.keys = static_cast<const int64_t*>(keys_v[v]) + o;
        L.n = nc;
        L.stage_off = stage_off;
        L.out = out_v[v] + o * dim;
        L.pos_base = static_cast<uint32_t>(o);
        L.out_bf16 = bf16_dst(s, v, o, dim);
        L.first_of_v = o == 0;
        L.last_of_v = o + csz >= n;
        launches.push_back(L);
        stage_off += nc;
      }
    }
  }
  const size_t G = groups.size();
  if (G > s->vt) return fail(HPSX_ERR_INTERNAL, "binned lookup: more groups than counters");
  size_t count_words = 0, entries = 0;
  for (BinGroup& g : groups) {
    const uint32_t P = c->tables[g.table].host_parts;
    g.count_off = count_words;
    count_words += P + 1;
    g.bins.num_bins = P;
    // uniform hashing keeps the bins within a few per cent of n/P; a bin that still overflows (one key repeated
    // thousands of times) spills into the group's overflow list, which can take every key of the group
    g.bins.bin_cap = static_cast<uint32_t>(g.n / P + g.n / (4 * P) + 256);
    g.bins.spill_cap = static_cast<uint32_t>(g.n);
    entries += static_cast<size_t>(P) * g.bins.bin_cap + g.bins.spill_cap;
  }
  {
    const int rc = ensure_binned_workspace(s, count_words, entries, launches.size() + 2 * G + 2);
    if (rc != HPSX_OK) return rc;
  }
  {
    size_t e = 0;
    for (BinGroup& g : groups) {
      g.bins.count = s->d_bin_count + g.count_off;
      g.bins.keys = s->d_bin_keys + e;
      g.bins.pos = s->d_bin_pos + e;
      e += static_cast<size_t>(g.bins.num_bins) * g.bins.bin_cap + g.bins.spill_cap;
    }
  }
  cudaStream_t A = s->stream, B = s->stream_b, C = s->stream_c;
  const bool host_out = s->host_out != nullptr && !batch;
  cudaEvent_t* ev_keys = s->ev_chunk.data();
  cudaEvent_t* ev_probe = ev_keys + launches.size();
  cudaEvent_t* ev_pulled = ev_probe + G;
  uint32_t* d_absent = s->d_counters + 2 * s->vt;
  uint32_t* d_inserted = s->d_counters + s->vt;

  // ---- enqueue
  const bool timeline = (s->debug_flags & 4) != 0;
  if (timeline) {
    while (s->ev_trace.size() < 6 * G + 1) {
      cudaEvent_t e;
      HPSX_CU(cudaEventCreate(&e));
      s->ev_trace.push_back(e);
    }
    HPSX_CU(cudaEventRecord(s->ev_trace[6 * G], A));
  }
  HPSX_CU(cudaMemsetAsync(s->d_bin_count, 0, count_words * sizeof(uint32_t), A));
  HPSX_CU(cudaMemsetAsync(d_absent, 0, G * sizeof(uint32_t), A));
  if (!keys_on_device) {
    for (size_t i = 0; i < launches.size(); ++i) {
      const ProbeLaunch& L = launches[i];
      HPSX_CU(cudaMemcpyAsync(s->d_keys + L.stage_off, L.keys, L.n * sizeof(int64_t), cudaMemcpyHostToDevice, C));
      HPSX_CU(cudaEventRecord(ev_keys[i], C));
      s->stats.h2d_bytes += L.n * sizeof(int64_t);
    }
  }
  // The pulling warp also inserts the row when this call owns the cache (one instance, dynamic cache) and no
  // probe can run beside the pull; otherwise (several instances: split lock; host output: pulls beside probes)
  // the rows are inserted from the output buffer in a separate pass.
  const bool overlap = host_out;  // pull of group g on stream B beside the probes of group g+1
  const bool fused = !c->is_static && !split && !overlap;
  cudaStream_t P = overlap ? B : A;  // where the pulls run
  // NVLink reads need ~2 MB in flight (latency x 700 GB/s) where PCIe needs ~130 KB: four CTAs per SM instead of one
  const int pull_ctas = c->tier.committed ? std::max(s->pull_grid_ctas, 148 * 4) : s->pull_grid_ctas;
  // Rows over PCIe: on a host table of a few GB the quad form of the pull kernel (four rows in flight per warp, half the
  // CTAs) reaches 50.2 GB/s where the one-row form reaches 48.5 (profiles/r02_summary.md); on tens of GB more rows in flight
  // spread the reads over more host memory and lose (engine.hpp pull_grid_ctas), so large tables keep the one-row form.
  auto small_host_table = [&](size_t t) {
    return s->model->tables[t]->rows() * s->model->tables[t]->dim() * sizeof(float) <= (8ull << 30);
  };
  size_t li = 0;
  for (size_t g = 0; g < G; ++g) {
    const BinGroup& grp = groups[g];
    const DeviceTable& dt = c->tables[grp.table];
    if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g], A));
    for (; li < launches.size() && launches[li].group == g; ++li) {
      const ProbeLaunch& L = launches[li];
      if (!keys_on_device) HPSX_CU(cudaStreamWaitEvent(A, ev_keys[li], 0));
      if (L.first_of_v) HPSX_CU(cudaEventRecord(s->ev[2 * L.v], A));
      HPSX_CU(launch_probe_gather(dt, keys_on_device ? L.keys : s->d_keys + L.stage_off, L.n, L.out, epoch, !c->is_static,
                                  nullptr, nullptr, nullptr, nullptr, s->probe_variant, A, nullptr, L.pos_base, L.out_bf16,
                                  &grp.bins, true));
      if (L.last_of_v) HPSX_CU(cudaEventRecord(s->ev[2 * L.v + 1], A));
      ++s->stats.kernel_launches;
    }
    if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g + 1], A));
    HPSX_CU(cudaEventRecord(ev_probe[g], A));
  }
  for (size_t g = 0; g < G; ++g) {
    const BinGroup& grp = groups[g];
    const DeviceTable& dt = c->tables[grp.table];
    // fused pulls rewrite cache slots: every probe of the request must be done (stream order on A gives that)
    if (overlap) HPSX_CU(cudaStreamWaitEvent(P, ev_probe[g], 0));
    if (g == 0) HPSX_CU(cudaEventRecord(s->ev_pull[0], P));
    if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g + 2], P));
    const bool quad_pcie = !c->tier.committed && !(s->debug_flags & 8) && small_host_table(grp.table);
    HPSX_CU(launch_pull_binned(dt, grp.bins, grp.out, grp.out_bf16, batch ? grp.outs.data() : nullptr,
                               batch ? static_cast<int>(R) : 0, d_absent + g, quad_pcie ? std::max(1, pull_ctas / 2) : pull_ctas, P,
                               fused ? 1 : 0, epoch, d_inserted + g, (c->tier.committed || quad_pcie) ? 4 : 1));
    if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g + 3], P));
    HPSX_CU(cudaEventRecord(ev_pulled[g], P));
    ++s->stats.kernel_launches;
    if (host_out) {
      // the chunk's rows are complete (hits by its probe, misses by its pull): start their trip to host memory now,
      // on the other direction of the link, while the next chunks are probed and pulled
      const size_t dim = dt.dim;
      HPSX_CU(cudaStreamWaitEvent(s->stream_d, ev_pulled[g], 0));
      HPSX_CU(cudaMemcpyAsync(s->host_out[grp.v] + grp.row_off * dim, grp.out + grp.row_off * dim, grp.n * dim * sizeof(float),
                              cudaMemcpyDeviceToHost, s->stream_d));
      s->stats.d2h_bytes += grp.n * dim * sizeof(float);
    }
  }
  HPSX_CU(cudaEventRecord(s->ev_pull[2], P));
  auto enqueue_inserts = [&]() -> int {
    for (size_t g = 0; g < G; ++g) {
      const BinGroup& grp = groups[g];
      if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g + 4], A));
      HPSX_CU(launch_insert_binned(c->tables[grp.table], grp.bins, grp.out, batch ? grp.outs.data() : nullptr,
                                   batch ? static_cast<int>(R) : 0, epoch, d_inserted + g, A));
      if (timeline) HPSX_CU(cudaEventRecord(s->ev_trace[6 * g + 5], A));
      ++s->stats.kernel_launches;
    }
    return HPSX_OK;
  };
  if (overlap && G > 0) HPSX_CU(cudaStreamWaitEvent(A, ev_pulled[G - 1], 0));  // B is in order: the last pull implies all
  if (!c->is_static && !split && !fused) {
    const int rc = enqueue_inserts();
    if (rc != HPSX_OK) return rc;
  }
  HPSX_CU(cudaEventRecord(s->ev_pull[1], A));
  HPSX_CU(cudaMemcpyAsync(s->h_bin_count, s->d_bin_count, count_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, A));
  HPSX_CU(cudaMemcpyAsync(s->h_counters + 2 * s->vt, d_absent, G * sizeof(uint32_t), cudaMemcpyDeviceToHost, A));
  HPSX_CU(cudaStreamSynchronize(A));
  if (host_out) {
    HPSX_CU(cudaStreamSynchronize(s->stream_d));
    s->host_out_done = true;
  }
  s->stats.d2h_bytes += (count_words + G) * sizeof(uint32_t);
  if (split && !c->is_static) {
    // several instances share the cache: the probes and pulls above ran under the SHARED lock; the rows they
    // brought are inserted in a short exclusive section
    rlock.unlock();
    wlock.lock();
    const int rc = enqueue_inserts();
    if (rc != HPSX_OK) return rc;
    HPSX_CU(cudaStreamSynchronize(A));
  }

  if (timeline && !c->is_static && !fused && !split) {
    auto at = [&](size_t i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s->ev_trace[6 * G], s->ev_trace[i]);
      return ms;
    };
    for (size_t g = 0; g < G; ++g)
      std::fprintf(stderr, "[hpsx] timeline group %zu (%zu keys): probe %.3f-%.3f | pull %.3f-%.3f | insert %.3f-%.3f ms\n", g,
                   groups[g].n, at(6 * g), at(6 * g + 1), at(6 * g + 2), at(6 * g + 3), at(6 * g + 4), at(6 * g + 5));
  }
  // ---- account
  for (const ProbeLaunch& L : launches)
    if (L.last_of_v) account_probe_time(s, L.v, n_v[L.v]);
  uint64_t total_keys = 0, total_miss = 0;
  for (size_t g = 0; g < G; ++g) {
    const BinGroup& grp = groups[g];
    const uint32_t* cnt = s->h_bin_count + grp.count_off;
    uint64_t m = std::min<uint32_t>(cnt[grp.bins.num_bins], grp.bins.spill_cap);
    for (uint32_t b = 0; b < grp.bins.num_bins; ++b) m += std::min<uint32_t>(cnt[b], grp.bins.bin_cap);
    const uint32_t absent = s->h_counters[2 * s->vt + g];
    s->stats.hits += grp.n - m;
    s->stats.misses += m;
    // rows pulled by the kernel: over PCIe from host memory, or from the NVLink tier's shards
    (c->tier.committed ? s->stats.tier_bytes : s->stats.h2d_bytes) += (m - absent) * c->tables[grp.table].dim * sizeof(float);
    s->stats.default_filled += absent;
    total_keys += grp.n;
    total_miss += m;
  }
  float ms = 0.f;
  if (total_miss != 0) {
    if (cudaEventElapsedTime(&ms, s->ev_pull[0], s->ev_pull[1]) == cudaSuccess) s->stats.insert_kernel_ms += ms;
    if (cudaEventElapsedTime(&ms, s->ev_pull[0], s->ev_pull[2]) == cudaSuccess) s->stats.pull_kernel_ms += ms;
  }
  if (total_keys != 0)
    s->miss_ratio = 0.5 * s->miss_ratio + 0.5 * static_cast<double>(total_miss) / static_cast<double>(total_keys);
  if (trace_on())
    std::fprintf(stderr, "[hpsx] binned lookup: %zu keys (%s), %zu group(s), %llu misses, host total %.3f ms\n", static_cast<size_t>(total_keys),
                 keys_on_device ? "device" : "host", G, static_cast<unsigned long long>(total_miss), now_ms() - tr0);
  return HPSX_OK;
}

// Direct-pull lookup (enable_pagelock): probe+gather, then the misses are resolved ON THE GPU: their rows are read
// straight from the page-locked host table over PCIe.  No CPU gather, no staging copy.  Requests with synchronous
// insertion (hit_rate_threshold >= 1, the samples' setting) take the binned pipeline above.  What remains here is
// the general form: insertion mode decided per request from the hit rate (the miss rows may then have to stay the
// default vector), and scattered delivery (`pos_per_table`, model-parallel return leg).  Large miss lists are
// resolved to host addresses and radix-sorted by host page first; short ones are pulled in list order with the miss
// count read on the device.
int gpu_lookup_direct(hpsx_session* s, const void* const* keys_per_table, bool keys_on_device,
                      float* const* out_per_table, const size_t* n_per_table, size_t num_tables,
                      const uint32_t* const* pos_per_table) {
  hpsx_cache* c = s->cache;
  const size_t T = s->model->tables.size();
  const uint32_t epoch = c->epoch.fetch_add(1, std::memory_order_relaxed);
  size_t total = 0;
  for (size_t t = 0; t < num_tables; ++t) total += n_per_table[t];
  ++s->stats.lookups;
  s->stats.keys += total;
  if (total == 0) return HPSX_OK;
  for (size_t t = 0; t < num_tables; ++t)
    if (n_per_table[t] != 0 && (!keys_per_table[t] || !out_per_table[t]))
      return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer for table " + std::to_string(t));
  // the pull kernels read the page-locked host rows and their HBM index: a database reload waits for them
  std::shared_lock<std::shared_mutex> pull_lock(c->pull_rw);
  std::unique_lock<std::shared_mutex> wlock(c->rw, std::defer_lock);
  std::shared_lock<std::shared_mutex> rlock(c->rw, std::defer_lock);

  if (s->model->tier_only() && pos_per_table == nullptr) {
    // Tables that live in the NVLink tier only (model-parallel rows, "synthetic_device:" tables): no local cache to
    // probe and nothing to insert — one gather kernel per table reads every row from its owner's shard.
    if (!c->tier.committed)
      return fail(HPSX_ERR_UNSUPPORTED, "model '" + s->model->cfg.model_name + "': its tables live in the NVLink tier, which is not attached");
    uint32_t* d_absent = s->d_counters + 2 * s->vt;
    if (num_tables > s->vt) return fail(HPSX_ERR_INTERNAL, "tier lookup: more tables than counters");
    HPSX_CU(cudaMemsetAsync(d_absent, 0, num_tables * sizeof(uint32_t), s->stream));
    HPSX_CU(cudaEventRecord(s->ev_pull[0], s->stream));
    size_t off = 0;
    for (size_t t = 0; t < num_tables; ++t) {
      const size_t n = n_per_table[t];
      if (n == 0) continue;
      const int64_t* d_keys = static_cast<const int64_t*>(keys_per_table[t]);
      if (!keys_on_device) {
        // pinned host keys (what Triton hands a backend) are read in place by the gather kernel, 256 B per warp tile
        // over PCIe while other warps' rows are in flight; pageable keys are staged first
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, keys_per_table[t]) == cudaSuccess && at.type == cudaMemoryTypeHost &&
            at.devicePointer != nullptr) {
          d_keys = static_cast<const int64_t*>(at.devicePointer);
        } else {
          cudaGetLastError();
          HPSX_CU(cudaMemcpyAsync(s->d_keys + off, keys_per_table[t], n * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
          d_keys = s->d_keys + off;
        }
        s->stats.h2d_bytes += n * sizeof(int64_t);
      }
      HPSX_CU(launch_tier_gather(c->tables[t % T], d_keys, n, out_per_table[t], d_absent + t, s->stream));
      ++s->stats.kernel_launches;
      off += n;
    }
    HPSX_CU(cudaEventRecord(s->ev_pull[1], s->stream));
    HPSX_CU(cudaMemcpyAsync(s->h_counters + 2 * s->vt, d_absent, num_tables * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    s->stats.d2h_bytes += num_tables * sizeof(uint32_t);
    for (size_t t = 0; t < num_tables; ++t) {
      const size_t n = n_per_table[t];
      if (n == 0) continue;
      const uint32_t absent = s->h_counters[2 * s->vt + t];
      s->stats.misses += n;  // of the (absent) local cache: every key is served by the tier
      s->stats.default_filled += absent;
      s->stats.tier_bytes += (n - absent) * c->tables[t % T].dim * sizeof(float);
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev_pull[0], s->ev_pull[1]) == cudaSuccess) {
      s->stats.insert_kernel_ms += ms;
      s->stats.pull_kernel_ms += ms;
    }
    return HPSX_OK;
  }

  if (binned_path_applies(s, n_per_table, num_tables, pos_per_table)) {
    // The insert pass rewrites cache slots: exclusive unless the cache is static (never inserts).  When several
    // instances share the cache (include/model_state.hpp:76-84) the call splits instead: probes and pulls under
    // the SHARED lock, then a short exclusive section for the inserts — one instance's probes run beside
    // another instance's pull.
    const bool split = s->model->split_lock && !c->is_static && c->sessions.load(std::memory_order_relaxed) > 1;
    if (c->is_static || split) rlock.lock(); else wlock.lock();
    return gpu_lookup_direct_binned(s, keys_per_table, keys_on_device, out_per_table, n_per_table, num_tables, epoch, split,
                                    rlock, wlock);
  }

  // A small request whose predicted miss count (from this session's recent miss ratio) is below the sort threshold
  // would pull in miss-list order anyway: it then takes the device-driven form — probe and pull back to back, the
  // pull reads the miss count on the device, ONE host wait — instead of reading the count back first.
  const bool speculative = num_tables <= T && total < kPipelineMinKeys && pos_per_table == nullptr &&
                           s->miss_ratio * static_cast<double>(total) < static_cast<double>(pull_sort_min());
  const bool sorted = !speculative;
  if (sorted) {
    const int rc = ensure_sort_workspace(s);
    if (rc != HPSX_OK) return rc;
  }
  const double tr0 = now_ms();
  // the fused pull kernel rewrites cache slots: exclusive unless the cache is static
  if (c->is_static) rlock.lock(); else wlock.lock();

  HPSX_CU(cudaMemsetAsync(s->d_counters, 0, s->vt * sizeof(uint32_t), s->stream));
  HPSX_CU(cudaMemsetAsync(s->d_counters + 2 * s->vt, 0, s->vt * sizeof(uint32_t), s->stream));
  std::vector<size_t> off(num_tables + 1, 0);
  auto pull = [&](size_t t, size_t m_hint) -> cudaError_t {
    const bool use_sorted = sorted && m_hint >= std::max<size_t>(pull_sort_min(), 1);
    return launch_pull_misses(c->tables[t % T], s->d_miss_keys + off[t], s->d_miss_pos + off[t], s->d_counters + t,
                              n_per_table[t], out_per_table[t], nullptr, !c->is_static, s->insert_mode,
                              s->model->cfg.hit_rate_threshold, epoch, s->d_counters + s->vt + t,
                              s->d_counters + 2 * s->vt + t, use_sorted ? s->d_addr[1] + off[t] : nullptr,
                              use_sorted ? s->d_sidx[1] + off[t] : nullptr, m_hint, s->stream, 0,
                              bf16_dst(s, t, 0, c->tables[t % T].dim), nullptr);
  };
  for (size_t t = 0; t < num_tables; ++t) {
    const size_t n = n_per_table[t];
    off[t + 1] = off[t] + n;
    if (n == 0) continue;
    const int64_t* d_keys;
    const size_t dim = c->tables[t % T].dim;
    if (keys_on_device) {
      d_keys = static_cast<const int64_t*>(keys_per_table[t]);
    } else {
      HPSX_CU(cudaMemcpyAsync(s->d_keys + off[t], keys_per_table[t], n * sizeof(int64_t),
                              cudaMemcpyHostToDevice, s->stream));
      s->stats.h2d_bytes += n * sizeof(int64_t);
      d_keys = s->d_keys + off[t];
    }
    HPSX_CU(cudaEventRecord(s->ev[2 * t], s->stream));
    HPSX_CU(launch_probe_gather(c->tables[t % T], d_keys, n, out_per_table[t], epoch, !c->is_static,
                                s->d_counters + t, s->d_miss_pos + off[t], s->d_miss_keys + off[t], nullptr,
                                s->probe_variant, s->stream, pos_per_table ? pos_per_table[t] : nullptr, 0,
                                bf16_dst(s, t, 0, dim)));
    HPSX_CU(cudaEventRecord(s->ev[2 * t + 1], s->stream));
    ++s->stats.kernel_launches;
    if (!sorted) {
      HPSX_CU(pull(t, 0));
      HPSX_CU(cudaEventRecord(s->ev_pull[2 * t + 1], s->stream));
      ++s->stats.kernel_launches;
    }
  }
  if (sorted) {
    HPSX_CU(cudaMemcpyAsync(s->h_counters, s->d_counters, s->vt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    s->stats.d2h_bytes += T * sizeof(uint32_t);
    for (size_t t = 0; t < num_tables; ++t)
      if (n_per_table[t] != 0) account_probe_time(s, t, n_per_table[t]);
    {
      // no table missed anything: the rows are all delivered, nothing to pull, no second host wait
      bool any_miss = false;
      for (size_t t = 0; t < num_tables; ++t) any_miss = any_miss || (n_per_table[t] != 0 && s->h_counters[t] != 0);
      if (!any_miss) {
        for (size_t t = 0; t < num_tables; ++t) s->stats.hits += n_per_table[t];
        s->miss_ratio *= 0.5;
        return HPSX_OK;
      }
    }
    NvtxRange miss_range("hpsx_direct_pull_misses");
    for (size_t t = 0; t < num_tables; ++t) {
      const uint32_t m = n_per_table[t] ? s->h_counters[t] : 0;
      if (m == 0) continue;
      HPSX_CU(cudaEventRecord(s->ev_pull[2 * t], s->stream));
      if (m >= std::max<size_t>(pull_sort_min(), 1)) {
        HPSX_CU(launch_resolve_and_sort_misses(c->tables[t % T], s->d_miss_keys + off[t], m, s->d_addr[0] + off[t],
                                               s->d_sidx[0] + off[t], s->d_addr[1] + off[t], s->d_sidx[1] + off[t],
                                               s->d_sort_temp, s->sort_temp_bytes, s->stream));
        ++s->stats.kernel_launches;  // resolve (the CUB radix-sort passes are library kernels, not counted)
      }
      HPSX_CU(pull(t, m));
      HPSX_CU(cudaEventRecord(s->ev_pull[2 * t + 1], s->stream));
      ++s->stats.kernel_launches;
    }
  }
  HPSX_CU(cudaMemcpyAsync(s->h_counters + s->vt, s->d_counters + s->vt, 2 * s->vt * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          s->stream));
  if (!sorted)
    HPSX_CU(cudaMemcpyAsync(s->h_counters, s->d_counters, s->vt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  HPSX_CU(cudaStreamSynchronize(s->stream));
  s->stats.d2h_bytes += 3 * T * sizeof(uint32_t);
  for (size_t t = 0; t < num_tables; ++t) {
    const size_t n = n_per_table[t];
    if (n == 0) continue;
    const uint32_t m = s->h_counters[t];
    float ms = 0.f;
    if (sorted) {
      if (m != 0 && cudaEventElapsedTime(&ms, s->ev_pull[2 * t], s->ev_pull[2 * t + 1]) == cudaSuccess)
        s->stats.insert_kernel_ms += ms;  // resolve + sort + pull
    } else {
      account_probe_time(s, t, n);
      if (cudaEventElapsedTime(&ms, s->ev[2 * t + 1], s->ev_pull[2 * t + 1]) == cudaSuccess) s->stats.insert_kernel_ms += ms;
    }
    s->stats.hits += n - m;
    s->stats.misses += m;
    const size_t row_bytes = s->model->tables[t % T]->dim() * sizeof(float);
    const uint32_t absent = s->h_counters[2 * s->vt + t];
    (c->tier.committed ? s->stats.tier_bytes : s->stats.h2d_bytes) += static_cast<uint64_t>(m - absent) * row_bytes;  // rows pulled by the kernel (PCIe, or the NVLink tier)
    s->stats.default_filled += (m != 0 && !decide_sync(s, n, m)) ? m : absent;
    s->miss_ratio = 0.5 * s->miss_ratio + 0.5 * static_cast<double>(m) / static_cast<double>(n);
    if (trace_on()) {
      float p_ms = 0.f, q_ms = 0.f;
      cudaEventElapsedTime(&p_ms, s->ev[2 * t], s->ev[2 * t + 1]);
      if (sorted && m != 0) cudaEventElapsedTime(&q_ms, s->ev_pull[2 * t], s->ev_pull[2 * t + 1]);
      std::fprintf(stderr, "[hpsx] direct lookup n=%zu (%s keys): probe %.3f ms | %u misses, resolve+sort+pull %.3f ms | host total %.3f ms\n",
                   n, keys_on_device ? "device" : "host", p_ms, m, q_ms, now_ms() - tr0);
    }
  }
  return HPSX_OK;
}

// `pos_per_table` (nullable, device memory): key i of table t is delivered to row pos[t][i] of out[t]
// instead of row i (model-parallel return leg; out[t] may be a peer GPU's buffer).
int gpu_lookup(hpsx_session* s, const void* const* keys_per_table, bool keys_on_device,
               float* const* out_per_table, const size_t* n_per_table, size_t num_tables,
               const uint32_t* const* pos_per_table = nullptr) {
  hpsx_cache* c = s->cache;
  NvtxRange range("hpsx_lookup");
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  if (c->direct_pull)
    return gpu_lookup_direct(s, keys_per_table, keys_on_device, out_per_table, n_per_table, num_tables,
                             pos_per_table);
  const size_t T = s->model->tables.size();
  const uint32_t epoch = c->epoch.fetch_add(1, std::memory_order_relaxed);
  size_t total = 0;
  for (size_t t = 0; t < num_tables; ++t) total += n_per_table[t];
  ++s->stats.lookups;
  s->stats.keys += total;
  if (total == 0) return HPSX_OK;
  const double tr0 = now_ms();

  HPSX_CU(cudaMemsetAsync(s->d_counters, 0, s->vt * sizeof(uint32_t), s->stream));
  std::vector<size_t> off(num_tables + 1, 0);
  {
    std::shared_lock<std::shared_mutex> lk(c->rw);
    for (size_t t = 0; t < num_tables; ++t) {
      const size_t n = n_per_table[t];
      off[t + 1] = off[t] + n;
      if (n == 0) continue;
      if (!keys_per_table[t] || !out_per_table[t])
        return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer for table " + std::to_string(t));
      const int64_t* d_keys;
      if (keys_on_device) {
        d_keys = static_cast<const int64_t*>(keys_per_table[t]);
      } else {
        HPSX_CU(cudaMemcpyAsync(s->d_keys + off[t], keys_per_table[t], n * sizeof(int64_t),
                                cudaMemcpyHostToDevice, s->stream));
        s->stats.h2d_bytes += n * sizeof(int64_t);
        d_keys = s->d_keys + off[t];
      }
      HPSX_CU(cudaEventRecord(s->ev[2 * t], s->stream));
      HPSX_CU(launch_probe_gather(c->tables[t % T], d_keys, n, out_per_table[t], epoch, !c->is_static,
                                  s->d_counters + t, s->d_miss_pos + off[t], s->d_miss_keys + off[t],
                                  s->hd_miss_keys + off[t], s->probe_variant, s->stream,
                                  pos_per_table ? pos_per_table[t] : nullptr, 0,
                                  bf16_dst(s, t, 0, c->tables[t % T].dim)));
      HPSX_CU(cudaEventRecord(s->ev[2 * t + 1], s->stream));
      ++s->stats.kernel_launches;
    }
    HPSX_CU(cudaMemcpyAsync(s->h_counters, s->d_counters, s->vt * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, s->stream));
    // one synchronisation: miss counts (copied) and miss keys (written by the kernel straight into
    // mapped pinned memory) are both on the host after it
    HPSX_CU(cudaStreamSynchronize(s->stream));
  }
  s->stats.d2h_bytes += T * sizeof(uint32_t);
  for (size_t t = 0; t < num_tables; ++t)
    if (n_per_table[t] != 0) account_probe_time(s, t, n_per_table[t]);

  const double tr1 = now_ms();
  const double gather0 = s->stats.host_gather_ms;
  bool any_sync = false;
  std::unique_lock<std::shared_mutex> wlock(c->rw, std::defer_lock);
  cudaEvent_t e0 = s->ev[0], e1 = s->ev[1];
  for (size_t t = 0; t < num_tables; ++t) {
    const uint32_t m = s->h_counters[t];
    const size_t n = n_per_table[t];
    s->stats.hits += n - m;
    if (m == 0) continue;
    s->stats.d2h_bytes += static_cast<uint64_t>(m) * sizeof(int64_t);  // zero-copy miss-key mirror
    if (!decide_sync(s, n, m)) {
      post_async_insert(s, t, off[t], m);
      continue;
    }
    if (!any_sync) HPSX_CU(cudaEventRecord(e0, s->stream));
    any_sync = true;
    const int rc = stream_miss_rows(s, t, off[t], m, out_per_table[t], !c->is_static, epoch, nullptr,
                                    &wlock);
    if (rc != HPSX_OK) return rc;
  }
  const double tr2 = now_ms();
  if (any_sync) {
    HPSX_CU(cudaEventRecord(e1, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) s->stats.insert_kernel_ms += ms;
  }
  if (trace_on())
    std::fprintf(stderr,
                 "[hpsx] lookup n=%zu probe+sync %.3f ms | miss loop %.3f ms (host gather %.3f) | "
                 "final sync %.3f ms | misses[0]=%u\n",
                 total, tr1 - tr0, tr2 - tr1, s->stats.host_gather_ms - gather0, now_ms() - tr2,
                 s->h_counters[0]);
  return HPSX_OK;
}

}  // namespace

namespace hpsx {
namespace eng {

// The per-call stage that keeps all miss rows of one request (pooled path, model-parallel groups).
// Stream-ordered (cudaMallocAsync/cudaFreeAsync): growing it never synchronises the device, which matters when
// another rank's flag-wait kernel is resident on the same GPU.
int ensure_pool_stage(hpsx_session* s, size_t m) {
  if (s->pool_stage_rows >= m) return HPSX_OK;
  if (s->d_pool_stage) HPSX_CU(cudaFreeAsync(s->d_pool_stage, s->stream));
  s->d_pool_stage = nullptr;
  s->pool_stage_rows = 0;
  const size_t rows = std::max(m, std::min<size_t>(2 * m, std::max<size_t>(s->cap_keys, m)));
  HPSX_CU(cudaMallocAsync(reinterpret_cast<void**>(&s->d_pool_stage), rows * s->max_dim * sizeof(float), s->stream));
  s->pool_stage_rows = rows;
  return HPSX_OK;
}

}  // namespace eng
}  // namespace hpsx

namespace {

int gpu_lookup_pooled(hpsx_session* s, size_t table, const int64_t* keys, bool keys_on_device,
                      size_t num_bags, size_t hotness, int combiner, float* d_pooled) {
  NvtxRange range("hpsx_lookup_pooled");
  hpsx_cache* c = s->cache;
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  const size_t T = s->model->tables.size();
  const size_t n = num_bags * hotness;
  ++s->stats.lookups;
  s->stats.keys += n;
  if (n == 0) return HPSX_OK;
  if (!keys || !d_pooled) return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer");
  if (!s->d_src) HPSX_CU(cudaMalloc(&s->d_src, s->cap_keys * sizeof(uint32_t)));
  const uint32_t epoch = c->epoch.fetch_add(1, std::memory_order_relaxed);
  std::shared_lock<std::shared_mutex> pull_lock(c->pull_rw, std::defer_lock);
  if (c->direct_pull) pull_lock.lock();  // the pull below reads the page-locked host rows and their HBM index
  HPSX_CU(cudaMemsetAsync(s->d_counters, 0, s->vt * sizeof(uint32_t), s->stream));
  const int64_t* d_keys = keys;
  uint32_t m = 0;
  {
    // The slot indices recorded by the probe stay valid only while no kernel rewrites cache slots:
    // the shared lock is held from the probe until the pooled gather has finished, and the rows of
    // the misses are inserted only afterwards.
    std::shared_lock<std::shared_mutex> lk(c->rw);
    if (!keys_on_device) {
      HPSX_CU(cudaMemcpyAsync(s->d_keys, keys, n * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
      s->stats.h2d_bytes += n * sizeof(int64_t);
      d_keys = s->d_keys;
    }
    HPSX_CU(launch_probe_index(c->tables[table], d_keys, n, epoch, !c->is_static, s->d_src,
                               s->d_counters + table, s->d_miss_pos, s->d_miss_keys, s->hd_miss_keys,
                               s->stream));
    ++s->stats.kernel_launches;
    HPSX_CU(cudaMemcpyAsync(s->h_counters, s->d_counters, s->vt * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
    s->stats.d2h_bytes += T * sizeof(uint32_t);
    m = s->h_counters[table];
    s->stats.hits += n - m;
    if (m > 0) {
      // the pooled sum needs every row: misses are always fetched before pooling
      s->stats.d2h_bytes += static_cast<uint64_t>(m) * sizeof(int64_t);
      {
        const int prc = ensure_pool_stage(s, m);
        if (prc != HPSX_OK) return prc;
      }
      if (c->direct_pull) {
        // rows pulled by the GPU straight from the page-locked host table into the stage
        HPSX_CU(cudaMemsetAsync(s->d_counters + 2 * s->vt, 0, s->vt * sizeof(uint32_t), s->stream));
        const bool use_sorted = m >= pull_sort_min();
        if (use_sorted) {
          const int wrc = ensure_sort_workspace(s);
          if (wrc != HPSX_OK) return wrc;
          HPSX_CU(launch_resolve_and_sort_misses(c->tables[table], s->d_miss_keys, m, s->d_addr[0], s->d_sidx[0],
                                                 s->d_addr[1], s->d_sidx[1], s->d_sort_temp, s->sort_temp_bytes,
                                                 s->stream));
          ++s->stats.kernel_launches;
        }
        HPSX_CU(launch_pull_misses(c->tables[table], s->d_miss_keys, s->d_miss_pos, s->d_counters + table, n,
                                   nullptr, s->d_pool_stage, false, 1, 0.f, epoch, nullptr,
                                   s->d_counters + 2 * s->vt + table, use_sorted ? s->d_addr[1] : nullptr,
                                   use_sorted ? s->d_sidx[1] : nullptr, m, s->stream));
        ++s->stats.kernel_launches;
        s->stats.misses += m;
        s->stats.h2d_bytes += static_cast<uint64_t>(m) * s->model->tables[table]->dim() * sizeof(float);
      } else {
        const int rc = stream_miss_rows(s, table, 0, m, nullptr, false, epoch, s->d_pool_stage, nullptr);
        if (rc != HPSX_OK) return rc;
      }
    }
    HPSX_CU(cudaEventRecord(s->ev[2 * table], s->stream));
    HPSX_CU(launch_pooled_gather(c->tables[table], s->d_src, s->d_pool_stage, num_bags, hotness,
                                 combiner == HPSX_COMBINER_MEAN, d_pooled, s->stream));
    HPSX_CU(cudaEventRecord(s->ev[2 * table + 1], s->stream));
    ++s->stats.kernel_launches;
    HPSX_CU(cudaStreamSynchronize(s->stream));
  }
  account_probe_time(s, table, n);
  if (m > 0 && !c->is_static) {
    std::unique_lock<std::shared_mutex> wlock(c->rw);
    HPSX_CU(launch_insert_merge(c->tables[table], s->d_miss_keys, s->d_miss_pos, s->d_pool_stage, m,
                                nullptr, true, epoch, s->d_counters + s->vt + table, s->stream));
    ++s->stats.kernel_launches;
    HPSX_CU(cudaStreamSynchronize(s->stream));
  }
  return HPSX_OK;
}


}  // namespace

hpsx_cache::~hpsx_cache() {
  if (device >= 0) {
    DeviceGuard guard(device);
    {
      std::unique_lock<std::mutex> lk(async_mu);
      async_cv.wait(lk, [this] { return async_pending == 0; });
    }
    for (auto& t : tables) {
      if (t.buckets) cudaFree(t.buckets);
      if (t.values) cudaFree(t.values);
    }
    hpsx::eng::tier_release(this);
    for (hpsx::IndexSlot* ix : indexes) cudaFree(ix);
    if (async_d_keys) cudaFree(async_d_keys);
    if (async_d_stage) cudaFree(async_d_stage);
    if (async_h_stage) cudaFreeHost(async_h_stage);
    if (async_stream) cudaStreamDestroy(async_stream);
  }
}

hpsx_session::~hpsx_session() {
  if (cache) cache->sessions.fetch_sub(1, std::memory_order_relaxed);
  if (device >= 0 && cache) {
    DeviceGuard guard(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(d_keys);
    cudaFree(d_miss_pos);
    cudaFree(d_miss_keys);
    cudaFree(d_counters);
    cudaFree(d_src);
    for (int i = 0; i < 2; ++i) {
      cudaFree(d_addr[i]);
      cudaFree(d_sidx[i]);
    }
    cudaFree(d_sort_temp);
    cudaFree(d_bin_count);
    cudaFree(d_bin_keys);
    cudaFree(d_bin_pos);
    if (h_bin_count) cudaFreeHost(h_bin_count);
    cudaFree(d_result);
    cudaFree(d_pool_stage);
    if (h_counters) cudaFreeHost(h_counters);
    if (h_miss_keys) cudaFreeHost(h_miss_keys);
    for (int b = 0; b < hpsx::kNumStages; ++b) {
      if (h_stage[b]) cudaFreeHost(h_stage[b]);
      cudaFree(d_stage[b]);
      if (stage_free[b]) cudaEventDestroy(stage_free[b]);
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_pull) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_chunk) cudaEventDestroy(e);
    if (stream_b) cudaStreamDestroy(stream_b);
    if (stream_c) cudaStreamDestroy(stream_c);
    if (stream_d) cudaStreamDestroy(stream_d);
    if (stream) cudaStreamDestroy(stream);
  }
}


// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int hpsx_abi_version(void) { return HPSX_ABI_VERSION; }
const char* hpsx_last_error(void) { return g_err.c_str(); }

int hpsx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int hpsx_ps_create(const hpsx_volatile_params* vdb, hpsx_ps** out) {
  HPSX_GUARD_BEGIN
  if (!out) return fail(HPSX_ERR_INVALID_ARG, "null output handle");
  std::unique_ptr<hpsx_ps> ps(new hpsx_ps());
  size_t threads = 0;
  if (vdb) {
    ps->vdb.num_partitions = vdb->num_partitions;
    if (vdb->allocation_rate) ps->vdb.allocation_rate = vdb->allocation_rate;
    ps->vdb.hpsx_pull_window_mb = vdb->pull_window_bytes >> 20;
    if (vdb->initial_cache_rate > 0) ps->vdb.initial_cache_rate = vdb->initial_cache_rate;
    threads = vdb->num_threads;
  }
  ps->vdb.num_partitions = resolve_partitions(ps->vdb.num_partitions);
  ps->pool.reset(new ThreadPool(threads));
  *out = ps.release();
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_destroy(hpsx_ps* ps) {
  HPSX_GUARD_BEGIN
  if (!ps) return HPSX_OK;
  // caches wait for their asynchronous jobs, which run on the pool: destroy them first
  for (auto& kv : ps->models) kv.second->caches.clear();
  delete ps;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_num_models(const hpsx_ps* ps, size_t* out) {
  if (!ps || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = ps->model_order.size();
  return HPSX_OK;
}

int hpsx_ps_model_name(const hpsx_ps* ps, size_t index, const char** out) {
  if (!ps || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (index >= ps->model_order.size()) return fail(HPSX_ERR_NOT_FOUND, "model index out of range");
  *out = ps->model_order[index].c_str();
  return HPSX_OK;
}

int hpsx_ps_has_model(const hpsx_ps* ps, const char* model_name) {
  return find_model(const_cast<hpsx_ps*>(ps), model_name) != nullptr ? 1 : 0;
}

static int add_model_cfg(hpsx_ps* ps, const ModelConfig& cfg, float load_factor) {
  const size_t T = cfg.embedding_vecsize_per_table.size();
  if (cfg.model_name.empty()) return fail(HPSX_ERR_INVALID_ARG, "model name is empty");
  if (T == 0) return fail(HPSX_ERR_INVALID_ARG, "model '" + cfg.model_name + "' has no tables");
  if (cfg.maxnum_catfeature_query_per_table_per_sample.size() != T ||
      cfg.default_value_for_each_table.size() != T)
    return fail(HPSX_ERR_INVALID_ARG, "model '" + cfg.model_name + "': per-table lists differ in length");
  if (cfg.embedding_cache_type == CacheType::UVM)
    return fail(HPSX_ERR_UNSUPPORTED, "embedding_cache_type 'uvm' is not supported");
  std::unique_ptr<Model> m(new Model());
  m->cfg = cfg;
  m->load_factor = load_factor > 0.f ? load_factor : 0.5f;
  m->direct_pull = cfg.enable_pagelock;
  m->split_lock = cfg.hpsx_split_lock;
  m->request_chunks = cfg.hpsx_request_chunks > 0 ? std::min<int>(cfg.hpsx_request_chunks, kMaxBatchRequests) : 4;
  m->pull_grid_ctas = cfg.hpsx_pull_grid_ctas > 0 ? std::min(cfg.hpsx_pull_grid_ctas, 148 * 8) : 148;
  m->probe_variant = cfg.hpsx_probe == "ldg" ? kProbeLdg : cfg.hpsx_probe == "tma" ? kProbeTma : kProbeV8;
  m->peer_tier = cfg.hpsx_peer_tier;
  for (size_t t = 0; t < T; ++t) {
    if (cfg.embedding_vecsize_per_table[t] == 0)
      return fail(HPSX_ERR_INVALID_ARG, "embedding_vecsize_per_table must be > 0");
    m->tables.emplace_back(new HostTable(cfg.embedding_vecsize_per_table[t],
                                         cfg.default_value_for_each_table[t], ps->vdb.num_partitions,
                                         ps->vdb.allocation_rate, ps->vdb.hpsx_pull_window_mb << 20));
  }
  m->device_rows.assign(T, 0);
  m->device_seed.assign(T, 0);
  for (size_t t = 0; t < T && t < cfg.sparse_files.size(); ++t) {
    if (cfg.sparse_files[t].empty()) continue;
    if (cfg.sparse_files[t].rfind("synthetic_device:", 0) == 0) {
      unsigned long long rows = 0, seed = 0;
      if (std::sscanf(cfg.sparse_files[t].c_str(), "synthetic_device:rows=%llu,seed=%llu", &rows, &seed) != 2 || rows == 0)
        return fail(HPSX_ERR_INVALID_ARG, "malformed table spec '" + cfg.sparse_files[t] +
                                              "' (want synthetic_device:rows=<N>,seed=<S>)");
      if (!cfg.hpsx_peer_tier || !cfg.enable_pagelock || !cfg.use_gpu_embedding_cache)
        return fail(HPSX_ERR_INVALID_ARG, "model '" + cfg.model_name + "': a synthetic_device table lives in the NVLink tier only; "
                                              "it needs gpucache, enable_pagelock and hpsx_peer_tier");
      m->device_rows[t] = rows;
      m->device_seed[t] = seed;
      continue;
    }
    const int rc = load_sparse_dir(ps, m->tables[t].get(), cfg.sparse_files[t]);
    if (rc != HPSX_OK) return rc;
  }
  for (size_t t = 0; t < T; ++t) {
    m->table_names.push_back(t < cfg.embedding_table_names.size()
                                 ? cfg.embedding_table_names[t]
                                 : "sparse_embedding" + std::to_string(t));  // docs default name
  }
  m->cfg.sparse_files.resize(T);
  for (size_t t = 0; t < T; ++t) {
    m->c_sparse_files.push_back(m->cfg.sparse_files[t].c_str());
    m->c_table_names.push_back(m->table_names[t].c_str());
  }
  std::lock_guard<std::mutex> lk(ps->mu);
  if (ps->models.count(cfg.model_name))
    return fail(HPSX_ERR_INVALID_ARG, "model '" + cfg.model_name + "' is already registered");
  ps->model_order.push_back(cfg.model_name);
  ps->models[cfg.model_name] = std::move(m);
  return HPSX_OK;
}

int hpsx_ps_add_model(hpsx_ps* ps, const hpsx_model_params* p) {
  HPSX_GUARD_BEGIN
  if (!ps || !p || !p->model_name) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (p->num_tables == 0 || !p->embedding_vecsize_per_table ||
      !p->maxnum_catfeature_query_per_table_per_sample || !p->default_value_for_each_table)
    return fail(HPSX_ERR_INVALID_ARG, "per-table parameter arrays are required");
  ModelConfig cfg;
  cfg.model_name = p->model_name;
  cfg.max_batch_size = p->max_batch_size;
  for (size_t t = 0; t < p->num_tables; ++t) {
    cfg.sparse_files.push_back(p->sparse_files && p->sparse_files[t] ? p->sparse_files[t] : "");
    if (p->table_names && p->table_names[t]) cfg.embedding_table_names.push_back(p->table_names[t]);
    cfg.embedding_vecsize_per_table.push_back(p->embedding_vecsize_per_table[t]);
    cfg.maxnum_catfeature_query_per_table_per_sample.push_back(
        p->maxnum_catfeature_query_per_table_per_sample[t]);
    cfg.default_value_for_each_table.push_back(p->default_value_for_each_table[t]);
  }
  cfg.use_gpu_embedding_cache = p->use_gpu_embedding_cache != 0;
  cfg.hit_rate_threshold = p->hit_rate_threshold;
  cfg.cache_size_percentage = p->cache_size_percentage;
  cfg.number_of_worker_buffers_in_pool = p->number_of_worker_buffers_in_pool;
  for (size_t i = 0; i < p->num_deployed_devices; ++i)
    cfg.deployed_devices.push_back(p->deployed_devices[i]);
  if (cfg.deployed_devices.empty()) cfg.deployed_devices.push_back(0);
  cfg.device_id = cfg.deployed_devices.back();
  cfg.embedding_cache_type = p->embedding_cache_type == HPSX_CACHE_STATIC ? CacheType::Static
                                                                          : CacheType::Dynamic;
  cfg.enable_pagelock = p->enable_pagelock != 0;
  cfg.hpsx_split_lock = p->split_lock >= 0;
  cfg.hpsx_request_chunks = p->request_chunks;
  cfg.hpsx_pull_grid_ctas = p->pull_grid_ctas;
  cfg.hpsx_probe = p->probe_variant == kProbeLdg && p->probe_variant_set ? "ldg" : p->probe_variant == kProbeTma ? "tma" : "";
  cfg.hpsx_peer_tier = p->peer_tier != 0;
  return add_model_cfg(ps, cfg, p->cache_load_factor);
  HPSX_GUARD_END
}

int hpsx_ps_create_from_json(const char* ps_json_path, hpsx_ps** out) {
  HPSX_GUARD_BEGIN
  if (!ps_json_path || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  PsConfig cfg;
  const ParseResult pr = parse_ps_config_file(ps_json_path, &cfg);
  if (!pr.ok) return fail(HPSX_ERR_INVALID_ARG, pr.message);
  switch (cfg.volatile_db.type) {
    case DatabaseType::HashMap:
    case DatabaseType::ParallelHashMap:
      break;
    default:
      return fail(HPSX_ERR_UNSUPPORTED,
                  std::string("volatile_db.type '") + to_string(cfg.volatile_db.type) +
                      "' is not supported: this engine serves the hash_map / parallel_hash_map "
                      "host database only");
  }
  if (cfg.persistent_db.type != DatabaseType::Disabled)
    return fail(HPSX_ERR_UNSUPPORTED, "persistent_db is not supported (type must be 'disabled')");
  hpsx_volatile_params vp{};
  vp.num_partitions = cfg.volatile_db.num_partitions;
  vp.allocation_rate = cfg.volatile_db.allocation_rate;
  vp.pull_window_bytes = cfg.volatile_db.hpsx_pull_window_mb << 20;
  vp.initial_cache_rate = cfg.volatile_db.initial_cache_rate;
  hpsx_ps* ps = nullptr;
  int rc = hpsx_ps_create(&vp, &ps);
  if (rc != HPSX_OK) return rc;
  ps->vdb = cfg.volatile_db;
  ps->vdb.num_partitions = resolve_partitions(cfg.volatile_db.num_partitions);
  for (const ModelConfig& m : cfg.models) {
    rc = add_model_cfg(ps, m, 0.f);
    if (rc == HPSX_OK && m.use_gpu_embedding_cache && m.init_ec)
      rc = hpsx_ps_create_embedding_cache_per_model(ps, m.model_name.c_str());
    if (rc != HPSX_OK) {
      const std::string keep = g_err;
      hpsx_ps_destroy(ps);
      return fail(rc, keep);
    }
  }
  *out = ps;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_load_table(hpsx_ps* ps, const char* model, size_t table, const int64_t* h_keys,
                       const float* h_vectors, size_t num_rows) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  if (table >= m->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (num_rows > 0 && (!h_keys || !h_vectors)) return fail(HPSX_ERR_INVALID_ARG, "null rows");
  m->tables[table]->insert(h_keys, h_vectors, num_rows, *ps->pool);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_load_table_procedural(hpsx_ps* ps, const char* model, size_t table, size_t num_rows,
                                  uint64_t seed) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  if (table >= m->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  m->tables[table]->fill_procedural(num_rows, seed, *ps->pool);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_load_table_procedural_shard(hpsx_ps* ps, const char* model, size_t table, size_t num_rows,
                                        uint64_t seed, uint32_t shard, uint32_t num_shards) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  if (table >= m->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (num_shards == 0 || shard >= num_shards) return fail(HPSX_ERR_INVALID_ARG, "shard must be < num_shards");
  m->tables[table]->fill_procedural(num_rows, seed, *ps->pool, shard, num_shards);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_table_rows(const hpsx_ps* ps, const char* model, size_t table, size_t* out) {
  Model* m = find_model(const_cast<hpsx_ps*>(ps), model);
  if (!m || !out) return fail(HPSX_ERR_NOT_FOUND, "unknown model");
  if (table >= m->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  *out = m->device_rows[table] != 0 ? static_cast<size_t>(m->device_rows[table]) : m->tables[table]->rows();
  return HPSX_OK;
}

int hpsx_ps_get_model_params(hpsx_ps* ps, const char* model, hpsx_model_params* out) {
  Model* m = find_model(ps, model);
  if (!m || !out) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  const ModelConfig& c = m->cfg;
  *out = hpsx_model_params{};
  out->model_name = c.model_name.c_str();
  out->max_batch_size = c.max_batch_size;
  out->num_tables = m->tables.size();
  out->sparse_files = m->c_sparse_files.data();
  out->table_names = m->c_table_names.data();
  out->embedding_vecsize_per_table = c.embedding_vecsize_per_table.data();
  out->maxnum_catfeature_query_per_table_per_sample = c.maxnum_catfeature_query_per_table_per_sample.data();
  out->default_value_for_each_table = c.default_value_for_each_table.data();
  out->use_gpu_embedding_cache = c.use_gpu_embedding_cache ? 1 : 0;
  out->hit_rate_threshold = c.hit_rate_threshold;
  out->cache_size_percentage = c.cache_size_percentage;
  out->number_of_worker_buffers_in_pool = c.number_of_worker_buffers_in_pool;
  out->deployed_devices = c.deployed_devices.data();
  out->num_deployed_devices = c.deployed_devices.size();
  out->embedding_cache_type = c.embedding_cache_type == CacheType::Static ? HPSX_CACHE_STATIC : HPSX_CACHE_DYNAMIC;
  out->cache_load_factor = m->load_factor;
  out->enable_pagelock = c.enable_pagelock ? 1 : 0;
  out->split_lock = m->split_lock ? 0 : -1;
  out->request_chunks = m->request_chunks;
  out->pull_grid_ctas = m->pull_grid_ctas;
  out->probe_variant = m->probe_variant;
  out->probe_variant_set = 1;
  out->peer_tier = m->peer_tier ? 1 : 0;
  return HPSX_OK;
}

int hpsx_ps_sync_models_from_json(hpsx_ps* ps, const char* ps_json_path, size_t* num_added) {
  HPSX_GUARD_BEGIN
  if (!ps || !ps_json_path) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (num_added) *num_added = 0;
  PsConfig cfg;
  const ParseResult pr = parse_ps_config_file(ps_json_path, &cfg);
  if (!pr.ok) return fail(HPSX_ERR_INVALID_ARG, pr.message);
  for (const ModelConfig& m : cfg.models) {
    if (find_model(ps, m.model_name.c_str()) != nullptr) continue;
    int rc = add_model_cfg(ps, m, 0.f);
    if (rc == HPSX_OK && m.use_gpu_embedding_cache && m.init_ec)
      rc = hpsx_ps_create_embedding_cache_per_model(ps, m.model_name.c_str());
    if (rc != HPSX_OK) return rc;
    if (num_added) ++*num_added;
  }
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_copy_to_host(int device, void* h_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return HPSX_OK;
  if (!h_dst || !d_src) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return HPSX_OK;
}

int hpsx_ps_lookup(hpsx_ps* ps, const char* model, size_t table, const int64_t* h_keys, size_t n,
                   float* h_vectors) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  if (table >= m->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (n > 0 && (!h_keys || !h_vectors)) return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer");
  const HostTable& ht = *m->tables[table];
  ht.fetch(h_keys, n, h_vectors, ht.dim(), *ps->pool);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_create_embedding_cache_per_model(hpsx_ps* ps, const char* model) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  if (!m->cfg.use_gpu_embedding_cache)
    return fail(HPSX_ERR_INVALID_ARG, "model '" + m->cfg.model_name + "' has gpucache = false");
  for (int dev : m->cfg.deployed_devices) {
    {
      std::lock_guard<std::mutex> lk(m->mu);
      if (m->caches.count(dev)) continue;
    }
    std::unique_ptr<hpsx_cache> c;
    const int rc = build_cache(ps, m, dev, &c);
    if (rc != HPSX_OK) return rc;
    std::lock_guard<std::mutex> lk(m->mu);
    m->caches[dev] = std::move(c);
  }
  if (m->peer_tier) {
    if (!m->direct_pull)
      return fail(HPSX_ERR_INVALID_ARG, "model '" + m->cfg.model_name + "': hpsx_peer_tier needs enable_pagelock");
    if (m->cfg.deployed_devices.size() >= 2 || m->tier_only()) return hpsx_ps_peer_tier_connect_local(ps, model);
  }
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_get_embedding_cache(hpsx_ps* ps, const char* model, int device, hpsx_cache** out) {
  Model* m = find_model(ps, model);
  if (!m || !out) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  std::lock_guard<std::mutex> lk(m->mu);
  auto it = m->caches.find(device);
  if (it == m->caches.end()) {
    *out = nullptr;
    return fail(HPSX_ERR_NOT_FOUND, "model '" + m->cfg.model_name + "' has no embedding cache on device " +
                                        std::to_string(device));
  }
  *out = it->second.get();
  return HPSX_OK;
}

int hpsx_ps_update_database_per_model(hpsx_ps* ps, const char* model) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  // Page-locked tables are read in place by the pull kernels: keep lookups out while rows are rewritten.
  std::vector<hpsx_cache*> caches;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    for (auto& kv : m->caches) caches.push_back(kv.second.get());
  }
  // lock order everywhere: async_mu (workspace owner), then pull_rw, then rw
  std::vector<std::unique_lock<std::mutex>> held_ws;
  std::vector<std::unique_lock<std::shared_mutex>> held;
  for (hpsx_cache* c : caches)
    if (c->direct_pull) held_ws.emplace_back(c->async_mu);
  // pull kernels of in-flight lookups read the page-locked rows and the HBM index (some of them outside c->rw:
  // the split-lock form, static caches of shard groups): wait for them, keep new ones out
  for (hpsx_cache* c : caches)
    if (c->direct_pull) held.emplace_back(c->pull_rw);
  for (hpsx_cache* c : caches)
    if (c->direct_pull) held.emplace_back(c->rw);
  for (size_t t = 0; t < m->tables.size() && t < m->cfg.sparse_files.size(); ++t) {
    if (m->cfg.sparse_files[t].empty() || m->device_rows[t] != 0) continue;
    const int rc = load_sparse_dir(ps, m->tables[t].get(), m->cfg.sparse_files[t]);
    if (rc != HPSX_OK) return rc;
  }
  for (hpsx_cache* c : caches) {
    if (!c->direct_pull) continue;
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
    const int rc = sync_direct_pull_index_impl(c, c->async_stream);
    if (rc != HPSX_OK) return rc;
    c->tier.committed = false;  // the index holds host addresses again
    c->tier.repointed = 0;
  }
  if (m->peer_tier) {
    std::vector<hpsx_cache*> dp;
    for (hpsx_cache* c : caches)
      if (c->direct_pull) dp.push_back(c);
    if (dp.size() >= 2 || m->tier_only()) return tier_connect_local_locked(m, dp);
  }
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_ps_refresh_embedding_cache(hpsx_ps* ps, const char* model, int device, size_t* refreshed_rows) {
  HPSX_GUARD_BEGIN
  if (refreshed_rows) *refreshed_rows = 0;
  hpsx_cache* c = nullptr;
  int rc = hpsx_ps_get_embedding_cache(ps, model, device, &c);
  if (rc != HPSX_OK) return rc;
  Model* m = c->model;
  DeviceGuard guard(c->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  std::lock_guard<std::mutex> ws(c->async_mu);  // one owner of the refresh / async-insert workspace
  uint32_t* d_updated = nullptr;
  HPSX_CU(cudaMalloc(&d_updated, sizeof(uint32_t)));
  HPSX_CU(cudaMemsetAsync(d_updated, 0, sizeof(uint32_t), c->async_stream));
  std::vector<int64_t> keys;
  for (size_t t = 0; t < c->tables.size() && rc == HPSX_OK; ++t) {
    keys.resize(c->slots[t]);
    size_t n = 0;
    rc = hpsx_cache_dump_keys(c, t, keys.data(), keys.size(), &n);
    if (rc != HPSX_OK) break;
    // cache_refresh_percentage_per_iteration (src/backend.cpp:411-416): share of the cache rewritten per
    // exclusive section, so that lookups interleave with a long refresh
    const double pct = m->cfg.cache_refresh_percentage_per_iteration;
    size_t chunk = pct > 0.0 ? static_cast<size_t>(pct * static_cast<double>(c->slots[t])) : c->async_rows;
    chunk = std::min(c->async_rows, std::max<size_t>(chunk, 1024));
    const HostTable& ht = *m->tables[t];
    const size_t dim = ht.dim();
    for (size_t off = 0; off < n; off += chunk) {
      const size_t mc = std::min(chunk, n - off);
      ht.fetch(keys.data() + off, mc, c->async_h_stage, dim, *ps->pool);
      std::unique_lock<std::shared_mutex> wl(c->rw);
      cudaError_t e = cudaMemcpyAsync(c->async_d_keys, keys.data() + off, mc * sizeof(int64_t), cudaMemcpyHostToDevice,
                                      c->async_stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->async_d_stage, c->async_h_stage, mc * dim * sizeof(float), cudaMemcpyHostToDevice,
                            c->async_stream);
      if (e == cudaSuccess) e = launch_update_values(c->tables[t], c->async_d_keys, c->async_d_stage, mc, d_updated, c->async_stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->async_stream);
      if (e != cudaSuccess) {
        rc = fail(HPSX_ERR_CUDA, std::string("cache refresh: ") + cudaGetErrorString(e));
        break;
      }
    }
  }
  uint32_t h = 0;
  if (rc == HPSX_OK && cudaMemcpy(&h, d_updated, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = fail(HPSX_ERR_CUDA, "cache refresh: reading the counter failed");
  cudaFree(d_updated);
  if (rc == HPSX_OK && refreshed_rows) *refreshed_rows = h;
  return rc;
  HPSX_GUARD_END
}

int hpsx_ps_destroy_embedding_cache_per_model(hpsx_ps* ps, const char* model) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  std::map<int, std::unique_ptr<hpsx_cache>> doomed;
  {
    std::lock_guard<std::mutex> lk(m->mu);
    doomed.swap(m->caches);
  }
  doomed.clear();
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_cache_num_tables(const hpsx_cache* cache, size_t* out) {
  if (!cache || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = cache->tables.size();
  return HPSX_OK;
}

int hpsx_cache_device(const hpsx_cache* cache, int* out) {
  if (!cache || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = cache->device;
  return HPSX_OK;
}

int hpsx_cache_capacity(const hpsx_cache* cache, size_t table, size_t* slots) {
  if (!cache || !slots) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (table >= cache->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  *slots = cache->slots[table];
  return HPSX_OK;
}

int hpsx_cache_resident(hpsx_cache* cache, size_t table, size_t* keys) {
  if (!cache || !keys) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (table >= cache->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  DeviceGuard guard(cache->device);
  std::shared_lock<std::shared_mutex> lk(cache->rw);
  unsigned long long* d = nullptr;
  HPSX_CU(cudaMalloc(&d, sizeof(unsigned long long)));
  cudaError_t e = launch_count_resident(cache->tables[table], d, nullptr);
  unsigned long long h = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  HPSX_CU(e);
  *keys = static_cast<size_t>(h);
  return HPSX_OK;
}

int hpsx_cache_dump_keys(hpsx_cache* cache, size_t table, int64_t* h_keys, size_t cap, size_t* n) {
  if (!cache || !n || (cap > 0 && !h_keys)) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (table >= cache->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  DeviceGuard guard(cache->device);
  std::shared_lock<std::shared_mutex> lk(cache->rw);
  unsigned long long* d_count = nullptr;
  int64_t* d_keys = nullptr;
  HPSX_CU(cudaMalloc(&d_count, sizeof(unsigned long long)));
  cudaError_t e = cudaMalloc(&d_keys, std::max<size_t>(cap, 1) * sizeof(int64_t));
  unsigned long long h = 0;
  if (e == cudaSuccess) e = launch_dump_keys(cache->tables[table], d_keys, cap, d_count, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost);
  const size_t got = std::min<size_t>(static_cast<size_t>(h), cap);
  if (e == cudaSuccess && got > 0)
    e = cudaMemcpy(h_keys, d_keys, got * sizeof(int64_t), cudaMemcpyDeviceToHost);
  cudaFree(d_count);
  if (d_keys) cudaFree(d_keys);
  HPSX_CU(e);
  *n = got;
  return HPSX_OK;
}

int hpsx_cache_drain_async(hpsx_cache* cache) {
  if (!cache) return fail(HPSX_ERR_INVALID_ARG, "null cache");
  std::unique_lock<std::mutex> lk(cache->async_mu);
  cache->async_cv.wait(lk, [cache] { return cache->async_pending == 0; });
  return HPSX_OK;
}

int hpsx_session_create(hpsx_ps* ps, const char* model, int device, hpsx_session** out) {
  HPSX_GUARD_BEGIN
  Model* m = find_model(ps, model);
  if (!m || !out) return fail(HPSX_ERR_NOT_FOUND, std::string("unknown model '") + (model ? model : "") + "'");
  std::unique_ptr<hpsx_session> s(new hpsx_session());
  s->ps = ps;
  s->model = m;
  const size_t T = m->tables.size();
  s->cap_per_table.resize(T);
  for (size_t t = 0; t < T; ++t) {
    s->cap_per_table[t] =
        m->cfg.max_batch_size * m->cfg.maxnum_catfeature_query_per_table_per_sample[t];
    s->cap_keys += s->cap_per_table[t];
    s->max_dim = std::max(s->max_dim, m->tables[t]->dim());
  }
  if (!m->cfg.use_gpu_embedding_cache || device < 0) {
    // CPU session: vectors are returned in host memory (hps_backend/src/hps.cc:640-642)
    *out = s.release();
    return HPSX_OK;
  }
  hpsx_cache* c = nullptr;
  int rc = hpsx_ps_get_embedding_cache(ps, model, device, &c);
  if (rc != HPSX_OK) return rc;
  s->cache = c;
  c->sessions.fetch_add(1, std::memory_order_relaxed);
  s->device = device;
  s->probe_variant = m->probe_variant;
  s->request_chunks = m->request_chunks;
  s->pull_grid_ctas = m->pull_grid_ctas;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  const size_t cap = std::max<size_t>(s->cap_keys, 1);
  HPSX_CU(cudaMalloc(&s->d_keys, cap * sizeof(int64_t)));
  HPSX_CU(cudaMalloc(&s->d_miss_pos, cap * sizeof(uint32_t)));
  HPSX_CU(cudaMalloc(&s->d_miss_keys, cap * sizeof(int64_t)));
  // counters and events are indexed by "virtual table" v = request * T + table, so that one call can serve a
  // batch of up to kMaxBatchRequests requests (hpsx_session_lookup_batch); a plain lookup uses v = table
  s->vt = T * kMaxBatchRequests;
  HPSX_CU(cudaMalloc(&s->d_counters, 3 * s->vt * sizeof(uint32_t)));
  HPSX_CU(cudaMallocHost(&s->h_counters, 3 * s->vt * sizeof(uint32_t)));
  HPSX_CU(cudaMemsetAsync(s->d_counters, 0, 3 * s->vt * sizeof(uint32_t), s->stream));
  // miss keys are written by the probe kernels straight into this mapped buffer (zero-copy)
  HPSX_CU(cudaHostAlloc(&s->h_miss_keys, cap * sizeof(int64_t),
                        cudaHostAllocMapped | cudaHostAllocPortable));
  HPSX_CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->hd_miss_keys), s->h_miss_keys, 0));
  const size_t stage_rows = std::min<size_t>(kStageChunkRows, std::max<size_t>(cap, kMinStageChunkRows));
  for (int b = 0; b < kNumStages; ++b) {
    HPSX_CU(cudaMallocHost(&s->h_stage[b], stage_rows * s->max_dim * sizeof(float)));
    HPSX_CU(cudaMalloc(&s->d_stage[b], stage_rows * s->max_dim * sizeof(float)));
    HPSX_CU(cudaEventCreateWithFlags(&s->stage_free[b], cudaEventDisableTiming));
  }
  HPSX_CU(cudaStreamSynchronize(s->stream));
  s->ev.resize(2 * s->vt);
  for (auto& e : s->ev) HPSX_CU(cudaEventCreate(&e));
  s->ev_pull.resize(2 * s->vt);
  for (auto& e : s->ev_pull) HPSX_CU(cudaEventCreate(&e));
  *out = s.release();
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_session_destroy(hpsx_session* s) {
  HPSX_GUARD_BEGIN
  delete s;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_session_device(const hpsx_session* s, int* out) {
  if (!s || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = s->device;
  return HPSX_OK;
}

int hpsx_session_stream(const hpsx_session* s, void** out) {
  if (!s || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *out = s->stream;
  return HPSX_OK;
}

int hpsx_session_lookup(hpsx_session* s, const void* const* h_keys_per_table,
                        float* const* vectors_per_table, const size_t* num_keys_per_table,
                        size_t num_tables) {
  HPSX_GUARD_BEGIN
  int rc = check_tables(s, num_keys_per_table, num_tables);
  if (rc != HPSX_OK) return rc;
  if (num_tables > 0 && (!h_keys_per_table || !vectors_per_table))
    return fail(HPSX_ERR_INVALID_ARG, "null pointer arrays");
  std::lock_guard<std::mutex> lk(s->mu);
  if (!s->cache) {
    ++s->stats.lookups;
    for (size_t t = 0; t < num_tables; ++t) {
      const size_t n = num_keys_per_table[t];
      if (n == 0) continue;
      if (!h_keys_per_table[t] || !vectors_per_table[t])
        return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer for table " + std::to_string(t));
      const HostTable& ht = *s->model->tables[t];
      const double t0 = now_ms();
      const size_t absent = ht.fetch(static_cast<const int64_t*>(h_keys_per_table[t]), n,
                                     vectors_per_table[t], ht.dim(), *s->ps->pool);
      s->stats.host_gather_ms += now_ms() - t0;
      s->stats.keys += n;
      s->stats.misses += n;
      s->stats.default_filled += absent;
    }
    return HPSX_OK;
  }
  return gpu_lookup(s, h_keys_per_table, false, vectors_per_table, num_keys_per_table, num_tables);
  HPSX_GUARD_END
}

int hpsx_session_lookup_device_keys(hpsx_session* s, const int64_t* const* d_keys_per_table,
                                    float* const* d_vectors_per_table,
                                    const size_t* num_keys_per_table, size_t num_tables) {
  HPSX_GUARD_BEGIN
  int rc = check_tables(s, num_keys_per_table, num_tables);
  if (rc != HPSX_OK) return rc;
  if (!s->cache)
    return fail(HPSX_ERR_UNSUPPORTED, "device keys need a GPU session (gpucache = true)");
  if (num_tables > 0 && (!d_keys_per_table || !d_vectors_per_table))
    return fail(HPSX_ERR_INVALID_ARG, "null pointer arrays");
  std::lock_guard<std::mutex> lk(s->mu);
  return gpu_lookup(s, reinterpret_cast<const void* const*>(d_keys_per_table), true,
                    d_vectors_per_table, num_keys_per_table, num_tables);
  HPSX_GUARD_END
}

int hpsx_session_lookup_scatter(hpsx_session* s, size_t table, const int64_t* d_keys, const uint32_t* d_pos,
                                size_t n, float* d_out_base) {
  HPSX_GUARD_BEGIN
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (!s->cache) return fail(HPSX_ERR_UNSUPPORTED, "scatter lookup needs a GPU session (gpucache = true)");
  const size_t T = s->model->tables.size();
  if (table >= T) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (n > s->cap_per_table[table])
    return fail(HPSX_ERR_INVALID_ARG, "scatter lookup exceeds the table's key capacity (max_batch_size * maxnum_catfeature)");
  if (n == 0) return HPSX_OK;
  if (!d_keys || !d_pos || !d_out_base) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  std::vector<const void*> keys(T, nullptr);
  std::vector<float*> out(T, nullptr);
  std::vector<size_t> cnt(T, 0);
  std::vector<const uint32_t*> pos(T, nullptr);
  keys[table] = d_keys;
  out[table] = d_out_base;
  cnt[table] = n;
  pos[table] = d_pos;
  std::lock_guard<std::mutex> lk(s->mu);
  return gpu_lookup(s, keys.data(), true, out.data(), cnt.data(), T, pos.data());
  HPSX_GUARD_END
}

// ------------------------------------------------------------------------------------------------
// device buffers shareable between the processes of one box (one process per GPU): CUDA IPC
// ------------------------------------------------------------------------------------------------
int hpsx_device_malloc(int device, size_t bytes, void** d_ptr) {
  if (!d_ptr) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(cudaMalloc(d_ptr, std::max<size_t>(bytes, 16)));
  return HPSX_OK;
}

int hpsx_device_free(int device, void* d_ptr) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(cudaFree(d_ptr));
  return HPSX_OK;
}

int hpsx_ipc_export(int device, void* d_ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!d_ptr || !handle64) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t h;
  HPSX_CU(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle64, &h, sizeof(h));
  return HPSX_OK;
}

int hpsx_ipc_open(int device, const void* handle64, void** d_ptr) {
  if (!handle64 || !d_ptr) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  HPSX_CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HPSX_OK;
}

int hpsx_ipc_close(int device, void* d_ptr) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(cudaIpcCloseMemHandle(d_ptr));
  return HPSX_OK;
}

int hpsx_session_lookup_ex(hpsx_session* s, const void* const* keys_per_table, int key_memory,
                           float* const* vectors_per_table, int vector_memory,
                           const size_t* num_keys_per_table, size_t num_tables) {
  HPSX_GUARD_BEGIN
  const bool keys_dev = key_memory == HPSX_MEM_DEVICE, vec_dev = vector_memory == HPSX_MEM_DEVICE;
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (!s->cache) {
    if (keys_dev || vec_dev)
      return fail(HPSX_ERR_UNSUPPORTED, "a CPU session (gpucache = false) takes host keys and host vectors");
    return hpsx_session_lookup(s, keys_per_table, vectors_per_table, num_keys_per_table, num_tables);
  }
  if (vec_dev) {
    return keys_dev ? hpsx_session_lookup_device_keys(s, reinterpret_cast<const int64_t* const*>(keys_per_table),
                                                      vectors_per_table, num_keys_per_table, num_tables)
                    : hpsx_session_lookup(s, keys_per_table, vectors_per_table, num_keys_per_table, num_tables);
  }
  // GPU session, host vectors: gather into the session's device result buffer, then D2H per table.
  int rc = check_tables(s, num_keys_per_table, num_tables);
  if (rc != HPSX_OK) return rc;
  if (num_tables > 0 && (!keys_per_table || !vectors_per_table))
    return fail(HPSX_ERR_INVALID_ARG, "null pointer arrays");
  std::lock_guard<std::mutex> lk(s->mu);
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  if (!s->d_result) {
    size_t floats = 0;
    for (size_t t = 0; t < s->cap_per_table.size(); ++t) floats += s->cap_per_table[t] * s->model->tables[t]->dim();
    HPSX_CU(cudaMalloc(&s->d_result, std::max<size_t>(floats, 1) * sizeof(float)));
  }
  std::vector<float*> d_out(num_tables, nullptr);
  size_t off = 0;
  for (size_t t = 0; t < num_tables; ++t) {
    d_out[t] = s->d_result + off;
    off += num_keys_per_table[t] * s->model->tables[t]->dim();
  }
  for (size_t t = 0; t < num_tables; ++t)
    if (num_keys_per_table[t] != 0 && !vectors_per_table[t])
      return fail(HPSX_ERR_INVALID_ARG, "null vector pointer for table " + std::to_string(t));
  s->host_out = vectors_per_table;
  s->host_out_done = false;
  rc = gpu_lookup(s, keys_per_table, keys_dev, d_out.data(), num_keys_per_table, num_tables);
  s->host_out = nullptr;
  if (rc != HPSX_OK) return rc;
  if (s->host_out_done) return HPSX_OK;  // the binned pipeline copied every chunk as it completed
  for (size_t t = 0; t < num_tables; ++t) {
    const size_t bytes = num_keys_per_table[t] * s->model->tables[t]->dim() * sizeof(float);
    if (bytes == 0) continue;
    if (!vectors_per_table[t]) return fail(HPSX_ERR_INVALID_ARG, "null vector pointer for table " + std::to_string(t));
    HPSX_CU(cudaMemcpyAsync(vectors_per_table[t], d_out[t], bytes, cudaMemcpyDeviceToHost, s->stream));
    s->stats.d2h_bytes += bytes;
  }
  HPSX_CU(cudaStreamSynchronize(s->stream));
  return HPSX_OK;
  HPSX_GUARD_END
}

static_assert(kMaxBatchRequests == HPSX_MAX_BATCH_REQUESTS, "header and engine disagree on the batch limit");

int hpsx_session_lookup_batch(hpsx_session* s, size_t num_requests, const void* const* keys, int key_memory,
                              float* const* vectors, int vector_memory, const size_t* num_keys) {
  HPSX_GUARD_BEGIN
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (num_requests == 0) return HPSX_OK;
  if (!keys || !vectors || !num_keys) return fail(HPSX_ERR_INVALID_ARG, "null pointer arrays");
  const size_t T = s->model->tables.size();
  if (num_requests > kMaxBatchRequests)
    return fail(HPSX_ERR_INVALID_ARG, "a batch holds at most " + std::to_string(kMaxBatchRequests) + " requests");
  size_t total = 0;
  for (size_t r = 0; r < num_requests; ++r) {
    const int rc = check_tables(s, num_keys + r * T, T);
    if (rc != HPSX_OK) return rc;
    for (size_t t = 0; t < T; ++t) total += num_keys[r * T + t];
  }
  const bool fused = s->cache != nullptr && vector_memory == HPSX_MEM_DEVICE && total <= s->cap_keys;
  if (!fused) {
    // CPU sessions, host output buffers and batches larger than the workspace: one request at a time
    for (size_t r = 0; r < num_requests; ++r) {
      const int rc = hpsx_session_lookup_ex(s, keys + r * T, key_memory, vectors + r * T, vector_memory, num_keys + r * T, T);
      if (rc != HPSX_OK) return rc;
    }
    return HPSX_OK;
  }
  std::lock_guard<std::mutex> lk(s->mu);
  return gpu_lookup(s, keys, key_memory == HPSX_MEM_DEVICE, vectors, num_keys, num_requests * T);
  HPSX_GUARD_END
}

int hpsx_session_lookup_bf16_mirror(hpsx_session* s, size_t table, const int64_t* keys, int key_memory, size_t n,
                                    float* d_vectors, void* d_vectors_bf16) {
  HPSX_GUARD_BEGIN
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (!s->cache) return fail(HPSX_ERR_UNSUPPORTED, "the bf16 mirror needs a GPU session (gpucache = true)");
  const size_t T = s->model->tables.size();
  if (table >= T) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (n > s->cap_per_table[table])
    return fail(HPSX_ERR_INVALID_ARG, "lookup exceeds the table's key capacity (max_batch_size * maxnum_catfeature)");
  if (n == 0) return HPSX_OK;
  if (!keys || !d_vectors || !d_vectors_bf16) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (s->model->tables[table]->dim() % 8 != 0)
    return fail(HPSX_ERR_UNSUPPORTED, "the bf16 mirror needs rows that are multiples of 8 floats");
  // synchronous insertion only: a row that arrives after the response has no defined place in the mirror
  std::vector<const void*> kp(T, nullptr);
  std::vector<float*> op(T, nullptr);
  std::vector<size_t> np(T, 0);
  kp[table] = keys;
  op[table] = d_vectors;
  np[table] = n;
  std::lock_guard<std::mutex> lk(s->mu);
  const int saved_mode = s->insert_mode;
  s->insert_mode = 1;
  s->bf16_table = table;
  s->bf16_out = d_vectors_bf16;
  const int rc = gpu_lookup(s, kp.data(), key_memory == HPSX_MEM_DEVICE, op.data(), np.data(), T);
  s->bf16_out = nullptr;
  s->insert_mode = saved_mode;
  if (rc == HPSX_ERR_CUDA && std::string(g_err).find("not supported") != std::string::npos)
    return fail(HPSX_ERR_UNSUPPORTED, "the bf16 mirror needs 32-byte aligned fp32 output and 16-byte aligned bf16 output");
  return rc;
  HPSX_GUARD_END
}

static int pooled_common(hpsx_session* s, size_t table, const int64_t* keys, bool on_device,
                         size_t num_bags, size_t hotness, int combiner, float* d_pooled) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (!s->cache)
    return fail(HPSX_ERR_UNSUPPORTED, "pooled lookup needs a GPU session (gpucache = true)");
  if (table >= s->model->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (combiner != HPSX_COMBINER_SUM && combiner != HPSX_COMBINER_MEAN)
    return fail(HPSX_ERR_INVALID_ARG, "unknown combiner");
  if (hotness == 0 && num_bags > 0) return fail(HPSX_ERR_INVALID_ARG, "hotness must be > 0");
  if (num_bags * hotness > s->cap_per_table[table])
    return fail(HPSX_ERR_INVALID_ARG, "pooled lookup exceeds the table's key capacity");
  std::lock_guard<std::mutex> lk(s->mu);
  return gpu_lookup_pooled(s, table, keys, on_device, num_bags, hotness, combiner, d_pooled);
}

int hpsx_session_lookup_pooled(hpsx_session* s, size_t table, const int64_t* h_keys, size_t num_bags,
                               size_t hotness, int combiner, float* d_pooled) {
  HPSX_GUARD_BEGIN
  return pooled_common(s, table, h_keys, false, num_bags, hotness, combiner, d_pooled);
  HPSX_GUARD_END
}

int hpsx_session_lookup_pooled_device_keys(hpsx_session* s, size_t table, const int64_t* d_keys,
                                           size_t num_bags, size_t hotness, int combiner,
                                           float* d_pooled) {
  HPSX_GUARD_BEGIN
  return pooled_common(s, table, d_keys, true, num_bags, hotness, combiner, d_pooled);
  HPSX_GUARD_END
}

// CPU session: rows fetched from the host table chunk by chunk and reduced in ascending slot order in
// fp32 — the same order as the oracle and the pooled kernel.
static int cpu_lookup_pooled(hpsx_session* s, size_t table, const int64_t* keys, size_t num_bags, size_t hotness,
                             int combiner, float* pooled) {
  const HostTable& ht = *s->model->tables[table];
  const size_t dim = ht.dim();
  constexpr size_t kChunkBags = 4096;
  std::vector<float> rows(std::min(num_bags, kChunkBags) * hotness * dim);
  ++s->stats.lookups;
  s->stats.keys += num_bags * hotness;
  s->stats.misses += num_bags * hotness;
  for (size_t b0 = 0; b0 < num_bags; b0 += kChunkBags) {
    const size_t nb = std::min(kChunkBags, num_bags - b0);
    s->stats.default_filled += ht.fetch(keys + b0 * hotness, nb * hotness, rows.data(), dim, *s->ps->pool);
    s->ps->pool->parallel_for(nb, [&](size_t b) {
      float* acc = pooled + (b0 + b) * dim;
      const float* r = rows.data() + b * hotness * dim;
      for (size_t d = 0; d < dim; ++d) acc[d] = 0.f;
      for (size_t j = 0; j < hotness; ++j)
        for (size_t d = 0; d < dim; ++d) acc[d] += r[j * dim + d];
      if (combiner == HPSX_COMBINER_MEAN)
        for (size_t d = 0; d < dim; ++d) acc[d] /= static_cast<float>(hotness);
    });
  }
  return HPSX_OK;
}

int hpsx_session_lookup_pooled_ex(hpsx_session* s, size_t table, const int64_t* keys, int key_memory,
                                  size_t num_bags, size_t hotness, int combiner, float* pooled,
                                  int pooled_memory) {
  HPSX_GUARD_BEGIN
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  const bool keys_dev = key_memory == HPSX_MEM_DEVICE, out_dev = pooled_memory == HPSX_MEM_DEVICE;
  if (table >= s->model->tables.size()) return fail(HPSX_ERR_NOT_FOUND, "table index out of range");
  if (combiner != HPSX_COMBINER_SUM && combiner != HPSX_COMBINER_MEAN)
    return fail(HPSX_ERR_INVALID_ARG, "unknown combiner");
  if (hotness == 0 && num_bags > 0) return fail(HPSX_ERR_INVALID_ARG, "hotness must be > 0");
  if (num_bags * hotness > s->cap_per_table[table])
    return fail(HPSX_ERR_INVALID_ARG, "pooled lookup exceeds the table's key capacity");
  if (num_bags == 0) return HPSX_OK;
  if (!keys || !pooled) return fail(HPSX_ERR_INVALID_ARG, "null key/vector pointer");
  if (!s->cache) {
    if (keys_dev || out_dev)
      return fail(HPSX_ERR_UNSUPPORTED, "a CPU session (gpucache = false) takes host keys and host vectors");
    std::lock_guard<std::mutex> lk(s->mu);
    return cpu_lookup_pooled(s, table, keys, num_bags, hotness, combiner, pooled);
  }
  if (out_dev) return pooled_common(s, table, keys, keys_dev, num_bags, hotness, combiner, pooled);
  // GPU session, host output: pool into the device result buffer, then D2H
  std::lock_guard<std::mutex> lk(s->mu);
  DeviceGuard guard(s->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  if (!s->d_result) {
    size_t floats = 0;
    for (size_t t = 0; t < s->cap_per_table.size(); ++t) floats += s->cap_per_table[t] * s->model->tables[t]->dim();
    HPSX_CU(cudaMalloc(&s->d_result, std::max<size_t>(floats, 1) * sizeof(float)));
  }
  const int rc = gpu_lookup_pooled(s, table, keys, keys_dev, num_bags, hotness, combiner, s->d_result);
  if (rc != HPSX_OK) return rc;
  const size_t bytes = num_bags * s->model->tables[table]->dim() * sizeof(float);
  HPSX_CU(cudaMemcpyAsync(pooled, s->d_result, bytes, cudaMemcpyDeviceToHost, s->stream));
  HPSX_CU(cudaStreamSynchronize(s->stream));
  s->stats.d2h_bytes += bytes;
  return HPSX_OK;
  HPSX_GUARD_END
}

// rows inserted into the cache are counted on the device (cumulative since the last reset)
static int read_inserted(hpsx_session* s, uint64_t* out) {
  *out = 0;
  if (!s->cache) return HPSX_OK;
  DeviceGuard guard(s->device);
  std::vector<uint32_t> h(s->vt, 0);
  HPSX_CU(cudaMemcpyAsync(h.data(), s->d_counters + s->vt, s->vt * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          s->stream));
  HPSX_CU(cudaStreamSynchronize(s->stream));
  for (uint32_t v : h) *out += v;
  return HPSX_OK;
}

int hpsx_session_get_stats(const hpsx_session* cs, hpsx_session_stats* out) {
  if (!cs || !out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  hpsx_session* s = const_cast<hpsx_session*>(cs);
  std::lock_guard<std::mutex> lk(s->mu);
  *out = s->stats;
  return read_inserted(s, &out->inserted);
}

int hpsx_session_reset_stats(hpsx_session* s) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  std::lock_guard<std::mutex> lk(s->mu);
  s->stats = hpsx_session_stats{};
  if (s->cache) {
    DeviceGuard guard(s->device);
    HPSX_CU(cudaMemsetAsync(s->d_counters + s->vt, 0, s->vt * sizeof(uint32_t), s->stream));
    HPSX_CU(cudaStreamSynchronize(s->stream));
  }
  return HPSX_OK;
}

int hpsx_session_set_insert_mode(hpsx_session* s, int mode) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  s->insert_mode = mode < 0 ? -1 : (mode > 0 ? 1 : 0);
  return HPSX_OK;
}

int hpsx_session_set_debug(hpsx_session* s, int flags) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  s->debug_flags = flags;
  return HPSX_OK;
}

int hpsx_session_set_probe_variant(hpsx_session* s, int variant) {
  if (!s) return fail(HPSX_ERR_INVALID_ARG, "null session");
  if (variant != kProbeLdg && variant != kProbeTma && variant != kProbeV8)
    return fail(HPSX_ERR_INVALID_ARG, "unknown probe variant");
  s->probe_variant = variant;
  return HPSX_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone device primitives
// ------------------------------------------------------------------------------------------------
int hpsx_unique(int device, const int64_t* d_keys, size_t n, int64_t* d_unique, uint32_t* d_inverse,
                size_t* h_num_unique, void* stream) {
  HPSX_GUARD_BEGIN
  if (!h_num_unique) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  *h_num_unique = 0;
  if (n == 0) return HPSX_OK;
  if (!d_keys || !d_unique || !d_inverse) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (n > (1ull << 31)) return fail(HPSX_ERR_UNSUPPORTED, "too many keys for one dedup call");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  int64_t* ws_keys = nullptr;
  uint32_t *ws_ids = nullptr, *d_counter = nullptr;
  HPSX_CU(cudaMalloc(&ws_keys, cap * sizeof(int64_t)));
  cudaError_t e = cudaMalloc(&ws_ids, cap * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_counter, 4 * sizeof(uint32_t));
  if (e == cudaSuccess)
    e = launch_unique(d_keys, n, ws_keys, ws_ids, cap, d_unique, d_inverse, d_counter, st);
  uint32_t h[4] = {0, 0, 0, 0};
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h, d_counter, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(ws_keys);
  if (ws_ids) cudaFree(ws_ids);
  if (d_counter) cudaFree(d_counter);
  HPSX_CU(e);
  *h_num_unique = h[0];
  return HPSX_OK;
  HPSX_GUARD_END
}

uint32_t hpsx_owner(int64_t key, uint32_t num_shards) { return owner_of(key, num_shards); }

int hpsx_owner_batch(const int64_t* h_keys, size_t n, uint32_t num_shards, uint32_t* h_owners) {
  if (n > 0 && (!h_keys || !h_owners)) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (num_shards == 0) return fail(HPSX_ERR_INVALID_ARG, "num_shards must be > 0");
  for (size_t i = 0; i < n; ++i) h_owners[i] = owner_of(h_keys[i], num_shards);
  return HPSX_OK;
}

int hpsx_route_keys(int device, const int64_t* d_keys, size_t n, uint32_t num_shards,
                    int64_t* d_routed_keys, uint32_t* d_perm, uint32_t* d_counts, uint32_t* h_counts,
                    void* stream) {
  HPSX_GUARD_BEGIN
  if (num_shards == 0 || num_shards > 64) return fail(HPSX_ERR_INVALID_ARG, "num_shards must be in [1,64]");
  if (!d_counts) return fail(HPSX_ERR_INVALID_ARG, "null d_counts");
  if (n > 0 && (!d_keys || !d_routed_keys || !d_perm)) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  if (n > 0xFFFFFFFFull) return fail(HPSX_ERR_UNSUPPORTED, "too many keys for one routing call");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* d_cursor = nullptr;
  HPSX_CU(cudaMalloc(&d_cursor, num_shards * sizeof(uint32_t)));
  cudaError_t e = launch_route_keys(d_keys, n, num_shards, d_routed_keys, d_perm, d_counts, d_cursor, st);
  if (e == cudaSuccess && h_counts)
    e = cudaMemcpyAsync(h_counts, d_counts, num_shards * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_cursor);
  HPSX_CU(e);
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_scatter_rows(int device, const float* d_rows, const uint32_t* d_perm, size_t n, size_t d,
                      float* d_out, void* stream) {
  HPSX_GUARD_BEGIN
  if (n == 0) return HPSX_OK;
  if (!d_rows || !d_perm || !d_out || d == 0) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(launch_scatter_rows(d_rows, d_perm, n, d, d_out, static_cast<cudaStream_t>(stream)));
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_gather_rows(int device, const float* d_table, const uint32_t* d_idx, size_t n, size_t dim,
                     float* d_out, void* stream) {
  HPSX_GUARD_BEGIN
  if (n == 0) return HPSX_OK;
  if (!d_table || !d_idx || !d_out) return fail(HPSX_ERR_INVALID_ARG, "null argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  HPSX_CU(launch_gather_rows(d_table, d_idx, n, dim, d_out, static_cast<cudaStream_t>(stream)));
  return HPSX_OK;
  HPSX_GUARD_END
}

}  // extern "C"
