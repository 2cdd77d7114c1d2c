// Internal C++ objects behind the opaque handles of include/hpsx.h.
//   hpsx_ps      ~ HugeCTR::HierParameterServerBase   (reference: hps_backend/src/backend.cpp:68-71)
//   hpsx_cache   ~ HugeCTR::EmbeddingCacheBase        (hps_backend/src/model_state.cpp:404-412)
//   hpsx_session ~ HugeCTR::LookupSessionBase         (hps_backend/src/model_instance_state.cpp:170-171)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/hpsx.h"
#include "host_ps.hpp"
#include "kernels.h"
#include "ps_config.hpp"

namespace hpsx {

constexpr size_t kStageChunkRows = 32768;    // most rows per host->device miss chunk (16 MiB at dim 128)
constexpr size_t kMinStageChunkRows = 2048;  // smaller chunks cost more in launch + pool wake-ups than they hide
constexpr int kNumStages = 3;                // pinned/device stage pairs rotating through the miss pipeline
constexpr size_t kPipelineMinKeys = 1u << 18; // smaller requests gain nothing from the chunked direct pull
constexpr size_t kMaxBatchRequests = 16;     // requests one hpsx_session_lookup_batch call may serve

struct Model;

}  // namespace hpsx

namespace hpsx {
// NVLink tier of one cache (kernels.h launch_tier_fill / launch_index_repoint; include/hpsx.h hpsx_cache_peer_tier_*).
struct PeerTier {
  struct Shard {
    unsigned char* base = nullptr;  // [int64 keys[cap] | pad to 512 B | float rows[cap][dim]], one cudaMalloc
    uint64_t cap = 0, rows = 0;
    bool ipc = false;               // mapped with cudaIpcOpenMemHandle (closed, not freed, on release)
    static size_t rows_offset(uint64_t cap) { return (static_cast<size_t>(cap) * sizeof(int64_t) + 511) & ~static_cast<size_t>(511); }
    static size_t bytes(uint64_t cap, size_t dim) { return rows_offset(cap) + static_cast<size_t>(cap) * dim * sizeof(float); }
  };
  uint32_t rank = 0, world = 0;            // world == 0: no tier
  std::vector<Shard> own;                  // [T]
  std::vector<std::vector<Shard>> peers;   // [T][world]; entry [t][rank] aliases own[t]
  bool committed = false;                  // the direct-pull index points at the shards
  uint64_t repointed = 0;                  // index entries that point into a shard (all tables)
};
}  // namespace hpsx

// One HBM embedding cache: all tables of one model on one device.
struct hpsx_cache {
  hpsx::Model* model = nullptr;
  int device = -1;
  bool is_static = false;
  std::vector<hpsx::DeviceTable> tables;
  std::vector<size_t> slots;            // capacity of every table (ways * buckets)
  bool direct_pull = false;             // enable_pagelock: misses are pulled by kernels from pinned host rows
  std::vector<hpsx::IndexSlot*> indexes;  // HBM mirrors of the host tables' key -> row-address index
  std::atomic<uint32_t> epoch{1};       // one tick per lookup call; LRU stamps are epochs
  std::atomic<int> sessions{0};         // lookup sessions (model instances) attached to this cache
  // Probes (readers) run concurrently; a kernel that rewrites slots (insert) excludes them, so a
  // row is never copied while it is being replaced.
  std::shared_mutex rw;
  // Direct pull: kernels read the page-locked host rows and their HBM index.  Lookups hold this shared for the whole
  // call (also where they run outside `rw`), a database reload holds it exclusively while it rewrites rows,
  // re-registers slabs and rebuilds the index.  Lock order: async_mu, pull_rw, rw.
  std::shared_mutex pull_rw;
  hpsx::PeerTier tier;                  // guarded by pull_rw (exclusive to change, shared while kernels read through it)
  // asynchronous insertion (hit_rate >= hit_rate_threshold): one workspace, jobs serialised
  std::mutex async_mu;
  std::condition_variable async_cv;
  size_t async_pending = 0;
  cudaStream_t async_stream = nullptr;
  int64_t* async_d_keys = nullptr;
  float* async_d_stage = nullptr;
  float* async_h_stage = nullptr;
  size_t async_rows = 0;  // capacity of the async staging buffers (rows of the widest table)
  size_t max_dim = 0;
  std::atomic<uint64_t> async_inserted{0};

  ~hpsx_cache();
};

namespace hpsx {

struct Model {
  ModelConfig cfg;
  float load_factor = 0.5f;
  bool direct_pull = false;  // cfg.enable_pagelock
  // engine extensions of ps.json / hpsx_model_params (ps_config.hpp: hpsx_*)
  bool split_lock = true;
  int request_chunks = 4;
  int pull_grid_ctas = 148;
  int probe_variant = kProbeV8;
  bool peer_tier = false;
  // sparse_files entry "synthetic_device:rows=<N>,seed=<S>": table t has NO host rows; its rows (keys [0, N), synth_value)
  // are generated on the devices, sharded by owner_of over the NVLink tier, which is then the only copy
  std::vector<unsigned long long> device_rows, device_seed;  // [T], rows == 0: an ordinary table
  bool tier_only() const {
    for (unsigned long long r : device_rows)
      if (r == 0) return false;
    return !device_rows.empty();
  }
  // C views of cfg for hpsx_ps_get_model_params (built once in add_model_cfg)
  std::vector<const char*> c_sparse_files, c_table_names;
  std::vector<std::string> table_names;
  std::vector<std::unique_ptr<HostTable>> tables;
  std::mutex mu;  // guards `caches`
  std::map<int, std::unique_ptr<hpsx_cache>> caches;
};

}  // namespace hpsx

struct hpsx_ps {
  hpsx::VolatileDbConfig vdb;
  std::unique_ptr<hpsx::ThreadPool> pool;
  std::mutex mu;  // guards `models`
  std::map<std::string, std::unique_ptr<hpsx::Model>> models;
  std::vector<std::string> model_order;
};

struct hpsx_session {
  hpsx_ps* ps = nullptr;
  hpsx::Model* model = nullptr;
  hpsx_cache* cache = nullptr;  // nullptr: CPU session
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_b = nullptr;     // binned direct pull: misses of chunk c are pulled while chunk c+1 is probed
  std::vector<cudaEvent_t> ev_chunk;   // untimed events of the binned pipeline: keys copied / probed / pulled, per group
  // bf16 mirror of the current lookup's output (hpsx_session_lookup_bf16_mirror): table index and device buffer
  size_t bf16_table = 0;
  void* bf16_out = nullptr;
  double miss_ratio = 1.0;             // running miss ratio of this session's lookups (predicts the next miss count)
  cudaStream_t stream_c = nullptr;     // key copies of the chunks of a request (copy engine)
  cudaStream_t stream_d = nullptr;     // host-output requests: D2H of a chunk's rows while later chunks are served
  // set by hpsx_session_lookup_ex around a lookup whose vectors go to HOST memory: per-table host destinations; the
  // binned pipeline copies every chunk as soon as it is complete and sets host_out_done
  float* const* host_out = nullptr;
  bool host_out_done = false;
  int request_chunks = 4;              // a request of >= kPipelineMinKeys keys is cut into this many chunks: the pull of
                                       // chunk c (and the key copy of chunk c+1) overlaps the probe of chunk c+1
  int pull_grid_ctas = 148;            // CTAs of the persistent binned pull kernel.  One per SM: on a 10 M-row table 148-444
                                       // CTAs all give 48-49 GB/s (518: 42), on a 100 M-row table 148 give 43 GB/s, 222
                                       // 35 and 370 32 — the more reads are in flight, the wider they spread over host memory
  int debug_flags = 0;                 // hpsx_session_set_debug: 1 inserts after ALL pulls, 2 pulls after ALL probes, 4 timeline
  std::vector<cudaEvent_t> ev_trace;   // timing events of the timeline
  // binned miss lists (MissBins) of the groups of one request
  uint32_t* d_bin_count = nullptr;
  uint32_t* h_bin_count = nullptr;     // pinned mirror
  size_t bin_count_cap = 0;            // words
  int64_t* d_bin_keys = nullptr;
  uint32_t* d_bin_pos = nullptr;
  size_t bin_entry_cap = 0;
  int probe_variant = hpsx::kProbeV8;  // falls back to the LDG.128 variant for rows that are not 32-B multiples
  int insert_mode = -1;

  size_t vt = 0;                       // virtual tables (request x table) the counters / events are sized for
  size_t cap_keys = 0;                 // sum over tables of max_batch * maxnum_catfeature
  std::vector<size_t> cap_per_table;
  int64_t* d_keys = nullptr;           // [cap_keys]
  uint32_t* d_miss_pos = nullptr;      // [cap_keys]
  int64_t* d_miss_keys = nullptr;      // [cap_keys]
  uint32_t* d_counters = nullptr;      // [0,T) miss counts, [T,2T) inserted (cumulative), [2T,3T) absent keys
  uint32_t* h_counters = nullptr;      // pinned mirror
  int64_t* h_miss_keys = nullptr;      // mapped pinned [cap_keys]: the probe kernels mirror miss keys here
  int64_t* hd_miss_keys = nullptr;     // device-visible address of h_miss_keys
  float* h_stage[hpsx::kNumStages] = {};  // pinned [<= kStageChunkRows * max_dim]
  float* d_stage[hpsx::kNumStages] = {};
  cudaEvent_t stage_free[hpsx::kNumStages] = {};
  uint32_t* d_src = nullptr;           // pooled path, lazily [cap_keys]
  float* d_result = nullptr;           // host-output lookups of a GPU session: [sum_t cap_t * dim_t], lazy
  float* d_pool_stage = nullptr;       // pooled path: all miss rows of one call
  size_t pool_stage_rows = 0;
  size_t max_dim = 0;
  std::vector<cudaEvent_t> ev;         // 2 per table: probe start / stop
  std::vector<cudaEvent_t> ev_pull;    // 2 per table: start / end of the direct-pull phase
  // address-sorted pull: [0] unsorted, [1] sorted host addresses / miss-list indices; CUB temp storage
  unsigned long long* d_addr[2] = {nullptr, nullptr};
  uint32_t* d_sidx[2] = {nullptr, nullptr};
  void* d_sort_temp = nullptr;
  size_t sort_temp_bytes = 0;
  hpsx_session_stats stats{};
  std::mutex mu;  // one lookup at a time per session (Triton guarantees it; tests may not)

  ~hpsx_session();
};

// Rank-local half of a model-parallel group (include/hpsx.h: hpsx_shard_group_*): one table of one model,
// rows sharded by owner_of(key, world) over the GPUs of one box.  The exchange arena is one cudaMalloc
// (so one CUDA IPC handle): [control words | inbox keys | inbox positions | output rows].
struct hpsx_shard_group {
  hpsx_session* s = nullptr;
  size_t table = 0;
  uint32_t rank = 0, world = 1;
  uint32_t slot_cap = 0;  // keys one rank may send to one owner = keys per request of this table
  size_t dim = 0;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0, off_keys = 0, off_pos = 0, off_out = 0;
  std::vector<unsigned char*> peer_arena;  // [world]; [rank] == arena
  std::vector<bool> peer_ipc;              // opened with cudaIpcOpenMemHandle (must be closed)
  hpsx::ShardPeers peers{};
  bool connected = false;
  uint32_t seq = 0;
  // miss list of the keys received from all peers (a rank can receive up to world * slot_cap keys)
  size_t miss_cap = 0;
  uint32_t* d_miss_pos = nullptr;
  int64_t* d_miss_keys = nullptr;
  int64_t* h_miss_keys = nullptr;   // mapped pinned mirror (staged miss path only)
  int64_t* hd_miss_keys = nullptr;
  uint32_t* h_ctrl = nullptr;       // pinned copy of the control words
  unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
  hpsx_shard_stats last{};

  // control words (uint32 index into the arena)
  // [kCursor, kDone] are local scratch, zeroed with one memset per lookup
  // [kSeen, kSeen + 16): after a flag-wait timeout, the last value read from every peer's flag cell (diagnosis)
  static constexpr uint32_t kCnt = 0, kFlagDispatch = 16, kFlagReturn = 32, kCursor = 48, kStatus = 64, kMissCount = 65,
                            kDone = 66, kSeen = 67, kWords = 96;
  uint32_t* ctrl() const { return reinterpret_cast<uint32_t*>(arena); }
  ~hpsx_shard_group();
};
