/*
 * hpsx.h — C ABI of the B200-native Hierarchical Parameter Server engine (libhpsx.so).
 *
 * This is the thin FFI between the C++ host code of the Triton `hps` backend
 * (libtriton_hps.so, see include/triton_hps_backend.h) and the hand-written sm_100a
 * CUDA kernels.  Every entry point replaces one call the reference glue makes into the
 * un-vendored libhuge_ctr_hps.so; the reference call site is cited next to each function
 * (paths relative to /root/reference/hps_backend unless noted).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, opaque handles.  No C++/torch types.
 *   - every function returns HPSX_OK (0) or a negative hpsx_status; the message for the
 *     last failure on the calling thread is available from hpsx_last_error().
 *   - "h_" pointers are host memory, "d_" pointers are device memory on the handle's GPU.
 *   - keys are int64 ("supportlonglong": true, src/backend.cpp:124-126); vectors are fp32.
 *   - nothing here falls back to a CPU implementation of a GPU path: if no CUDA device is
 *     usable the GPU entry points fail with HPSX_ERR_CUDA.
 */
#ifndef HPSX_H_
#define HPSX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPSX_ABI_VERSION 5

typedef enum hpsx_status {
  HPSX_OK = 0,
  HPSX_ERR_INVALID_ARG = -1, /* maps to TRITONSERVER_ERROR_INVALID_ARG */
  HPSX_ERR_NOT_FOUND = -2,   /* unknown model / table / device */
  HPSX_ERR_UNSUPPORTED = -3, /* maps to TRITONSERVER_ERROR_UNSUPPORTED */
  HPSX_ERR_CUDA = -4,        /* CUDA runtime failure (reference throws: include/hps_buffer.hpp:79-87) */
  HPSX_ERR_IO = -5,          /* sparse model files unreadable */
  HPSX_ERR_INTERNAL = -6
} hpsx_status;

/* Opaque handles.  hpsx_ps      ~ HugeCTR::HierParameterServerBase   (src/backend.cpp:68-71)
 *                  hpsx_cache   ~ HugeCTR::EmbeddingCacheBase        (src/model_state.cpp:404-412)
 *                  hpsx_session ~ HugeCTR::LookupSessionBase         (src/model_instance_state.cpp:170-171) */
typedef struct hpsx_ps hpsx_ps;
typedef struct hpsx_cache hpsx_cache;
typedef struct hpsx_session hpsx_session;

typedef enum hpsx_cache_type {
  HPSX_CACHE_DYNAMIC = 0, /* LRU set-associative cache with insertion (default, src/backend.cpp:489-490) */
  HPSX_CACHE_STATIC = 1   /* loaded once, never replaced (src/backend.cpp:483-484) */
} hpsx_cache_type;

typedef enum hpsx_combiner { HPSX_COMBINER_SUM = 0, HPSX_COMBINER_MEAN = 1 } hpsx_combiner;

/* ~ HugeCTR::InferenceParams as filled by HPSBackend::ParseParameterServer (src/backend.cpp:318-523).
 * Arrays have `num_tables` entries and are copied by the callee. */
typedef struct hpsx_model_params {
  const char* model_name;                  /* "model"                       :325-328 */
  size_t max_batch_size;                   /* "max_batch_size"              :341-344 */
  size_t num_tables;                       /* = len(sparse_files)           :353-358 */
  const char* const* sparse_files;         /* may be NULL when tables are loaded from memory / procedurally */
  const char* const* table_names;          /* "embedding_table_names"       :462-467 (NULL -> sparse_embedding<i>) */
  const size_t* embedding_vecsize_per_table;                  /* :454-460 */
  const size_t* maxnum_catfeature_query_per_table_per_sample; /* :443-452 */
  const float* default_value_for_each_table;                  /* :427-433 */
  int use_gpu_embedding_cache;             /* "gpucache"                    :364-369 */
  float hit_rate_threshold;                /* "hit_rate_threshold"          :372-377 */
  float cache_size_percentage;             /* "gpucacheper"                 :380-385 */
  size_t number_of_worker_buffers_in_pool; /* "num_of_worker_buffer_in_pool":397-402 */
  const int* deployed_devices;             /* "deployed_device_list"        :418-425 */
  size_t num_deployed_devices;
  int embedding_cache_type;                /* hpsx_cache_type               :479-492 */
  /* engine extensions (not in the reference config; 0 selects the default) */
  float cache_load_factor;                 /* slots = gpucacheper*rows/load_factor, default 0.5 */
  int enable_pagelock;                     /* "enable_pagelock"             :506-511.  Page-locks the host
                                              tables; this engine then resolves cache misses on the GPU:
                                              kernels read the missing rows straight from host DRAM over
                                              PCIe through an HBM-resident key index ("direct pull"),
                                              instead of CPU gather + cudaMemcpyAsync */
  /* further engine extensions; ps.json spells them "hpsx_split_lock", "hpsx_request_chunks",
   * "hpsx_pull_grid_ctas", "hpsx_probe" (all optional) */
  int split_lock;                          /* >= 0 (default): instances that share a cache probe and pull under a shared
                                              lock and insert in a short exclusive section; < 0: whole-call exclusive lock */
  int request_chunks;                      /* a direct-pull request of >= 2^18 keys is cut into this many chunks so that
                                              the PCIe pull of chunk c overlaps the probe of chunk c+1; 0 -> 4 */
  int pull_grid_ctas;                      /* CTAs (of 256 threads) of the persistent binned pull kernel; 0 -> 148 */
  int probe_variant;                       /* probe+gather kernel, see hpsx_session_set_probe_variant; used when
                                              probe_variant_set != 0, else the default (4) */
  int probe_variant_set;
  int peer_tier;                           /* "hpsx_peer_tier": with enable_pagelock and >= 2 deployed devices, the rows of
                                              the host tables are also kept sharded over the devices' HBM and cache misses
                                              are read over NVLink instead of PCIe (hpsx_cache_peer_tier_*) */
} hpsx_model_params;

/* ~ HugeCTR::VolatileDatabaseParams, hash_map / parallel_hash_map only (src/backend.cpp:129-216). */
typedef struct hpsx_volatile_params {
  size_t num_partitions;     /* "num_partitions" :155-159; 0 -> min(cores,16) (docs/hierarchical_parameter_server.md:410-412) */
  size_t allocation_rate;    /* "allocation_rate" :161-165; 0 -> 256 MiB */
  double initial_cache_rate; /* "initial_cache_rate" :194-198; <=0 -> 1.0 */
  size_t num_threads;        /* worker pool; 0 -> HCTR_DEFAULT_CONCURRENCY or hardware_concurrency (src/thread_pool.cpp:25-41) */
  size_t pull_window_bytes;  /* engine extension (ps.json: volatile_db "hpsx_pull_window_mb"): host-memory window one
                                partition of a table — one bin of the direct-pull miss lists — should fit; 0 -> 16 MiB */
} hpsx_volatile_params;

/* Counters of one lookup session, cumulative since creation / last reset. */
typedef struct hpsx_session_stats {
  uint64_t lookups;            /* hpsx_session_lookup* calls */
  uint64_t keys;               /* key occurrences delivered */
  uint64_t hits;               /* served from the HBM cache */
  uint64_t misses;             /* went to the host parameter server */
  uint64_t inserted;           /* rows written into the cache */
  uint64_t default_filled;     /* rows answered with default_value (async mode / absent keys) */
  uint64_t h2d_bytes;          /* bytes copied host->device on behalf of lookups */
  uint64_t d2h_bytes;          /* bytes copied device->host on behalf of lookups */
  uint64_t kernel_launches;    /* CUDA kernels launched by this session */
  double probe_kernel_ms;      /* CUDA-event time of the probe+gather kernels (sum) */
  uint64_t probe_kernel_launches;
  uint64_t probe_kernel_keys;  /* keys those launches processed */
  double insert_kernel_ms;     /* CUDA-event time of the miss phase: first pull/merge kernel start to last insert end (sum) */
  double host_gather_ms;       /* wall time spent in the host parameter-server gather */
  double pull_kernel_ms;       /* binned direct pull: first pull kernel start to last pull kernel end (sum); the pulls of a
                                  chunked request run beside the probes of its later chunks */
  uint64_t tier_bytes;         /* bytes of missed rows read from the NVLink tier (peer or local HBM) instead of over PCIe;
                                  not part of h2d_bytes */
} hpsx_session_stats;

/* ---------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------- */
int hpsx_abi_version(void);
/* Message of the last failing call on this thread ("" if none).  Valid until the next call. */
const char* hpsx_last_error(void);
/* Number of CUDA devices the engine can use (0 on a CPU-only box, never an error). */
int hpsx_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * parameter server  (~ HierParameterServerBase)
 * ------------------------------------------------------------------------------------------- */
/* ~ HierParameterServerBase::create(ps_json_path)                         src/backend.cpp:68-69
 * Parses ps.json, loads every model's sparse files into the host (volatile) database and, for
 * gpucache models with init_ec, creates their embedding caches on deployed_device_list. */
int hpsx_ps_create_from_json(const char* ps_json_path, hpsx_ps** out);
/* Programmatic creation without a file (tests, bench). `vdb` may be NULL for defaults. */
int hpsx_ps_create(const hpsx_volatile_params* vdb, hpsx_ps** out);
int hpsx_ps_destroy(hpsx_ps* ps);

/* ~ get_hps_model_configuration_map()                                     src/backend.cpp:70-71 */
int hpsx_ps_num_models(const hpsx_ps* ps, size_t* out);
int hpsx_ps_model_name(const hpsx_ps* ps, size_t index, const char** out);
int hpsx_ps_has_model(const hpsx_ps* ps, const char* model_name);

/* ~ update_database_per_model(InferenceParams)                            src/model_state.cpp:389
 * Registers the model and (when sparse_files != NULL) loads `<dir>/key` + `<dir>/emb_vector`
 * (docs/architecture.md:185-218) of every table into the host database. */
int hpsx_ps_add_model(hpsx_ps* ps, const hpsx_model_params* params);
/* Insert/overwrite rows of one table from host arrays (key file + vector file contents). */
int hpsx_ps_load_table(hpsx_ps* ps, const char* model, size_t table, const int64_t* h_keys,
                       const float* h_vectors, size_t num_rows);
/* Fill one table with keys [0,num_rows) and procedural rows
 *   row(k)[j] = bitcast_f32(0x3F800000 | (splitmix64(k*131 + j + seed) >> 41)) - 1.5
 *  — the synthetic table of the benchmark configs; no files needed. */
int hpsx_ps_load_table_procedural(hpsx_ps* ps, const char* model, size_t table, size_t num_rows,
                                  uint64_t seed);
/* Model-parallel variant: of the keys [0,num_rows) load only those with
 * hpsx_owner(key, num_shards) == shard — one shard of a table too large for one GPU / one host. */
int hpsx_ps_load_table_procedural_shard(hpsx_ps* ps, const char* model, size_t table, size_t num_rows,
                                        uint64_t seed, uint32_t shard, uint32_t num_shards);
int hpsx_ps_table_rows(const hpsx_ps* ps, const char* model, size_t table, size_t* out);
/* ~ get_hps_model_configuration_map().at(model) -> InferenceParams      src/backend.cpp:70-71, hps.cc:221-223
 * Fills `out` with pointers into the server's own storage; they stay valid until the server is
 * destroyed.  cache_load_factor is the effective value. */
int hpsx_ps_get_model_params(hpsx_ps* ps, const char* model, hpsx_model_params* out);
/* ~ HPSBackend::ParseParameterServer(ps.json) for online deployment       src/hps.cc:207-219, src/backend.cpp:102-526
 * Re-reads ps.json and registers (and loads the sparse files of) every model that the server does
 * not know yet; known models are left untouched.  `num_added` may be NULL. */
int hpsx_ps_sync_models_from_json(hpsx_ps* ps, const char* ps_json_path, size_t* num_added);

/* ~ HierParameterServerBase::lookup(h_keys, n, h_vectors, model, table)   (CPU path, gpucache=false;
 * semantics docs/hierarchical_parameter_server.md:67-78,244-246): volatile-db fetch, absent keys get
 * default_value_for_each_table[table]. `h_vectors` is [n, vecsize] row-major. */
int hpsx_ps_lookup(hpsx_ps* ps, const char* model, size_t table, const int64_t* h_keys, size_t n,
                   float* h_vectors);

/* ---------------------------------------------------------------------------------------------
 * embedding cache  (~ EmbeddingCacheBase, one per (model, device))
 * ------------------------------------------------------------------------------------------- */
/* ~ create_embedding_cache_per_model(InferenceParams)                     src/model_state.cpp:391
 * Allocates the HBM hash table + value slab of every table on every deployed device and warms it
 * with the first `gpucacheper` fraction of each table's rows. */
int hpsx_ps_create_embedding_cache_per_model(hpsx_ps* ps, const char* model);
/* ~ get_embedding_cache(model, device) — NULL/NOT_FOUND when absent        src/model_state.cpp:379,411 */
int hpsx_ps_get_embedding_cache(hpsx_ps* ps, const char* model, int device, hpsx_cache** out);
/* ~ update_database_per_model(InferenceParams)                             src/model_state.cpp:132,389
 * Re-reads the model's sparse_files into the host database: rows are inserted or overwritten in place.  Tables
 * that are page-locked (enable_pagelock) get their new rows registered and their HBM row index rebuilt, with
 * lookups excluded meanwhile.  HBM caches keep the vectors they hold until hpsx_ps_refresh_embedding_cache. */
int hpsx_ps_update_database_per_model(hpsx_ps* ps, const char* model);
/* ~ refresh_embedding_cache(model, device)                                  src/model_state.cpp:135,161
 * Re-reads the vector of every key resident in the (model, device) cache from the host database and rewrites
 * the cached row (value-update kernel), cache_refresh_percentage_per_iteration of the cache per exclusive
 * section.  Residency does not change.  `refreshed_rows` (nullable) receives the number of rows rewritten. */
int hpsx_ps_refresh_embedding_cache(hpsx_ps* ps, const char* model, int device, size_t* refreshed_rows);
/* ~ destory_embedding_cache_per_model(model)                              src/model_state.cpp:111 */
int hpsx_ps_destroy_embedding_cache_per_model(hpsx_ps* ps, const char* model);
/* ~ get_cache_config().num_emb_table_                                     src/model_instance_state.cpp:107-109 */
int hpsx_cache_num_tables(const hpsx_cache* cache, size_t* out);
int hpsx_cache_device(const hpsx_cache* cache, int* out);
/* Slots allocated / keys resident in one table's HBM cache (resident is counted on the device). */
int hpsx_cache_capacity(const hpsx_cache* cache, size_t table, size_t* slots);
int hpsx_cache_resident(hpsx_cache* cache, size_t table, size_t* keys);
/* Copy the resident keys of one table to host (order unspecified). `cap` entries available. */
int hpsx_cache_dump_keys(hpsx_cache* cache, size_t table, int64_t* h_keys, size_t cap, size_t* n);

/* ---------------------------------------------------------------------------------------------
 * lookup session  (~ LookupSessionBase, one per model instance)
 * ------------------------------------------------------------------------------------------- */
/* ~ LookupSessionBase::create(inference_params, embedding_cache)          src/model_instance_state.cpp:170-171
 * `device` < 0 or a model with gpucache=false gives a CPU session (vectors returned in host memory,
 * src/hps.cc:640-642). */
int hpsx_session_create(hpsx_ps* ps, const char* model, int device, hpsx_session** out);
int hpsx_session_destroy(hpsx_session* s);
int hpsx_session_device(const hpsx_session* s, int* out);
/* The CUDA stream (cudaStream_t) the session launches on, for callers that time with events. */
int hpsx_session_stream(const hpsx_session* s, void** out);

/* ~ LookupSessionBase::lookup(h_keys_per_table, d_vectors_per_table, num_keys_per_table)
 *                                                                          src/model_instance_state.cpp:194-195
 * For every table t: vectors_per_table[t][i*d_t .. (i+1)*d_t) = row_t(keys_per_table[t][i]) or the
 * table's default value.  Blocks until the vectors are complete.  With a GPU session the vector
 * pointers are device memory (they may be the Triton output buffer itself); with a CPU session
 * they are host memory. */
int hpsx_session_lookup(hpsx_session* s, const void* const* h_keys_per_table,
                        float* const* vectors_per_table, const size_t* num_keys_per_table,
                        size_t num_tables);
/* Same, keys already resident in device memory (the engine-ABI arm of bench.py). */
int hpsx_session_lookup_device_keys(hpsx_session* s, const int64_t* const* d_keys_per_table,
                                    float* const* d_vectors_per_table,
                                    const size_t* num_keys_per_table, size_t num_tables);
/* Where a buffer handed to hpsx_session_lookup_ex lives. */
typedef enum hpsx_memory { HPSX_MEM_HOST = 0, HPSX_MEM_DEVICE = 1 } hpsx_memory;
/* General form used by the Triton shell, which must take whatever buffers Triton hands it
 * (src/hps.cc:586-597 input buffer, :638-648 output buffer whose memory type Triton may override):
 * keys and vectors each in host or device memory.  GPU session + host vectors: the rows are
 * gathered into the session's device result buffer and copied D2H (reference: hps.cc:681-685).
 * CPU session: both must be host memory. */
int hpsx_session_lookup_ex(hpsx_session* s, const void* const* keys_per_table, int key_memory,
                           float* const* vectors_per_table, int vector_memory,
                           const size_t* num_keys_per_table, size_t num_tables);
/* Cross-request batching (the reference serves the requests of one Execute call one by one,
 * src/hps.cc:392-406): `num_requests` <= HPSX_MAX_BATCH_REQUESTS requests in ONE pass — all probes are launched
 * back to back, the host waits once for all miss counts, every miss list is resolved, and the host waits once more.
 * Arrays are request-major: entry [r * T + t] is table t of request r (T = tables of the model).  Falls back to one
 * hpsx_session_lookup_ex per request for CPU sessions, host output buffers, or when the batch holds more keys than
 * one request may (max_batch_size * sum of maxnum_catfeature). */
#define HPSX_MAX_BATCH_REQUESTS 16
int hpsx_session_lookup_batch(hpsx_session* s, size_t num_requests, const void* const* keys, int key_memory,
                              float* const* vectors, int vector_memory, const size_t* num_keys);
/* Lookup of one table that ALSO writes a bf16 mirror of the vectors (d_vectors_bf16: device, [n, vecsize] bf16,
 * 16-byte aligned; d_vectors: device fp32, 32-byte aligned), produced by the same kernels that write the fp32 rows
 * (probe+gather and the miss kernels) — the elementwise conversion the dense head needs is fused into the
 * producer.  Insertion is synchronous for this call.  Rows must be multiples of 8 floats. */
int hpsx_session_lookup_bf16_mirror(hpsx_session* s, size_t table, const int64_t* keys, int key_memory, size_t n,
                                    float* d_vectors, void* d_vectors_bf16);
/* Model-parallel return leg fused into the gather: key i of `table` is delivered to row
 * d_pos[i] of d_out_base, which may be ANOTHER GPU's buffer opened with hpsx_ipc_open — the rows then leave
 * the owner's gather kernel as NVLink peer stores, no all-to-all of vectors and no scatter pass.  Misses are
 * resolved as usual and land in the same rows.  Blocks until this GPU's stores have been issued and the
 * kernels have completed. */
int hpsx_session_lookup_scatter(hpsx_session* s, size_t table, const int64_t* d_keys, const uint32_t* d_pos,
                                size_t n, float* d_out_base);
/* Device buffers that other processes of the box (one per GPU) can map: cudaMalloc + CUDA IPC handles
 * (64 opaque bytes, exchanged by the caller, e.g. over torch.distributed). */
int hpsx_device_malloc(int device, size_t bytes, void** d_ptr);
int hpsx_device_free(int device, void* d_ptr);
int hpsx_ipc_export(int device, void* d_ptr, void* handle64);
int hpsx_ipc_open(int device, const void* handle64, void** d_ptr);
int hpsx_ipc_close(int device, void* d_ptr);

/* ---------------------------------------------------------------------------------------------
 * model-parallel group: rows of one table sharded over the GPUs of one box by
 * hpsx_owner(key, world); the reference has no such mode (one full cache per device,
 * hps_backend/src/model_state.cpp:395-419).  One hpsx_shard_group per rank, built on that rank's session
 * (whose parameter server holds the shard owner == rank).  A lookup is ONE fused exchange over NVLink peer
 * memory: keys + request positions are stored straight into the owners' inboxes, each owner's probe+gather
 * kernel stores the rows straight into the requesters' output buffers, and two flag waves (device-side, no
 * NCCL call, no host round trip) order the steps.  Every rank of the group must call lookup the same number of
 * times (it is a collective); misses are always resolved synchronously.
 * ------------------------------------------------------------------------------------------- */
typedef struct hpsx_shard_group hpsx_shard_group;
#define HPSX_SHARD_HANDLE_BYTES 64 /* one CUDA IPC memory handle */
typedef struct hpsx_shard_stats {
  uint64_t keys_sent_remote;     /* keys of the last request owned by other ranks */
  uint64_t keys_received;        /* keys this rank looked up for the group in the last request (own included) */
  uint64_t keys_received_remote;
  uint64_t misses;               /* of keys_received */
  uint32_t status;               /* 0 ok; bit 0 a rank failed, bit 1 timeout, bit 2 capacity */
  uint32_t sent[16];             /* per owner */
  uint32_t received[16];         /* per requester */
} hpsx_shard_stats;
/* Allocates the rank's exchange arena (inbox for world x cap keys, output for cap rows; cap = the session's
 * key capacity for `table`) and writes its IPC handle to handle64 (may be NULL for same-process groups). */
int hpsx_shard_group_create(hpsx_session* s, size_t table, uint32_t rank, uint32_t world, hpsx_shard_group** out,
                            void* handle64);
/* One process per GPU: all_handles = world x 64 bytes, rank-major (exchanged by the caller, e.g. all_gather). */
int hpsx_shard_group_connect_ipc(hpsx_shard_group* g, const void* all_handles);
/* All ranks in this process (e.g. one Triton server driving 8 GPUs): groups[world], entry `rank` == g. */
int hpsx_shard_group_connect_local(hpsx_shard_group* g, hpsx_shard_group* const* groups);
/* Collective.  d_keys: this rank's request (device memory, n <= cap).  On return *d_out is this rank's output
 * buffer [n, dim] (owned by the group, valid until the next lookup), rows in request order. */
int hpsx_shard_group_lookup(hpsx_shard_group* g, const int64_t* d_keys, size_t n, float** d_out);
int hpsx_shard_group_get_stats(const hpsx_shard_group* g, hpsx_shard_stats* out);
/* Rows of the group's output buffer (= keys one request may hold); the buffer address never changes. */
int hpsx_shard_group_capacity(const hpsx_shard_group* g, size_t* rows);
int hpsx_shard_group_set_timeout_ms(hpsx_shard_group* g, uint64_t ms);
int hpsx_shard_group_destroy(hpsx_shard_group* g);

/* ---------------------------------------------------------------------------------------------
 * NVLink tier (engine extension; ps.json model key "hpsx_peer_tier").  The reference deploys one full cache per
 * device and every device's cache misses go to the one host parameter server (src/model_state.cpp:395-419,
 * include/backend.hpp:70-74) — over a host fabric that, measured on this pool, gives eight GPUs less than half the
 * per-GPU rate it gives one.  With the tier the rows of a page-locked (enable_pagelock) host table are ALSO kept
 * sharded over the HBM of the box's GPUs: rank r holds the rows with hpsx_owner(key, world) == r, every rank maps every
 * shard (peer access inside one process, CUDA IPC between processes), and the key -> row-address index the
 * direct-pull kernels use points into the shards.  A cache miss is then read over NVLink from its owner's HBM (or from
 * local HBM) by the same kernels, one-sidedly: no collective, no flag, the owner's SMs are not involved.  Values are
 * those of the host table at build time; hpsx_ps_update_database_per_model points the index back at host memory
 * (and, for a model with "hpsx_peer_tier", rebuilds the tier of all its caches in this process).
 *   one process, several devices:  hpsx_ps_peer_tier_connect_local (automatic for "hpsx_peer_tier": true)
 *   one process per device:        build -> export (every table) -> exchange handles -> attach_ipc (every peer and
 *                                  table) -> commit; detach on every rank before any rank destroys its cache.
 * ------------------------------------------------------------------------------------------- */
typedef struct hpsx_peer_tier_info {
  uint32_t rank, world;           /* world == 0: no tier */
  int committed;                  /* the index points at the shards */
  uint64_t own_rows;              /* rows in this rank's shards (all tables) */
  uint64_t own_bytes;             /* HBM the shards occupy */
  uint64_t index_entries_in_tier; /* keys whose misses are served from a shard (all tables) */
} hpsx_peer_tier_info;
int hpsx_cache_peer_tier_build(hpsx_cache* cache, uint32_t rank, uint32_t world);
/* handle64: HPSX_SHARD_HANDLE_BYTES; rows / cap: what the importing rank passes to attach_ipc */
int hpsx_cache_peer_tier_export(hpsx_cache* cache, size_t table, void* handle64, uint64_t* rows, uint64_t* cap);
int hpsx_cache_peer_tier_attach_ipc(hpsx_cache* cache, size_t table, uint32_t peer, const void* handle64, uint64_t rows,
                                    uint64_t cap);
/* same process (all tables of `peer_cache`, which was built as rank `peer` of the same world) */
int hpsx_cache_peer_tier_attach_local(hpsx_cache* cache, uint32_t peer, hpsx_cache* peer_cache);
int hpsx_cache_peer_tier_commit(hpsx_cache* cache);
/* Index back to host memory, peers unmapped, own shards freed (other ranks must have detached first). */
int hpsx_cache_peer_tier_detach(hpsx_cache* cache);
int hpsx_cache_peer_tier_info(hpsx_cache* cache, hpsx_peer_tier_info* out);
/* Every cache of `model` in this process becomes one rank of a tier (ranks in ascending device order). */
int hpsx_ps_peer_tier_connect_local(hpsx_ps* ps, const char* model);

/* Blocking device -> host copy on `device` (small control tensors such as NUMKEYS that Triton
 * delivered in GPU memory). */
int hpsx_copy_to_host(int device, void* h_dst, const void* d_src, size_t bytes);

/* Fused slot-wise gather + reduce (north-star stage a8): keys of table `table` laid out
 * [num_bags, hotness]; d_pooled[b*d .. ) = sum_j row(key[b,j]) (MEAN: divided by hotness), fp32,
 * accumulated in ascending j.  Misses are resolved (sync insert) before pooling. */
int hpsx_session_lookup_pooled(hpsx_session* s, size_t table, const int64_t* h_keys,
                               size_t num_bags, size_t hotness, int combiner, float* d_pooled);
int hpsx_session_lookup_pooled_device_keys(hpsx_session* s, size_t table, const int64_t* d_keys,
                                           size_t num_bags, size_t hotness, int combiner,
                                           float* d_pooled);

/* General form (keys and pooled vectors each in host or device memory; CPU sessions pool on the host
 * in the same ascending-slot fp32 order) — what the Triton shell calls when a model opts into pooling. */
int hpsx_session_lookup_pooled_ex(hpsx_session* s, size_t table, const int64_t* keys, int key_memory,
                                  size_t num_bags, size_t hotness, int combiner, float* pooled,
                                  int pooled_memory);

int hpsx_session_get_stats(const hpsx_session* s, hpsx_session_stats* out);
int hpsx_session_reset_stats(hpsx_session* s);
/* Force the insertion mode of subsequent lookups: <0 use hit_rate_threshold (default), 0 always
 * asynchronous (misses answered with the default vector), 1 always synchronous. */
int hpsx_session_set_insert_mode(hpsx_session* s, int mode);
/* Select the probe+gather kernel (default 4, or the model's "hpsx_probe"): 4 = 256-bit row vectors with L2
 * evict_first and bucket keys kept in L2 (rows must be multiples of 32 B, else 0 is used), 0 = LDG.128 register
 * copies (any row size), 1 = bulk-async (TMA engine) row staging through shared memory. */
int hpsx_session_set_probe_variant(hpsx_session* s, int variant);
/* Measurement aid for the binned direct pull (bench experiments; not a serving knob): bit 0 = insert only after ALL
 * pulls of the request, bit 1 = start the pulls only after ALL probes, bit 2 = print a per-kernel timeline of every
 * request on stderr, bit 3 = the one-row form of the pull kernel also on small host tables (default there: four rows in flight per warp). */
int hpsx_session_set_debug(hpsx_session* s, int flags);
/* Block until background (asynchronous) insertions queued by this session's cache are done. */
int hpsx_cache_drain_async(hpsx_cache* cache);

/* ---------------------------------------------------------------------------------------------
 * dense MLP head: the dense model that follows the hps model in the reference's ensembles
 * (samples/hps-triton-ensemble/01_model_training.ipynb cells 7,11: fc_1 -> fc_2 -> fc_3 over the reshaped lookup
 * vectors; 02_model_inference_hps_tf_ensemble.ipynb:336-395 deploys it as a second Triton model).  Here it reads the
 * lookup's device output in place.  Y = act(X W^T + b) per layer, bf16 operands with fp32 accumulation on the
 * tcgen05 tensor cores; the result is fp32.
 * ------------------------------------------------------------------------------------------- */
typedef struct hpsx_mlp hpsx_mlp;
/* dims[num_layers + 1]; weights[l]: host fp32 [dims[l+1], dims[l]] row-major; biases[l]: host fp32 [dims[l+1]] or
 * NULL (biases may be NULL altogether); relu[l] != 0 applies max(x, 0) (relu may be NULL = linear layers, like the
 * sample).  Every dims[l] (layer input width) must be a multiple of 8; a layer with one output unit must be last. */
int hpsx_mlp_create(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                    const float* const* biases, const int* relu, hpsx_mlp** out);
/* Same with the arithmetic chosen: HPSX_MLP_BF16 (above), or HPSX_MLP_TF32 — weights and activations stay fp32 in
 * memory, hpsx_mlp_forward reads the lookup's fp32 vectors in place (no conversion pass), the tensor cores round the
 * operands to TF32 and accumulate in fp32: within 1e-3 (relative to the output's scale) of a pure fp32 model, tested
 * (tests/test_dense_mlp_gpu.py); the reference's dense model is fp32 (01_model_training.ipynb cells 7,11).
 * hpsx_mlp_forward_bf16 is not available on a TF32 head. */
typedef enum hpsx_mlp_precision { HPSX_MLP_BF16 = 0, HPSX_MLP_TF32 = 1 } hpsx_mlp_precision;
int hpsx_mlp_create_ex(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                       const float* const* biases, const int* relu, int precision, hpsx_mlp** out);
/* d_in: device fp32 [batch, dims[0]]; d_out: device fp32 [batch, dims[num_layers]].  Asynchronous on `stream`
 * (a cudaStream_t, NULL = default stream). */
int hpsx_mlp_forward(hpsx_mlp* m, const float* d_in, size_t batch, float* d_out, void* stream);
/* Same with bf16 activations [batch, dims[0]] (16-byte aligned) written by hpsx_session_lookup_bf16_mirror: no
 * conversion pass. */
int hpsx_mlp_forward_bf16(hpsx_mlp* m, const void* d_in_bf16, size_t batch, float* d_out, void* stream);
int hpsx_mlp_destroy(hpsx_mlp* m);

/* ---------------------------------------------------------------------------------------------
 * stand-alone device primitives (testable without a parameter server)
 * ------------------------------------------------------------------------------------------- */
/* K1: dedup `n` device keys.  d_unique[0..*h_num_unique) holds each distinct key
 * once (order unspecified), d_inverse[i] is the index into d_unique of d_keys[i].  `stream` is a
 * cudaStream_t (NULL = default stream); the call synchronises that stream before returning. */
int hpsx_unique(int device, const int64_t* d_keys, size_t n, int64_t* d_unique, uint32_t* d_inverse,
                size_t* h_num_unique, void* stream);
/* Shard routing of the model-parallel mode: owner(key) in [0, num_shards). */
uint32_t hpsx_owner(int64_t key, uint32_t num_shards);
/* Host-side batch form: h_owners[i] = hpsx_owner(h_keys[i], num_shards). */
int hpsx_owner_batch(const int64_t* h_keys, size_t n, uint32_t num_shards, uint32_t* h_owners);
/* Bucket `n` device keys by owner: d_counts[num_shards] (u32), d_perm[n] = original positions grouped
 * by owner (stable within a CTA-tile, unspecified across), d_routed_keys[n] = keys in that order.
 * h_counts receives the counts.  Synchronises `stream`. */
int hpsx_route_keys(int device, const int64_t* d_keys, size_t n, uint32_t num_shards,
                    int64_t* d_routed_keys, uint32_t* d_perm, uint32_t* d_counts, uint32_t* h_counts,
                    void* stream);
/* out[perm[i]*d .. ) = rows[i*d .. ) — return path of routed lookups (inverse permutation). */
int hpsx_scatter_rows(int device, const float* d_rows, const uint32_t* d_perm, size_t n, size_t d,
                      float* d_out, void* stream);

/* Measurement primitive (bench.py): d_out[i] = d_table[d_idx[i]] for rows of 128 floats, same launch
 * shape as the probe+gather kernel without hashing — the practical "HBM random-gather" ceiling. */
int hpsx_gather_rows(int device, const float* d_table, const uint32_t* d_idx, size_t n, size_t dim,
                     float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPSX_H_ */
