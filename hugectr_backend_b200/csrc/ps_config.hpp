// ps.json -> parameter structs.  Same key surface, defaults, mandatory-key errors and
// number-or-string leniency as the reference glue (hps_backend/src/backend.cpp:102-526,
// hps_backend/src/triton_helpers.cpp:42-442); parsed with the in-tree JSON DOM instead of TritonJson.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "json.hpp"

namespace hpsx {

enum class DatabaseType { Disabled, HashMap, ParallelHashMap, RedisCluster, RocksDB, Invalid };
enum class OverflowPolicy { EvictRandom, EvictLeastUsed, EvictOldest, Invalid };
enum class UpdateSourceType { Null, KafkaMessageQueue, Invalid };
enum class CacheType { Dynamic = 0, Static = 1, UVM = 2 };

// ~ HugeCTR::VolatileDatabaseParams (backend.cpp:129-216); defaults from
// docs/hierarchical_parameter_server.md:400-503.
struct VolatileDbConfig {
  DatabaseType type = DatabaseType::ParallelHashMap;
  std::string address = "127.0.0.1:7000", user_name = "default", password;
  size_t num_partitions = 0;  // 0 -> min(cores, 16)
  size_t allocation_rate = 256ull << 20;
  size_t hpsx_pull_window_mb = 0;  // engine extension: locality window of the direct pull (host_ps.hpp); 0 -> 16
  size_t max_batch_size = 65536;
  size_t overflow_margin = SIZE_MAX;
  OverflowPolicy overflow_policy = OverflowPolicy::EvictRandom;
  double overflow_resolution_target = 0.8;
  double initial_cache_rate = 1.0;
  bool cache_missed_embeddings = false;
  std::vector<std::string> update_filters{"^hps_.+$"};
};

// ~ HugeCTR::PersistentDatabaseParams (backend.cpp:219-259).  Parsed and carried; the RocksDB
// backend itself is out of scope.
struct PersistentDbConfig {
  DatabaseType type = DatabaseType::Disabled;
  std::string path;
  size_t num_threads = 16;
  bool read_only = false;
  size_t max_batch_size = 65536;
  std::vector<std::string> update_filters{"^hps_.+$"};
};

// ~ HugeCTR::UpdateSourceParams (backend.cpp:262-308).  Parsed and carried only.
struct UpdateSourceConfig {
  UpdateSourceType type = UpdateSourceType::Null;
  std::string brokers = "127.0.0.1:9092";
  size_t receive_buffer_size = 256 * 1024;
  size_t poll_timeout_ms = 500;
  size_t max_batch_size = 8192;
  size_t failure_backoff_ms = 50;
  size_t max_commit_interval = 32;
};

// ~ HugeCTR::InferenceParams as filled at backend.cpp:318-523.
struct ModelConfig {
  std::string model_name;
  std::string network_file, dense_file;
  size_t max_batch_size = 0;
  std::vector<std::string> sparse_files;
  int device_id = 0;
  bool use_gpu_embedding_cache = true;
  float hit_rate_threshold = 0.55f;
  float cache_size_percentage = 0.55f;
  bool i64_input_key = true;
  size_t number_of_worker_buffers_in_pool = 1;
  size_t number_of_refresh_buffers_in_pool = 1;
  float cache_refresh_percentage_per_iteration = 0.0f;
  std::vector<int> deployed_devices;
  std::vector<float> default_value_for_each_table;
  size_t maxnum_des_feature_per_sample = 26;
  std::vector<size_t> maxnum_catfeature_query_per_table_per_sample;
  std::vector<size_t> embedding_vecsize_per_table;
  std::vector<std::string> embedding_table_names;
  size_t label_dim = 1;
  size_t slot_num = 10;
  CacheType embedding_cache_type = CacheType::Dynamic;
  bool init_ec = true;
  bool fp8_quant = false;
  bool enable_pagelock = false;
  // engine extensions (optional keys, not in the reference's surface): see include/hpsx.h hpsx_model_params
  bool hpsx_split_lock = true;
  int hpsx_request_chunks = 0;
  int hpsx_pull_grid_ctas = 0;
  std::string hpsx_probe;  // "", "v8", "ldg", "tma"
  bool hpsx_peer_tier = false;  // NVLink tier over the deployed devices (needs enable_pagelock and >= 2 devices)
  // refresh knobs: carried for the Triton shell (model_state.cpp:312-335)
  float refresh_delay = 0.0f, refresh_interval = 0.0f;
};

struct PsConfig {
  bool support_int64_key = true;
  VolatileDbConfig volatile_db;
  PersistentDbConfig persistent_db;
  UpdateSourceConfig update_source;
  std::vector<ModelConfig> models;
};

// Error reporting: empty string = success.  `invalid_arg` messages use the reference's wording for
// missing mandatory keys so that log scrapers keep working.
struct ParseResult {
  bool ok = true;
  std::string message;
};

ParseResult parse_ps_config(const json::Value& root, PsConfig* out);
ParseResult parse_ps_config_file(const std::string& path, PsConfig* out);

// Enum string mappers with the reference's aliases (triton_helpers.cpp:183-339).
DatabaseType parse_database_type(std::string s);
OverflowPolicy parse_overflow_policy(std::string s);
UpdateSourceType parse_update_source_type(std::string s);
const char* to_string(DatabaseType t);
const char* to_string(OverflowPolicy t);
const char* to_string(UpdateSourceType t);

// Scalar readers shared with the Triton shell (config.pbtxt-as-JSON uses the same leniency).
// Return false when the key is absent; throw std::invalid_argument on a malformed value.
bool json_get(const json::Value& obj, const char* key, bool* v);
bool json_get(const json::Value& obj, const char* key, double* v);
bool json_get(const json::Value& obj, const char* key, float* v);
bool json_get(const json::Value& obj, const char* key, int32_t* v);
bool json_get(const json::Value& obj, const char* key, int64_t* v);
bool json_get(const json::Value& obj, const char* key, size_t* v);
bool json_get(const json::Value& obj, const char* key, std::string* v);
bool json_get(const json::Value& obj, const char* key, std::vector<std::string>* v);
bool json_get(const json::Value& obj, const char* key, std::vector<float>* v);
bool json_get(const json::Value& obj, const char* key, std::vector<int32_t>* v);
bool json_get(const json::Value& obj, const char* key, std::vector<size_t>* v);

}  // namespace hpsx
