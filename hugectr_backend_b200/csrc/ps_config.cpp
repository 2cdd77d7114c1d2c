#include "ps_config.hpp"

#include <algorithm>
#include <cctype>
#include <fstream>
#include <iterator>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace hpsx {
namespace {

std::string lower(std::string s) {
  for (char& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  return s;
}

// scalar leniency of the reference: a JSON number/bool, or a string holding one
// (triton_helpers.cpp:47-60,74-79,132-137,151-156)
bool scalar_bool(const json::Value& v, const char* key) {
  if (v.is_bool()) return v.as_bool();
  if (v.is_number()) return v.as_double() != 0.0;
  if (v.is_string()) {
    const std::string t = lower(v.as_string());
    if (t == "true") return true;
    if (t == "false") return false;
    return std::stoll(t) != 0;
  }
  throw std::invalid_argument(std::string("The parameter '") + key + "' is not a boolean.");
}
double scalar_double(const json::Value& v, const char* key) {
  if (v.is_number()) return v.as_double();
  if (v.is_string()) return std::stod(v.as_string());
  throw std::invalid_argument(std::string("The parameter '") + key + "' is not a number.");
}
int64_t scalar_int(const json::Value& v, const char* key) {
  if (v.is_number()) {
    if (!v.is_integer())
      throw std::invalid_argument(std::string("The parameter '") + key + "' is not an integer.");
    return v.as_int();
  }
  if (v.is_string()) return std::stoll(v.as_string());
  throw std::invalid_argument(std::string("The parameter '") + key + "' is not an integer.");
}
size_t scalar_size(const json::Value& v, const char* key) {
  if (v.is_number()) {
    if (!v.is_integer() || v.as_double() < 0)
      throw std::invalid_argument(std::string("The parameter '") + key +
                                  "' is not an unsigned integer.");
    // literals above INT64_MAX (overflow_margin = 2^64-1) must not wrap
    return static_cast<size_t>(std::stoull(v.raw_number()));
  }
  if (v.is_string()) return static_cast<size_t>(std::stoull(v.as_string()));
  throw std::invalid_argument(std::string("The parameter '") + key + "' is not an unsigned integer.");
}

std::string mandatory_error(const char* key) {
  return std::string("The parameter '") + key +
         "' is mandatory. Please confirm that it has been added to the configuration file.";
}

template <typename T>
void need(const json::Value& obj, const char* key, T* v, bool required) {
  if (!json_get(obj, key, v) && required) throw std::invalid_argument(mandatory_error(key));
}

std::string normalise_enum(std::string s, bool dash_too) {
  for (char& c : s) {
    if (c == ' ' || (dash_too && c == '-'))
      c = '_';
    else
      c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  }
  return s;
}

}  // namespace

bool json_get(const json::Value& obj, const char* key, bool* v) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  *v = scalar_bool(*m, key);
  return true;
}
bool json_get(const json::Value& obj, const char* key, double* v) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  *v = scalar_double(*m, key);
  return true;
}
bool json_get(const json::Value& obj, const char* key, float* v) {
  double d = *v;
  if (!json_get(obj, key, &d)) return false;
  if (d < std::numeric_limits<float>::lowest() || d > std::numeric_limits<float>::max()) {
    std::ostringstream os;
    os << "The parameter '" << key << "' = " << d << " was truncated because it is out of bounds!";
    throw std::invalid_argument(os.str());
  }
  *v = static_cast<float>(d);
  return true;
}
bool json_get(const json::Value& obj, const char* key, int64_t* v) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  *v = scalar_int(*m, key);
  return true;
}
bool json_get(const json::Value& obj, const char* key, int32_t* v) {
  int64_t t = *v;
  if (!json_get(obj, key, &t)) return false;
  if (t < std::numeric_limits<int32_t>::min() || t > std::numeric_limits<int32_t>::max()) {
    std::ostringstream os;
    os << "The parameter '" << key << "' = " << t << " was truncated because it is out of bounds!";
    throw std::invalid_argument(os.str());
  }
  *v = static_cast<int32_t>(t);
  return true;
}
bool json_get(const json::Value& obj, const char* key, size_t* v) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  *v = scalar_size(*m, key);
  return true;
}
bool json_get(const json::Value& obj, const char* key, std::string* v) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  if (!m->is_string())
    throw std::invalid_argument(std::string("The parameter '") + key + "' is not a string.");
  *v = m->as_string();
  return true;
}

namespace {
template <typename T, typename F>
bool get_vector(const json::Value& obj, const char* key, std::vector<T>* v, F&& conv) {
  const json::Value* m = obj.find(key);
  if (!m) return false;
  if (!m->is_array())
    throw std::invalid_argument(std::string("The parameter '") + key + "' is not an array.");
  v->clear();
  for (const json::Value& e : m->items()) v->push_back(conv(e));
  return true;
}
}  // namespace

bool json_get(const json::Value& obj, const char* key, std::vector<std::string>* v) {
  return get_vector(obj, key, v, [&](const json::Value& e) {
    if (!e.is_string())
      throw std::invalid_argument(std::string("The parameter '") + key +
                                  "' must be an array of strings.");
    return e.as_string();
  });
}
bool json_get(const json::Value& obj, const char* key, std::vector<float>* v) {
  return get_vector(obj, key, v,
                    [&](const json::Value& e) { return static_cast<float>(scalar_double(e, key)); });
}
bool json_get(const json::Value& obj, const char* key, std::vector<int32_t>* v) {
  return get_vector(obj, key, v,
                    [&](const json::Value& e) { return static_cast<int32_t>(scalar_int(e, key)); });
}
bool json_get(const json::Value& obj, const char* key, std::vector<size_t>* v) {
  return get_vector(obj, key, v, [&](const json::Value& e) { return scalar_size(e, key); });
}

DatabaseType parse_database_type(std::string s) {
  s = normalise_enum(std::move(s), true);
  for (const char* n : {"disabled", "disable", "none"})
    if (s == n) return DatabaseType::Disabled;
  for (const char* n : {"hash_map", "hashmap", "hash", "map"})
    if (s == n) return DatabaseType::HashMap;
  for (const char* n : {"parallel_hash_map", "parallel_hashmap", "parallel_hash", "parallel_map"})
    if (s == n) return DatabaseType::ParallelHashMap;
  for (const char* n : {"redis_cluster", "redis"})
    if (s == n) return DatabaseType::RedisCluster;
  for (const char* n : {"rocks_db", "rocksdb", "rocks"})
    if (s == n) return DatabaseType::RocksDB;
  return DatabaseType::Invalid;
}
OverflowPolicy parse_overflow_policy(std::string s) {
  s = normalise_enum(std::move(s), false);
  for (const char* n : {"evict_random", "random"})
    if (s == n) return OverflowPolicy::EvictRandom;
  for (const char* n : {"evict_least_used", "least_used"})
    if (s == n) return OverflowPolicy::EvictLeastUsed;
  for (const char* n : {"evict_oldest", "oldest"})
    if (s == n) return OverflowPolicy::EvictOldest;
  return OverflowPolicy::Invalid;
}
UpdateSourceType parse_update_source_type(std::string s) {
  s = normalise_enum(std::move(s), false);
  for (const char* n : {"null", "none"})
    if (s == n) return UpdateSourceType::Null;
  for (const char* n : {"kafka_message_queue", "kafka_mq", "kafka"})
    if (s == n) return UpdateSourceType::KafkaMessageQueue;
  return UpdateSourceType::Invalid;
}
const char* to_string(DatabaseType t) {
  switch (t) {
    case DatabaseType::Disabled: return "disabled";
    case DatabaseType::HashMap: return "hash_map";
    case DatabaseType::ParallelHashMap: return "parallel_hash_map";
    case DatabaseType::RedisCluster: return "redis_cluster";
    case DatabaseType::RocksDB: return "rocks_db";
    default: return "<invalid>";
  }
}
const char* to_string(OverflowPolicy t) {
  switch (t) {
    case OverflowPolicy::EvictRandom: return "evict_random";
    case OverflowPolicy::EvictLeastUsed: return "evict_least_used";
    case OverflowPolicy::EvictOldest: return "evict_oldest";
    default: return "<invalid>";
  }
}
const char* to_string(UpdateSourceType t) {
  switch (t) {
    case UpdateSourceType::Null: return "null";
    case UpdateSourceType::KafkaMessageQueue: return "kafka_message_queue";
    default: return "<invalid>";
  }
}

namespace {

template <typename E, typename F>
void get_enum(const json::Value& obj, const char* key, E* v, F&& map, const char* type_name) {
  std::string s;
  if (!json_get(obj, key, &s) || s.empty()) return;  // optional: keep the default
  const E e = map(s);
  if (e == E::Invalid)
    throw std::invalid_argument(std::string("Unable to map parameter '") + key + "' = \"" + s +
                                "\" to " + type_name + "!");
  *v = e;
}

void parse_volatile(const json::Value& j, VolatileDbConfig* p) {
  get_enum(j, "type", &p->type, parse_database_type, "DatabaseType_t");
  need(j, "address", &p->address, false);
  need(j, "user_name", &p->user_name, false);
  need(j, "password", &p->password, false);
  need(j, "num_partitions", &p->num_partitions, false);
  need(j, "allocation_rate", &p->allocation_rate, false);
  need(j, "hpsx_pull_window_mb", &p->hpsx_pull_window_mb, false);
  need(j, "max_batch_size", &p->max_batch_size, false);
  need(j, "overflow_margin", &p->overflow_margin, false);
  get_enum(j, "overflow_policy", &p->overflow_policy, parse_overflow_policy,
           "DatabaseOverflowPolicy_t");
  need(j, "overflow_resolution_target", &p->overflow_resolution_target, false);
  need(j, "initial_cache_rate", &p->initial_cache_rate, false);
  need(j, "cache_missed_embeddings", &p->cache_missed_embeddings, false);
  need(j, "update_filters", &p->update_filters, false);
}

void parse_persistent(const json::Value& j, PersistentDbConfig* p) {
  get_enum(j, "type", &p->type, parse_database_type, "DatabaseType_t");
  need(j, "path", &p->path, false);
  need(j, "num_threads", &p->num_threads, false);
  need(j, "read_only", &p->read_only, false);
  need(j, "max_batch_size", &p->max_batch_size, false);
  need(j, "update_filters", &p->update_filters, false);
}

void parse_update_source(const json::Value& j, UpdateSourceConfig* p) {
  get_enum(j, "type", &p->type, parse_update_source_type, "UpdateSourceType_t");
  need(j, "brokers", &p->brokers, false);
  need(j, "receive_buffer_size", &p->receive_buffer_size, false);
  need(j, "poll_timeout_ms", &p->poll_timeout_ms, false);
  need(j, "max_batch_size", &p->max_batch_size, false);
  need(j, "failure_backoff_ms", &p->failure_backoff_ms, false);
  need(j, "max_commit_interval", &p->max_commit_interval, false);
}

void parse_model(const json::Value& j, bool i64, ModelConfig* m) {
  need(j, "model", &m->model_name, true);
  need(j, "network_file", &m->network_file, false);
  need(j, "max_batch_size", &m->max_batch_size, true);
  need(j, "dense_file", &m->dense_file, false);
  need(j, "sparse_files", &m->sparse_files, true);
  need(j, "gpucache", &m->use_gpu_embedding_cache, true);
  need(j, "hit_rate_threshold", &m->hit_rate_threshold, m->use_gpu_embedding_cache);
  need(j, "gpucacheper", &m->cache_size_percentage, m->use_gpu_embedding_cache);
  m->i64_input_key = i64;
  need(j, "num_of_worker_buffer_in_pool", &m->number_of_worker_buffers_in_pool, true);
  need(j, "num_of_refresher_buffer_in_pool", &m->number_of_refresh_buffers_in_pool, false);
  need(j, "cache_refresh_percentage_per_iteration", &m->cache_refresh_percentage_per_iteration,
       false);
  need(j, "deployed_device_list", &m->deployed_devices, true);
  if (m->deployed_devices.empty())
    throw std::invalid_argument("The parameter 'deployed_device_list' must not be empty.");
  m->device_id = m->deployed_devices.back();  // backend.cpp:422
  need(j, "default_value_for_each_table", &m->default_value_for_each_table, true);
  need(j, "maxnum_des_feature_per_sample", &m->maxnum_des_feature_per_sample, false);
  need(j, "maxnum_catfeature_query_per_table_per_sample",
       &m->maxnum_catfeature_query_per_table_per_sample, true);
  need(j, "embedding_vecsize_per_table", &m->embedding_vecsize_per_table, true);
  need(j, "embedding_table_names", &m->embedding_table_names, false);
  need(j, "label_dim", &m->label_dim, false);
  need(j, "slot_num", &m->slot_num, false);
  std::string cache_type;
  need(j, "embedding_cache_type", &cache_type, false);
  cache_type = lower(cache_type);
  // "stochastic" cannot be selected in the reference either (backend.cpp:482,487): -> dynamic
  if (cache_type == "static")
    m->embedding_cache_type = CacheType::Static;
  else if (cache_type == "uvm")
    m->embedding_cache_type = CacheType::UVM;
  else
    m->embedding_cache_type = CacheType::Dynamic;
  need(j, "init_ec", &m->init_ec, false);
  need(j, "fp8_quant", &m->fp8_quant, false);
  need(j, "enable_pagelock", &m->enable_pagelock, false);
  // engine extensions
  need(j, "hpsx_split_lock", &m->hpsx_split_lock, false);
  need(j, "hpsx_request_chunks", &m->hpsx_request_chunks, false);
  need(j, "hpsx_pull_grid_ctas", &m->hpsx_pull_grid_ctas, false);
  need(j, "hpsx_probe", &m->hpsx_probe, false);
  need(j, "hpsx_peer_tier", &m->hpsx_peer_tier, false);
  m->hpsx_probe = lower(m->hpsx_probe);
  if (!m->hpsx_probe.empty() && m->hpsx_probe != "v8" && m->hpsx_probe != "ldg" && m->hpsx_probe != "tma")
    throw std::invalid_argument("The parameter 'hpsx_probe' of model '" + m->model_name + "' must be v8, ldg or tma.");
  need(j, "refresh_delay", &m->refresh_delay, false);
  need(j, "refresh_interval", &m->refresh_interval, false);

  // cross-field consistency the lookup path depends on (HugeCTR checks these when it builds the
  // embedding cache config; the glue itself does not)
  const size_t T = m->sparse_files.size();
  auto same = [&](size_t n, const char* key) {
    if (n != T)
      throw std::invalid_argument(std::string("The parameter '") + key + "' of model '" +
                                  m->model_name + "' must have one entry per sparse file (" +
                                  std::to_string(T) + "), got " + std::to_string(n) + ".");
  };
  same(m->embedding_vecsize_per_table.size(), "embedding_vecsize_per_table");
  same(m->maxnum_catfeature_query_per_table_per_sample.size(),
       "maxnum_catfeature_query_per_table_per_sample");
  same(m->default_value_for_each_table.size(), "default_value_for_each_table");
  if (!m->embedding_table_names.empty())
    same(m->embedding_table_names.size(), "embedding_table_names");
}

}  // namespace

ParseResult parse_ps_config(const json::Value& root, PsConfig* out) {
  ParseResult r;
  try {
    if (!root.is_object()) throw std::invalid_argument("ps.json: top level must be an object");
    *out = PsConfig();
    need(root, "supportlonglong", &out->support_int64_key, true);
    if (const json::Value* j = root.find("volatile_db")) parse_volatile(*j, &out->volatile_db);
    if (const json::Value* j = root.find("persistent_db")) parse_persistent(*j, &out->persistent_db);
    if (const json::Value* j = root.find("update_source"))
      parse_update_source(*j, &out->update_source);
    if (const json::Value* arr = root.find("models")) {
      if (!arr->is_array()) throw std::invalid_argument("The parameter 'models' is not an array.");
      for (const json::Value& jm : arr->items()) {
        ModelConfig m;
        parse_model(jm, out->support_int64_key, &m);
        // a repeated model name replaces the earlier entry (backend.cpp:517-520)
        auto it = std::find_if(out->models.begin(), out->models.end(),
                               [&](const ModelConfig& x) { return x.model_name == m.model_name; });
        if (it != out->models.end())
          *it = std::move(m);
        else
          out->models.push_back(std::move(m));
      }
    }
  } catch (const std::exception& e) {
    r.ok = false;
    r.message = e.what();
  }
  return r;
}

ParseResult parse_ps_config_file(const std::string& path, PsConfig* out) {
  std::ifstream f(path);
  if (!f.is_open()) {
    ParseResult r;
    r.ok = false;
    r.message = "Failed to open Parameter Server Configuration '" + path +
                "', please check whether the file path is correct!";
    return r;
  }
  const std::string text{std::istreambuf_iterator<char>{f}, std::istreambuf_iterator<char>{}};
  try {
    const json::Value root = json::Value::parse(text);
    return parse_ps_config(root, out);
  } catch (const std::exception& e) {
    ParseResult r;
    r.ok = false;
    r.message = std::string("ps.json '") + path + "': " + e.what();
    return r;
  }
}

}  // namespace hpsx
