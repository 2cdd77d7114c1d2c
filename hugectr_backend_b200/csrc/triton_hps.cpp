// libtriton_hps.so — the Triton `hps` backend shell over the hpsx engine (include/triton_hps_backend.h).
//
// State chain (attached with *SetState, freed in the matching *Finalize):
//   Backend  -> ServerState   : ps.json path + the one hpsx_ps + model version map
//                               (reference HPSBackend: hps_backend/src/backend.cpp:59-99)
//   Model    -> ModelState    : validated config.pbtxt + InferenceParams view + instance-group GPUs
//                               (reference ModelState: hps_backend/src/model_state.cpp:66-432)
//   Instance -> InstanceState : one hpsx_session (stream, pinned + device workspaces)
//                               (reference ModelInstanceState: hps_backend/src/model_instance_state.cpp:73-197)
//
// The shell makes no CUDA call of its own: every device operation goes through include/hpsx.h.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <sstream>
#include <condition_variable>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hpsx.h"
#include "../../include/triton_hps_backend.h"
#include "json.hpp"
#include "ps_config.hpp"

namespace {

using hpsx::json::Value;

// ------------------------------------------------------------------------------------------------
// small helpers: logging, errors, time
// ------------------------------------------------------------------------------------------------
template <typename... Args>
std::string cat(const Args&... args) {
  std::ostringstream os;
  (void)std::initializer_list<int>{((os << args), 0)...};
  return os.str();
}

#define HPS_LOG(LEVEL, ...)                                                                        \
  do {                                                                                             \
    TRITONSERVER_Error* le__ =                                                                     \
        TRITONSERVER_LogMessage(TRITONSERVER_LOG_##LEVEL, __FILE__, __LINE__, cat(__VA_ARGS__).c_str()); \
    if (le__ != nullptr) TRITONSERVER_ErrorDelete(le__);                                           \
  } while (0)

#define HPS_ERROR(CODE, ...) TRITONSERVER_ErrorNew(TRITONSERVER_ERROR_##CODE, cat(__VA_ARGS__).c_str())

#define HPS_RETURN_IF_ERROR(X)                \
  do {                                        \
    TRITONSERVER_Error* re__ = (X);           \
    if (re__ != nullptr) return re__;         \
  } while (0)

#define HPS_LOG_IF_ERROR(X, MSG)                                                    \
  do {                                                                              \
    TRITONSERVER_Error* le2__ = (X);                                                \
    if (le2__ != nullptr) {                                                         \
      HPS_LOG(ERROR, MSG, ": ", TRITONSERVER_ErrorMessage(le2__));                  \
      TRITONSERVER_ErrorDelete(le2__);                                              \
    }                                                                               \
  } while (0)

uint64_t now_ns() {
  using namespace std::chrono;
  return static_cast<uint64_t>(duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count());
}

// hpsx status -> TRITONSERVER_Error (message from the engine's thread-local slot)
TRITONSERVER_Error* engine_error(int rc, const std::string& what) {
  TRITONSERVER_Error_Code code = TRITONSERVER_ERROR_INTERNAL;
  switch (rc) {
    case HPSX_ERR_INVALID_ARG: code = TRITONSERVER_ERROR_INVALID_ARG; break;
    case HPSX_ERR_NOT_FOUND: code = TRITONSERVER_ERROR_NOT_FOUND; break;
    case HPSX_ERR_UNSUPPORTED: code = TRITONSERVER_ERROR_UNSUPPORTED; break;
    case HPSX_ERR_IO: code = TRITONSERVER_ERROR_UNAVAILABLE; break;
    default: break;
  }
  return TRITONSERVER_ErrorNew(code, (what + ": " + hpsx_last_error()).c_str());
}

#define HPS_RETURN_IF_ENGINE_ERROR(CALL, WHAT)              \
  do {                                                      \
    const int rc__ = (CALL);                                \
    if (rc__ != HPSX_OK) return engine_error(rc__, WHAT);   \
  } while (0)

std::string shape_to_string(const int64_t* shape, size_t dims) {
  std::string s = "[";
  for (size_t i = 0; i < dims; ++i) s += (i ? "," : "") + std::to_string(shape[i]);
  return s + "]";
}

// Parses a Triton message (backend config, model config) into the in-tree JSON DOM.
TRITONSERVER_Error* message_to_json(TRITONSERVER_Message* msg, Value* out, std::string* text) {
  const char* base = nullptr;
  size_t size = 0;
  HPS_RETURN_IF_ERROR(TRITONSERVER_MessageSerializeToJson(msg, &base, &size));
  if (text != nullptr) text->assign(base ? base : "", size);
  try {
    *out = Value::parse(base ? base : "", size);
  } catch (const std::exception& e) {
    return HPS_ERROR(INVALID_ARG, "failed to parse JSON message: ", e.what());
  }
  return nullptr;
}

// "dims": [-1] — numbers, or strings holding numbers (protobuf JSON renders int64 as strings).
bool parse_dims(const Value& obj, const char* key, std::vector<int64_t>* out) {
  const Value* a = obj.find(key);
  if (a == nullptr || !a->is_array()) return false;
  out->clear();
  for (const Value& v : a->items()) {
    if (v.is_number()) {
      out->push_back(static_cast<int64_t>(v.as_double()));
    } else if (v.is_string()) {
      try {
        out->push_back(std::stoll(v.as_string()));
      } catch (const std::exception&) {
        return false;
      }
    } else {
      return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// Backend state
// ------------------------------------------------------------------------------------------------
struct ServerState {
  std::string ps_json_path;
  hpsx_ps* ps = nullptr;
  std::mutex mu;                             // guards `versions` (reference: backend.cpp:85,96)
  std::map<std::string, uint64_t> versions;  // model -> version last initialised

  ~ServerState() {
    if (ps != nullptr) hpsx_ps_destroy(ps);
  }
  uint64_t version_of(const std::string& model) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = versions.find(model);
    return it == versions.end() ? 0 : it->second;
  }
  void set_version(const std::string& model, uint64_t v) {
    std::lock_guard<std::mutex> lk(mu);
    versions[model] = v;
  }
};

// ------------------------------------------------------------------------------------------------
// Model state
// ------------------------------------------------------------------------------------------------
struct ModelState {
  TRITONBACKEND_Model* triton_model = nullptr;
  ServerState* server = nullptr;
  std::string name;
  uint64_t version = 0;
  Value config;                // config.pbtxt as JSON
  hpsx_model_params params{};  // view into the parameter server's storage
  std::vector<int> gpus;       // instance_group gpus (CPU mode: {0}, like model_state.cpp:297)
  size_t cat_num = 0;          // sum of maxnum_catfeature_query_per_table_per_sample (model_state.cpp:337-345)
  size_t max_batch_size = 0;   // ps.json's value wins over config.pbtxt (model_state.cpp:366)
  float refresh_interval = 0.f, refresh_delay = 0.f;
  bool freeze_sparse = false;
  std::string output_name;
  // Opt-in extension (north-star stage a8): config.pbtxt parameters `hps_pooling` = "sum" | "mean" and
  // `hps_pooling_hotness` = "h0,h1,..." (keys per slot of every table, default 1).  OUTPUT0 then holds the
  // slot-wise reduced vectors: sum_t (n_t / h_t) * d_t floats.  Without them the reference's un-pooled
  // contract applies.
  int pooling = -1;  // -1 off, else hpsx_combiner
  std::vector<size_t> pooling_hotness;
  // Opt-in extension: config.pbtxt parameter `hps_report_stats` = "1" adds the response parameters CacheHits /
  // CacheMisses (totals of the Execute call the response belongs to) — how a client, or bench.py, learns the hit
  // rate of a deployed model.  Costs one small device-to-host read per call.
  bool report_stats = false;

  bool gpucache() const { return params.use_gpu_embedding_cache != 0; }
  size_t num_tables() const { return params.num_tables; }

  TRITONSERVER_Error* validate();         // ~ ValidateModelConfig   model_state.cpp:180-261
  TRITONSERVER_Error* parse();            // ~ ParseModelConfig      model_state.cpp:263-371
  TRITONSERVER_Error* ensure_caches();    // ~ Create_EmbeddingCache model_state.cpp:373-432

  // Cache refresh (model_state.cpp:124-178, 413-427; include/timer.hpp): a periodic thread when
  // refresh_interval > 0, and one asynchronous database reload + refresh when a new version of an already
  // served model is initialised.  Both run beside lookups and are joined before the state is destroyed.
  uint64_t previous_version = 0;  // version the parameter server last initialised for this model (0: none)
  std::mutex refresh_mu;
  std::condition_variable refresh_cv;
  bool stopping = false;
  std::thread periodic;
  std::vector<std::thread> one_shot;
  void refresh_all_devices();           // ~ Refresh_Embedding_Cache
  void reload_and_refresh(int device);  // ~ EmbeddingCacheRefresh
  void start_refresh_threads();
  ~ModelState();
};

TRITONSERVER_Error* ModelState::validate() {
  const Value* inputs = config.find("input");
  if (inputs == nullptr || !inputs->is_array())
    return HPS_ERROR(INVALID_ARG, "model configuration of '", name, "' has no 'input' array");
  if (inputs->size() != 2) return HPS_ERROR(INVALID_ARG, "expect 2 input, got ", inputs->size());
  for (size_t i = 0; i < 2; ++i) {
    const Value& in = inputs->at(i);
    std::string iname, dtype;
    if (!in.is_object() || !hpsx::json_get(in, "name", &iname))
      return HPS_ERROR(INVALID_ARG, "input ", i, " of model '", name, "' has no name");
    if (iname != "KEYS" && iname != "NUMKEYS")
      return HPS_ERROR(INVALID_ARG, "expected input name as KEYS and NUMKEYS, but got ", iname);
    if (!hpsx::json_get(in, "data_type", &dtype))
      return HPS_ERROR(INVALID_ARG, "input ", iname, " has no data_type");
    if (iname == "KEYS" && dtype != "TYPE_INT64")
      return HPS_ERROR(INVALID_ARG, "expected KEYS input datatype as TYPE_INT64, got ", dtype);
    if (iname == "NUMKEYS" && dtype != "TYPE_INT32")
      return HPS_ERROR(INVALID_ARG, "expected NUMKEYS input datatype as TYPE_INT32, got ", dtype);
    std::vector<int64_t> dims;
    if (!parse_dims(in, "dims", &dims) || dims.empty())
      return HPS_ERROR(INVALID_ARG, "input ", iname, " has no dims");
    if (dims[0] != -1)
      return HPS_ERROR(INVALID_ARG, "expected input shape equal -1, got ",
                       shape_to_string(dims.data(), dims.size()));
  }
  if (inputs->at(0).find("name")->as_string() == inputs->at(1).find("name")->as_string())
    return HPS_ERROR(INVALID_ARG, "expected one KEYS and one NUMKEYS input, got two ",
                     inputs->at(0).find("name")->as_string());

  const Value* outputs = config.find("output");
  if (outputs == nullptr || !outputs->is_array())
    return HPS_ERROR(INVALID_ARG, "model configuration of '", name, "' has no 'output' array");
  if (outputs->size() != 1) return HPS_ERROR(INVALID_ARG, "expect 1 output, got ", outputs->size());
  const Value& out = outputs->at(0);
  std::string dtype;
  if (!out.is_object() || !hpsx::json_get(out, "data_type", &dtype))
    return HPS_ERROR(INVALID_ARG, "output of model '", name, "' has no data_type");
  if (dtype != "TYPE_FP32")
    return HPS_ERROR(INVALID_ARG, "expected  output datatype as TYPE_FP32, got ", dtype);
  std::vector<int64_t> dims;
  if (!parse_dims(out, "dims", &dims) || dims.empty())
    return HPS_ERROR(INVALID_ARG, "output of model '", name, "' has no dims");
  if (dims[0] != -1)
    return HPS_ERROR(INVALID_ARG, "expected  output shape equal -1, got ",
                     shape_to_string(dims.data(), dims.size()));
  hpsx::json_get(out, "name", &output_name);
  return nullptr;
}

TRITONSERVER_Error* ModelState::parse() {
  const Value* groups = config.find("instance_group");
  if (groups == nullptr || !groups->is_array() || groups->size() == 0)
    return HPS_ERROR(INVALID_ARG, "expect at least one instance in instance group , got ",
                     groups != nullptr ? groups->size() : 0);
  gpus.clear();
  for (const Value& g : groups->items()) {
    std::string kind;
    if (!g.is_object() || !hpsx::json_get(g, "kind", &kind))
      return HPS_ERROR(INVALID_ARG, "instance_group entry of model '", name, "' has no kind");
    if (gpucache()) {
      if (kind != "KIND_GPU")
        return HPS_ERROR(INVALID_ARG, "expect GPU kind instance in instance group , got ", kind);
      std::vector<int64_t> list;
      if (!parse_dims(g, "gpus", &list))
        return HPS_ERROR(INVALID_ARG, "instance_group entry of model '", name, "' has no gpus list");
      for (int64_t id : list)
        if (std::find(gpus.begin(), gpus.end(), static_cast<int>(id)) == gpus.end())
          gpus.push_back(static_cast<int>(id));
    } else if (gpus.empty()) {
      gpus.push_back(0);
    }
    int64_t count = 1;
    try {
      hpsx::json_get(g, "count", &count);
    } catch (const std::exception& e) {
      return HPS_ERROR(INVALID_ARG, "instance_group count of model '", name, "': ", e.what());
    }
    if (count > static_cast<int64_t>(params.number_of_worker_buffers_in_pool))
      return HPS_ERROR(INVALID_ARG,
                       "expect the number of instance(in instance_group) not larger than "
                       "num_of_worker_buffer_in_pool that is configured in the Parameter Server json file (",
                       params.number_of_worker_buffers_in_pool, "), got ", count);
  }

  // config.pbtxt `parameters { key: "refresh_interval" value { string_value: "…" } }`
  if (const Value* p = config.find("parameters"); p != nullptr && p->is_object()) {
    try {
      if (const Value* v = p->find("refresh_interval"); v && v->is_object())
        hpsx::json_get(*v, "string_value", &refresh_interval);
      if (const Value* v = p->find("refresh_delay"); v && v->is_object())
        hpsx::json_get(*v, "string_value", &refresh_delay);
      if (const Value* v = p->find("freeze_sparse"); v && v->is_object())
        hpsx::json_get(*v, "string_value", &freeze_sparse);
      if (const Value* v = p->find("hps_report_stats"); v && v->is_object())
        hpsx::json_get(*v, "string_value", &report_stats);
      std::string pool_mode, pool_hot;
      if (const Value* v = p->find("hps_pooling"); v && v->is_object()) hpsx::json_get(*v, "string_value", &pool_mode);
      if (const Value* v = p->find("hps_pooling_hotness"); v && v->is_object())
        hpsx::json_get(*v, "string_value", &pool_hot);
      if (!pool_mode.empty() && pool_mode != "none") {
        if (pool_mode == "sum")
          pooling = HPSX_COMBINER_SUM;
        else if (pool_mode == "mean")
          pooling = HPSX_COMBINER_MEAN;
        else
          return HPS_ERROR(INVALID_ARG, "model '", name, "': hps_pooling must be sum, mean or none, got ", pool_mode);
        pooling_hotness.assign(params.num_tables, 1);
        size_t t = 0;
        std::stringstream ss(pool_hot);
        for (std::string tok; std::getline(ss, tok, ',');) {
          if (t >= params.num_tables)
            return HPS_ERROR(INVALID_ARG, "model '", name, "': hps_pooling_hotness lists more entries than tables");
          const long long h = std::stoll(tok);
          if (h <= 0) return HPS_ERROR(INVALID_ARG, "model '", name, "': hps_pooling_hotness entries must be > 0");
          pooling_hotness[t++] = static_cast<size_t>(h);
        }
      }
    } catch (const std::exception& e) {
      return HPS_ERROR(INVALID_ARG, "model '", name, "' parameters: ", e.what());
    }
  }

  cat_num = 0;
  for (size_t t = 0; t < params.num_tables; ++t)
    cat_num += params.maxnum_catfeature_query_per_table_per_sample[t];
  if (cat_num == 0) return HPS_ERROR(INVALID_ARG, "expected at least one categorical feature, got ", cat_num);

  int64_t pbtxt_max_batch = 0;
  try {
    hpsx::json_get(config, "max_batch_size", &pbtxt_max_batch);
  } catch (const std::exception& e) {
    return HPS_ERROR(INVALID_ARG, "max_batch_size of model '", name, "': ", e.what());
  }
  if (pbtxt_max_batch < 0)
    return HPS_ERROR(INVALID_ARG,
                     "expected max_batch_size should greater than or equal to 0 (the configuration should be "
                     "consistent in Parameter Server json file and config.pbtxt file), got ",
                     pbtxt_max_batch);
  max_batch_size = params.max_batch_size;
  HPS_LOG(INFO, "model ", name, ": max_batch_size ", max_batch_size, " (ps.json), ", cat_num,
          " keys per sample, ", params.num_tables, " tables, gpucache ", gpucache() ? "on" : "off");
  return nullptr;
}

TRITONSERVER_Error* ModelState::ensure_caches() {
  if (!gpucache()) {
    start_refresh_threads();  // CPU models still reload their database when a new version arrives
    return nullptr;
  }
  for (int dev : gpus) {
    const int* b = params.deployed_devices;
    const int* e = b + params.num_deployed_devices;
    if (std::find(b, e, dev) == e)
      return HPS_ERROR(INVALID_ARG, "Please confirm that device ", dev,
                       " is added to 'deployed_device_list' in the ps configuration file");
  }
  // creates the caches on every deployed device that has none yet (no-op when init_ec already did)
  HPS_RETURN_IF_ENGINE_ERROR(hpsx_ps_create_embedding_cache_per_model(server->ps, name.c_str()),
                             "creating the embedding cache of model " + name);
  for (int dev : gpus) {
    hpsx_cache* c = nullptr;
    HPS_RETURN_IF_ENGINE_ERROR(hpsx_ps_get_embedding_cache(server->ps, name.c_str(), dev, &c),
                               "fetching the embedding cache of model " + name);
    HPS_LOG(INFO, "******Embedding cache of model ", name, " ready on device ", dev);
  }
  start_refresh_threads();
  return nullptr;
}

void ModelState::refresh_all_devices() {
  const uint64_t t0 = now_ns();
  for (int dev : gpus) {
    if (!gpucache()) continue;
    HPS_LOG(INFO, "The model ", name, " is periodically refreshing the embedding cache asynchronously on device ", dev);
    size_t rows = 0;
    const int rc = hpsx_ps_refresh_embedding_cache(server->ps, name.c_str(), dev, &rows);
    if (rc != HPSX_OK)
      HPS_LOG(ERROR, "refreshing the embedding cache of model ", name, " on device ", dev, " failed: ", hpsx_last_error());
    else
      HPS_LOG(INFO, "The model ", name, " has refreshed the embedding cache asynchronously on device ", dev, " (", rows,
              " rows)");
  }
  HPS_LOG(INFO, "Refresh embedding table execution time is ", (now_ns() - t0) / 1000000, " ms");
}

void ModelState::reload_and_refresh(int device) {
  HPS_LOG(INFO, "The model ", name, " is refreshing the embedding cache asynchronously on device ", device, ".");
  if (!freeze_sparse) {
    const int rc = hpsx_ps_update_database_per_model(server->ps, name.c_str());
    if (rc != HPSX_OK) HPS_LOG(ERROR, "updating the database of model ", name, " failed: ", hpsx_last_error());
  }
  if (gpucache()) {
    const int rc = hpsx_ps_refresh_embedding_cache(server->ps, name.c_str(), device, nullptr);
    if (rc != HPSX_OK)
      HPS_LOG(ERROR, "refreshing the embedding cache of model ", name, " on device ", device, " failed: ", hpsx_last_error());
  }
  HPS_LOG(INFO, "The model ", name, " has completed the asynchronous refresh of the embedding cache on device ", device, ".");
}

void ModelState::start_refresh_threads() {
  // a new version of a model the parameter server already serves: reload its sparse files (unless
  // freeze_sparse) and refresh the caches, once, in the background (model_state.cpp:413-420)
  if (previous_version > 0 && previous_version != version)
    for (int dev : gpus) one_shot.emplace_back([this, dev] { reload_and_refresh(dev); });
  if (refresh_interval > 1e-6f) {
    HPS_LOG(INFO, "model ", name, ": refreshing the embedding cache every ", refresh_interval, " s");
    periodic = std::thread([this] {
      const auto period = std::chrono::duration<double>(std::max(0.01, static_cast<double>(refresh_interval)));
      std::unique_lock<std::mutex> lk(refresh_mu);
      while (!stopping) {
        if (refresh_cv.wait_for(lk, period, [this] { return stopping; })) break;
        lk.unlock();
        refresh_all_devices();
        lk.lock();
      }
    });
  }
}

ModelState::~ModelState() {
  {
    std::lock_guard<std::mutex> lk(refresh_mu);
    stopping = true;
  }
  refresh_cv.notify_all();
  if (periodic.joinable()) periodic.join();
  for (std::thread& t : one_shot)
    if (t.joinable()) t.join();
}

// ------------------------------------------------------------------------------------------------
// Instance state
// ------------------------------------------------------------------------------------------------
struct InstanceState {
  TRITONBACKEND_ModelInstance* triton_instance = nullptr;
  ModelState* model = nullptr;
  std::string name;
  int device = 0;
  hpsx_session* session = nullptr;
  std::vector<int64_t> key_staging;  // only for inputs Triton delivers in several buffers

  ~InstanceState() {
    if (session != nullptr) hpsx_session_destroy(session);
  }
};

// Sends `err` as the (final) response of request r and forgets the response: the reference's
// GUARDED_RESPOND_IF_ERROR (include/hps_buffer.hpp:62-76).
void respond_error(std::vector<TRITONBACKEND_Response*>& responses, uint32_t r, TRITONSERVER_Error* err) {
  if (responses[r] != nullptr) {
    HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSend(responses[r], TRITONSERVER_RESPONSE_COMPLETE_FINAL, err),
                     "failed to send error response");
    responses[r] = nullptr;
  }
  TRITONSERVER_ErrorDelete(err);
}

struct InputView {
  const void* data = nullptr;  // contiguous bytes (Triton's buffer or the instance staging)
  uint64_t bytes = 0;
  bool on_device = false;
};

// Collects an input into one contiguous range.  A single buffer is used in place whatever its
// memory type; several buffers are concatenated into host staging (the reference copies every buffer
// to offset 0, src/hps.cc:586-597 — SURVEY.md Appendix B.3).
TRITONSERVER_Error* gather_input(InstanceState* inst, TRITONBACKEND_Input* input, uint32_t buffer_count,
                                 uint64_t total_bytes, std::vector<int64_t>* staging, InputView* view) {
  if (buffer_count == 0 || total_bytes == 0) {
    *view = InputView{};
    return nullptr;
  }
  if (buffer_count == 1) {
    const void* buf = nullptr;
    uint64_t bytes = 0;
    TRITONSERVER_MemoryType mt = TRITONSERVER_MEMORY_CPU_PINNED;
    int64_t mt_id = 0;
    HPS_RETURN_IF_ERROR(TRITONBACKEND_InputBuffer(input, 0, &buf, &bytes, &mt, &mt_id));
    if (mt == TRITONSERVER_MEMORY_GPU && inst->session != nullptr && mt_id != inst->device && inst->model->gpucache())
      return HPS_ERROR(UNSUPPORTED, "input buffer lives on GPU ", mt_id, " but the instance runs on GPU ",
                       inst->device);
    view->data = buf;
    view->bytes = bytes;
    view->on_device = mt == TRITONSERVER_MEMORY_GPU;
    return nullptr;
  }
  staging->resize((total_bytes + 7) / 8);
  char* dst = reinterpret_cast<char*>(staging->data());
  uint64_t off = 0;
  for (uint32_t b = 0; b < buffer_count; ++b) {
    const void* buf = nullptr;
    uint64_t bytes = 0;
    TRITONSERVER_MemoryType mt = TRITONSERVER_MEMORY_CPU_PINNED;
    int64_t mt_id = 0;
    HPS_RETURN_IF_ERROR(TRITONBACKEND_InputBuffer(input, b, &buf, &bytes, &mt, &mt_id));
    if (off + bytes > total_bytes) return HPS_ERROR(INVALID_ARG, "input buffers exceed the input's byte size");
    if (mt == TRITONSERVER_MEMORY_GPU) {
      HPS_RETURN_IF_ENGINE_ERROR(hpsx_copy_to_host(static_cast<int>(mt_id), dst + off, buf, bytes),
                                 "copying an input buffer to the host");
    } else {
      std::memcpy(dst + off, buf, bytes);
    }
    off += bytes;
  }
  view->data = dst;
  view->bytes = off;
  view->on_device = false;
  return nullptr;
}

// Everything one request needs for its lookup, resolved up front so that the requests of one Execute call can be
// served in one engine pass (the reference walks them one by one, src/hps.cc:392-406).
struct RequestPlan {
  bool has_output = false;  // false: nothing was requested (src/hps.cc:549-551), respond without a lookup
  bool keys_on_device = false, out_on_device = false;
  int64_t num_samples = 0;
  uint64_t num_keys = 0;
  std::vector<const void*> keys_pt;
  std::vector<float*> out_pt;
  std::vector<size_t> n_per_table;
  std::vector<int64_t> key_staging;  // only for inputs Triton delivers in several buffers
};

// Validates one request, sizes and obtains its output buffer and splits keys/output per table.  A non-null
// error is turned into an error response by the caller.
TRITONSERVER_Error* prepare_request(InstanceState* inst, TRITONBACKEND_Request* request,
                                    TRITONBACKEND_Response* response, RequestPlan* plan) {
  ModelState* ms = inst->model;
  int64_t* num_samples = &plan->num_samples;
  uint32_t input_count = 0, requested_output_count = 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestInputCount(request, &input_count));
  HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestOutputCount(request, &requested_output_count));
  if (input_count != 2)
    return HPS_ERROR(INVALID_ARG, "expected 2 inputs (KEYS, NUMKEYS) in request, got ", input_count);
  for (uint32_t i = 0; i < 2; ++i) {
    const char* in_name = nullptr;
    HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestInputName(request, i, &in_name));
    if (in_name == nullptr || (std::strcmp(in_name, "KEYS") != 0 && std::strcmp(in_name, "NUMKEYS") != 0))
      return HPS_ERROR(INVALID_ARG, "expected input name as KEYS and NUMKEYS in request, but got ",
                       in_name ? in_name : "(null)");
  }
  TRITONBACKEND_Input *keys_in = nullptr, *numkeys_in = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestInput(request, "KEYS", &keys_in));
  HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestInput(request, "NUMKEYS", &numkeys_in));

  TRITONSERVER_DataType keys_dt, numkeys_dt;
  const int64_t *keys_shape = nullptr, *numkeys_shape = nullptr;
  uint32_t keys_dims = 0, numkeys_dims = 0, keys_buffers = 0, numkeys_buffers = 0;
  uint64_t keys_bytes = 0, numkeys_bytes = 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_InputProperties(keys_in, nullptr, &keys_dt, &keys_shape, &keys_dims,
                                                    &keys_bytes, &keys_buffers));
  HPS_RETURN_IF_ERROR(TRITONBACKEND_InputProperties(numkeys_in, nullptr, &numkeys_dt, &numkeys_shape,
                                                    &numkeys_dims, &numkeys_bytes, &numkeys_buffers));
  if (keys_dt != TRITONSERVER_TYPE_INT64)
    return HPS_ERROR(INVALID_ARG, "expected KEYS datatype TYPE_INT64, got ", TRITONSERVER_DataTypeString(keys_dt));
  if (numkeys_dt != TRITONSERVER_TYPE_INT32)
    return HPS_ERROR(INVALID_ARG, "expected NUMKEYS datatype TYPE_INT32, got ",
                     TRITONSERVER_DataTypeString(numkeys_dt));

  const uint64_t num_keys = keys_bytes / sizeof(int64_t);
  plan->num_keys = num_keys;
  *num_samples = static_cast<int64_t>(num_keys / ms->cat_num);
  if (requested_output_count == 0) return nullptr;  // nothing to produce (src/hps.cc:549-551)
  const char* out_name = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_RequestOutputName(request, 0, &out_name));

  if (static_cast<uint64_t>(*num_samples) > ms->max_batch_size)
    return HPS_ERROR(UNSUPPORTED, "The number of Input samples greater than max batch size");

  // NUMKEYS: int32 [1, T] (src/hps.cc:616-618 reads shape[1] entries); any shape with <= T elements is
  // accepted, missing trailing tables count as 0 keys.
  const size_t T = ms->num_tables();
  const uint64_t numkeys_count = numkeys_bytes / sizeof(int32_t);
  if (numkeys_count == 0 || numkeys_count > T)
    return HPS_ERROR(INVALID_ARG, "NUMKEYS ", shape_to_string(numkeys_shape, numkeys_dims), " must hold between 1 and ",
                     T, " entries (one per embedding table of model ", ms->name, ")");
  std::vector<int64_t> numkeys_staging;
  InputView nk_view;
  HPS_RETURN_IF_ERROR(gather_input(inst, numkeys_in, numkeys_buffers, numkeys_bytes, &numkeys_staging, &nk_view));
  std::vector<int32_t> numkeys(numkeys_count);
  if (nk_view.on_device) {
    HPS_RETURN_IF_ENGINE_ERROR(hpsx_copy_to_host(inst->device, numkeys.data(), nk_view.data, numkeys_bytes),
                               "copying NUMKEYS to the host");
  } else {
    std::memcpy(numkeys.data(), nk_view.data, numkeys_count * sizeof(int32_t));
  }
  std::vector<size_t> n_per_table(T, 0);
  uint64_t key_sum = 0;
  int64_t out_floats = 0;  // 64-bit: one Criteo-shape response is 218 M floats
  for (size_t t = 0; t < numkeys_count; ++t) {
    if (numkeys[t] < 0) return HPS_ERROR(INVALID_ARG, "NUMKEYS[", t, "] is negative (", numkeys[t], ")");
    n_per_table[t] = static_cast<size_t>(numkeys[t]);
    key_sum += n_per_table[t];
    size_t out_rows = n_per_table[t];
    if (ms->pooling >= 0) {
      if (n_per_table[t] % ms->pooling_hotness[t] != 0)
        return HPS_ERROR(INVALID_ARG, "NUMKEYS[", t, "] = ", numkeys[t], " is not a multiple of the table's pooling hotness ",
                         ms->pooling_hotness[t]);
      out_rows = n_per_table[t] / ms->pooling_hotness[t];
    }
    out_floats += static_cast<int64_t>(out_rows) * static_cast<int64_t>(ms->params.embedding_vecsize_per_table[t]);
    if (n_per_table[t] > ms->max_batch_size * ms->params.maxnum_catfeature_query_per_table_per_sample[t])
      return HPS_ERROR(UNSUPPORTED, "NUMKEYS[", t, "] = ", numkeys[t], " exceeds max_batch_size * "
                       "maxnum_catfeature_query_per_table_per_sample = ",
                       ms->max_batch_size * ms->params.maxnum_catfeature_query_per_table_per_sample[t]);
  }
  if (key_sum != num_keys)
    return HPS_ERROR(INVALID_ARG, "NUMKEYS sums to ", key_sum, " but KEYS holds ", num_keys, " keys");

  InputView key_view;
  HPS_RETURN_IF_ERROR(gather_input(inst, keys_in, keys_buffers, keys_bytes, &plan->key_staging, &key_view));
  if (key_view.on_device && !ms->gpucache())
    return HPS_ERROR(UNSUPPORTED, "KEYS arrived in GPU memory but model ", ms->name, " runs without a GPU cache");

  // output tensor: FP32 [sum_t n_t * d_t]  (src/hps.cc:620-630)
  TRITONBACKEND_Output* output = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ResponseOutput(response, &output, out_name, TRITONSERVER_TYPE_FP32, &out_floats, 1));
  void* out_buf = nullptr;
  TRITONSERVER_MemoryType out_mt = ms->gpucache() ? TRITONSERVER_MEMORY_GPU : TRITONSERVER_MEMORY_CPU;
  int64_t out_mt_id = ms->gpucache() ? inst->device : 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_OutputBuffer(output, &out_buf, static_cast<uint64_t>(out_floats) * sizeof(float),
                                                 &out_mt, &out_mt_id));
  const bool out_on_device = out_mt == TRITONSERVER_MEMORY_GPU;
  if (out_on_device && !ms->gpucache())
    return HPS_ERROR(UNSUPPORTED, "output buffer was placed in GPU memory but model ", ms->name,
                     " runs without a GPU cache");
  if (out_on_device && out_mt_id != inst->device)
    return HPS_ERROR(UNSUPPORTED, "output buffer lives on GPU ", out_mt_id, " but the instance runs on GPU ",
                     inst->device);
  if (out_floats > 0 && out_buf == nullptr) return HPS_ERROR(INTERNAL, "Triton returned a null output buffer");

  // per-table pointers by prefix sums (src/model_instance_state.cpp:180-193); the kernels write the
  // rows directly into Triton's buffer — no result buffer, no D2D hand-off (src/hps.cc:676-680)
  plan->keys_pt.assign(T, nullptr);
  plan->out_pt.assign(T, nullptr);
  const int64_t* kbase = static_cast<const int64_t*>(key_view.data);
  float* obase = static_cast<float*>(out_buf);
  size_t koff = 0, ooff = 0;
  for (size_t t = 0; t < T; ++t) {
    plan->keys_pt[t] = kbase + koff;
    plan->out_pt[t] = obase + ooff;
    koff += n_per_table[t];
    ooff += (ms->pooling >= 0 ? n_per_table[t] / ms->pooling_hotness[t] : n_per_table[t]) *
            ms->params.embedding_vecsize_per_table[t];
  }
  plan->n_per_table = std::move(n_per_table);
  plan->keys_on_device = key_view.on_device;
  plan->out_on_device = out_on_device;
  plan->has_output = true;
  return nullptr;
}

// The lookup of one prepared request.
TRITONSERVER_Error* run_single(InstanceState* inst, const RequestPlan& plan) {
  ModelState* ms = inst->model;
  const size_t T = ms->num_tables();
  if (ms->pooling >= 0) {
    // fused slot-wise gather + reduce, one table at a time
    for (size_t t = 0; t < T; ++t) {
      if (plan.n_per_table[t] == 0) continue;
      const size_t h = ms->pooling_hotness[t];
      const int prc = hpsx_session_lookup_pooled_ex(
          inst->session, t, static_cast<const int64_t*>(plan.keys_pt[t]), plan.keys_on_device ? HPSX_MEM_DEVICE : HPSX_MEM_HOST,
          plan.n_per_table[t] / h, h, ms->pooling, plan.out_pt[t], plan.out_on_device ? HPSX_MEM_DEVICE : HPSX_MEM_HOST);
      if (prc != HPSX_OK) return engine_error(prc, "pooled embedding lookup of model " + ms->name);
    }
    return nullptr;
  }
  const int rc = hpsx_session_lookup_ex(inst->session, plan.keys_pt.data(), plan.keys_on_device ? HPSX_MEM_DEVICE : HPSX_MEM_HOST,
                                        plan.out_pt.data(), plan.out_on_device ? HPSX_MEM_DEVICE : HPSX_MEM_HOST,
                                        plan.n_per_table.data(), T);
  if (rc != HPSX_OK) return engine_error(rc, "embedding lookup of model " + ms->name);
  return nullptr;
}

// Requests [first, last) of `plans` in ONE engine pass (all are un-pooled GPU lookups into device buffers).
TRITONSERVER_Error* run_batch(InstanceState* inst, const std::vector<RequestPlan>& plans, const std::vector<uint32_t>& group) {
  ModelState* ms = inst->model;
  const size_t T = ms->num_tables();
  std::vector<const void*> keys;
  std::vector<float*> out;
  std::vector<size_t> n;
  for (uint32_t r : group) {
    keys.insert(keys.end(), plans[r].keys_pt.begin(), plans[r].keys_pt.end());
    out.insert(out.end(), plans[r].out_pt.begin(), plans[r].out_pt.end());
    n.insert(n.end(), plans[r].n_per_table.begin(), plans[r].n_per_table.end());
  }
  (void)T;
  const int rc = hpsx_session_lookup_batch(inst->session, group.size(), keys.data(),
                                           plans[group[0]].keys_on_device ? HPSX_MEM_DEVICE : HPSX_MEM_HOST, out.data(),
                                           HPSX_MEM_DEVICE, n.data());
  if (rc != HPSX_OK) return engine_error(rc, "batched embedding lookup of model " + ms->name);
  return nullptr;
}

}  // namespace

// ================================================================================================
// exported C ABI
// ================================================================================================
extern "C" {

TRITONSERVER_Error* TRITONBACKEND_Initialize(TRITONBACKEND_Backend* backend) {
  const char* name = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendName(backend, &name));
  HPS_LOG(INFO, "TRITONBACKEND_Initialize: ", name ? name : "(null)");

  uint32_t major = 0, minor = 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ApiVersion(&major, &minor));
  HPS_LOG(INFO, "Triton TRITONBACKEND API version: ", major, ".", minor, "; '", name ? name : "",
          "' TRITONBACKEND API version: ", TRITONBACKEND_API_VERSION_MAJOR, ".", TRITONBACKEND_API_VERSION_MINOR);
  if (major != TRITONBACKEND_API_VERSION_MAJOR || minor < TRITONBACKEND_API_VERSION_MINOR)
    return HPS_ERROR(UNSUPPORTED, "Triton backend API version does not support this backend");

  TRITONSERVER_Message* cfg_msg = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendConfig(backend, &cfg_msg));
  TRITONBACKEND_ArtifactType artifact_type;
  const char* location = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendArtifacts(backend, &artifact_type, &location));
  HPS_LOG(INFO, "The Hierarchical Parameter Server Backend Repository location: ", location ? location : "");

  // {"cmdline":{"ps":"/path/ps.json", ...}}  (tritonserver --backend-config=hps,ps=<file>)
  Value cfg;
  std::string cfg_text;
  HPS_RETURN_IF_ERROR(message_to_json(cfg_msg, &cfg, &cfg_text));
  HPS_LOG(INFO, "The HPS configuration: ", cfg_text);
  std::string ps_path;
  if (const Value* cmdline = cfg.find("cmdline"); cmdline != nullptr && cmdline->is_object()) {
    if (const Value* ps = cmdline->find("ps"); ps != nullptr && ps->is_string()) ps_path = ps->as_string();
  }
  if (ps_path.empty())
    return HPS_ERROR(INVALID_ARG,
                     "the path of the Parameter Server json configuration is missing: start tritonserver with "
                     "--backend-config=hps,ps=<ps.json>");

  std::unique_ptr<ServerState> state(new ServerState());
  state->ps_json_path = ps_path;
  HPS_LOG(INFO, "*****The HierarchicalParameterServer is creating... *****");
  HPS_RETURN_IF_ENGINE_ERROR(hpsx_ps_create_from_json(ps_path.c_str(), &state->ps),
                             "creating the Hierarchical Parameter Server from " + ps_path);
  HPS_LOG(INFO, "*****The HierarchicalParameterServer has been created successfully! *****");
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendSetState(backend, state.get()));
  state.release();
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_Finalize(TRITONBACKEND_Backend* backend) {
  void* vstate = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendState(backend, &vstate));
  HPS_LOG(INFO, "TRITONBACKEND_Backend Finalize: HPSBackend");
  delete static_cast<ServerState*>(vstate);
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_ModelInitialize(TRITONBACKEND_Model* model) {
  const char* name = nullptr;
  uint64_t version = 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelName(model, &name));
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelVersion(model, &version));
  HPS_LOG(INFO, "TRITONBACKEND_ModelInitialize: ", name, " (version ", version, ")");
  TRITONBACKEND_ArtifactType artifact_type;
  const char* location = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelRepository(model, &artifact_type, &location));
  HPS_LOG(INFO, "Repository location: ", location ? location : "");

  TRITONBACKEND_Backend* backend = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelBackend(model, &backend));
  void* vbackend = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_BackendState(backend, &vbackend));
  ServerState* server = static_cast<ServerState*>(vbackend);
  if (server == nullptr || server->ps == nullptr)
    return HPS_ERROR(INTERNAL, "the hps backend has no parameter server (TRITONBACKEND_Initialize failed?)");

  // online deployment: a model that was added to ps.json after the server started (src/hps.cc:207-219)
  if (!hpsx_ps_has_model(server->ps, name)) {
    HPS_LOG(INFO, "Parsing the latest Parameter Server json config file for deploying model ", name, " online");
    size_t added = 0;
    HPS_RETURN_IF_ENGINE_ERROR(hpsx_ps_sync_models_from_json(server->ps, server->ps_json_path.c_str(), &added),
                               "re-reading " + server->ps_json_path);
    if (!hpsx_ps_has_model(server->ps, name))
      return HPS_ERROR(INVALID_ARG, "Please make sure that the configuration of model ", name,
                       " has been added to the Parameter Server json configuration file ", server->ps_json_path);
  }

  std::unique_ptr<ModelState> ms(new ModelState());
  ms->triton_model = model;
  ms->server = server;
  ms->name = name;
  ms->version = version;
  ms->previous_version = server->version_of(name);
  HPS_RETURN_IF_ENGINE_ERROR(hpsx_ps_get_model_params(server->ps, name, &ms->params),
                             std::string("reading the parameters of model ") + name);
  TRITONSERVER_Message* cfg_msg = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelConfig(model, 1 /* config_version */, &cfg_msg));
  TRITONSERVER_Error* perr = message_to_json(cfg_msg, &ms->config, nullptr);
  HPS_LOG_IF_ERROR(TRITONSERVER_MessageDelete(cfg_msg), "failed to delete the model configuration message");
  HPS_RETURN_IF_ERROR(perr);

  HPS_RETURN_IF_ERROR(ms->validate());
  HPS_RETURN_IF_ERROR(ms->parse());
  HPS_RETURN_IF_ERROR(ms->ensure_caches());
  server->set_version(name, version);
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelSetState(model, ms.get()));
  ms.release();
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_ModelFinalize(TRITONBACKEND_Model* model) {
  void* vstate = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelState(model, &vstate));
  ModelState* ms = static_cast<ModelState*>(vstate);
  if (ms != nullptr) HPS_LOG(INFO, "TRITONBACKEND_ModelFinalize: delete model state of ", ms->name);
  // The caches stay with the parameter server: another version of the model may be loading right now
  // and shares them (reference keeps them unless version_ps_ == version_, src/model_state.cpp:108-122).
  delete ms;
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_ModelInstanceInitialize(TRITONBACKEND_ModelInstance* instance) {
  const char* name = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceName(instance, &name));
  TRITONBACKEND_Model* model = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceModel(instance, &model));
  void* vmodel = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelState(model, &vmodel));
  ModelState* ms = static_cast<ModelState*>(vmodel);
  if (ms == nullptr) return HPS_ERROR(INTERNAL, "model instance ", name ? name : "", " has no model state");
  int32_t device_id = 0;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceDeviceId(instance, &device_id));
  TRITONSERVER_InstanceGroupKind kind = TRITONSERVER_INSTANCEGROUPKIND_AUTO;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceKind(instance, &kind));
  HPS_LOG(INFO, "TRITONBACKEND_ModelInstanceInitialize: ", name ? name : "", " (device ", device_id, ")");
  if (ms->gpucache()) {
    if (kind != TRITONSERVER_INSTANCEGROUPKIND_GPU)
      return HPS_ERROR(INVALID_ARG, "model ", ms->name, " uses the GPU embedding cache: instance ", name ? name : "",
                       " must be KIND_GPU");
    if (std::find(ms->gpus.begin(), ms->gpus.end(), static_cast<int>(device_id)) == ms->gpus.end())
      return HPS_ERROR(INVALID_ARG, "instance ", name ? name : "", " runs on device ", device_id,
                       ", which is not in the instance_group gpus of model ", ms->name);
  }
  std::unique_ptr<InstanceState> inst(new InstanceState());
  inst->triton_instance = instance;
  inst->model = ms;
  inst->name = name ? name : "";
  inst->device = ms->gpucache() ? device_id : 0;
  HPS_LOG(INFO, "******Loading HPS ******");
  HPS_RETURN_IF_ENGINE_ERROR(
      hpsx_session_create(ms->server->ps, ms->name.c_str(), ms->gpucache() ? device_id : -1, &inst->session),
      "creating the lookup session of instance " + inst->name);
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceSetState(instance, inst.get()));
  inst.release();
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_ModelInstanceFinalize(TRITONBACKEND_ModelInstance* instance) {
  void* vstate = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceState(instance, &vstate));
  HPS_LOG(INFO, "TRITONBACKEND_ModelInstanceFinalize: delete instance state");
  delete static_cast<InstanceState*>(vstate);
  return nullptr;
}

TRITONSERVER_Error* TRITONBACKEND_ModelInstanceExecute(TRITONBACKEND_ModelInstance* instance,
                                                       TRITONBACKEND_Request** requests,
                                                       const uint32_t request_count) {
  void* vstate = nullptr;
  HPS_RETURN_IF_ERROR(TRITONBACKEND_ModelInstanceState(instance, &vstate));
  InstanceState* inst = static_cast<InstanceState*>(vstate);
  if (inst == nullptr || inst->session == nullptr)
    return HPS_ERROR(INTERNAL, "model instance has no HPS lookup session");

  // One response per request; failing to create them fails the whole call and leaves the requests
  // with Triton (src/hps.cc:381-390).
  std::vector<TRITONBACKEND_Response*> responses;
  responses.reserve(request_count);
  for (uint32_t r = 0; r < request_count; ++r) {
    TRITONBACKEND_Response* response = nullptr;
    HPS_RETURN_IF_ERROR(TRITONBACKEND_ResponseNew(&response, requests[r]));
    responses.push_back(response);
  }
  // From here on the requests are ours: exactly one FINAL response and one release each.
  uint64_t min_exec_start_ns = UINT64_MAX, max_exec_end_ns = 0, total_batch_size = 0;
  uint64_t batch_compute_start_ns = UINT64_MAX, batch_compute_end_ns = 0;
  std::vector<bool> answered(request_count, false);  // a success response went out
  try {
    // phase 1: validate every request and obtain its output buffer
    std::vector<RequestPlan> plans(request_count);
    std::vector<uint64_t> exec_start(request_count, 0), comp_start(request_count, 0), comp_end(request_count, 0);
    std::vector<bool> ok(request_count, false);
    for (uint32_t r = 0; r < request_count; ++r) {
      exec_start[r] = now_ns();
      min_exec_start_ns = std::min(min_exec_start_ns, exec_start[r]);
      TRITONSERVER_Error* err = prepare_request(inst, requests[r], responses[r], &plans[r]);
      if (err != nullptr) {
        HPS_LOG(ERROR, "request ", r, " of instance ", inst->name, ": ", TRITONSERVER_ErrorMessage(err),
                ", error response sent");
        respond_error(responses, r, err);
        continue;
      }
      ok[r] = true;
    }
    // phase 2: lookups.  Consecutive un-pooled GPU requests with device output buffers share one engine pass
    // (cross-request batching) as long as they fit one request's key budget; everything else runs alone.
    ModelState* ms = inst->model;
    hpsx_session_stats stats0{}, stats1{};
    const bool report = ms->report_stats && hpsx_session_get_stats(inst->session, &stats0) == HPSX_OK;
    const uint64_t key_budget = static_cast<uint64_t>(ms->max_batch_size) * ms->cat_num;
    auto batchable = [&](uint32_t r) {
      return ok[r] && plans[r].has_output && ms->gpucache() && ms->pooling < 0 && plans[r].out_on_device;
    };
    uint32_t r = 0;
    while (r < request_count) {
      if (!ok[r]) {
        ++r;
        continue;
      }
      std::vector<uint32_t> group{r};
      if (batchable(r)) {
        uint64_t keys_in_group = plans[r].num_keys;
        for (uint32_t q = r + 1; q < request_count && group.size() < HPSX_MAX_BATCH_REQUESTS; ++q) {
          if (!ok[q]) continue;  // already answered with an error
          if (!batchable(q) || plans[q].keys_on_device != plans[r].keys_on_device ||
              keys_in_group + plans[q].num_keys > key_budget)
            break;
          keys_in_group += plans[q].num_keys;
          group.push_back(q);
        }
      }
      const uint64_t t0 = now_ns();
      TRITONSERVER_Error* err = nullptr;
      if (group.size() > 1)
        err = run_batch(inst, plans, group);
      else if (plans[r].has_output)
        err = run_single(inst, plans[r]);
      const uint64_t t1 = now_ns();
      for (uint32_t q : group) {
        comp_start[q] = t0;
        comp_end[q] = t1;
        if (err != nullptr) {
          HPS_LOG(ERROR, "request ", q, " of instance ", inst->name, ": ", TRITONSERVER_ErrorMessage(err),
                  ", error response sent");
          respond_error(responses, q,
                        TRITONSERVER_ErrorNew(TRITONSERVER_ErrorCode(err), TRITONSERVER_ErrorMessage(err)));
          ok[q] = false;
        }
      }
      if (err != nullptr) TRITONSERVER_ErrorDelete(err);
      r = group.back() + 1;
    }
    const uint64_t t_phase3 = now_ns();
    const bool reported = report && hpsx_session_get_stats(inst->session, &stats1) == HPSX_OK;
    // phase 3: responses and statistics
    for (uint32_t q = 0; q < request_count; ++q) {
      if (!ok[q]) continue;
      if (reported) {
        HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSetIntParameter(responses[q], "CacheHits",
                                                               static_cast<int64_t>(stats1.hits - stats0.hits)),
                         "failed return cache hits");
        HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSetIntParameter(responses[q], "CacheMisses",
                                                               static_cast<int64_t>(stats1.misses - stats0.misses)),
                         "failed return cache misses");
      }
      HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSetIntParameter(responses[q], "NumSample", plans[q].num_samples),
                       "failed return Number of samples");
      HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSetIntParameter(responses[q], "DeviceID", inst->device),
                       "failed return device id");
      HPS_LOG_IF_ERROR(TRITONBACKEND_ResponseSend(responses[q], TRITONSERVER_RESPONSE_COMPLETE_FINAL, nullptr),
                       "failed sending response");
      answered[q] = true;
      const uint64_t exec_end_ns = now_ns();
      if (comp_start[q] == 0) comp_start[q] = comp_end[q] = exec_start[q];
      max_exec_end_ns = std::max(max_exec_end_ns, exec_end_ns);
      batch_compute_start_ns = std::min(batch_compute_start_ns, comp_start[q]);
      batch_compute_end_ns = std::max(batch_compute_end_ns, comp_end[q]);
      total_batch_size += static_cast<uint64_t>(plans[q].num_samples);
      HPS_LOG_IF_ERROR(TRITONBACKEND_ModelInstanceReportStatistics(instance, requests[q], true /* success */,
                                                                   exec_start[q], comp_start[q], comp_end[q],
                                                                   exec_end_ns),
                       "failed reporting request statistics");
    }
    static const bool trace = std::getenv("HPS_TRACE") != nullptr;
    if (trace && request_count > 0)
      std::fprintf(stderr, "[hps] execute %u request(s): prepare %.3f ms | lookup %.3f ms | respond %.3f ms\n", request_count,
                   (comp_start[0] ? comp_start[0] - exec_start[0] : 0) / 1e6,
                   (comp_end[request_count - 1] - comp_start[0]) / 1e6, (now_ns() - t_phase3) / 1e6);
  } catch (const std::exception& e) {
    // nothing may unwind through the C ABI: fail whatever has not been answered yet
    for (uint32_t r = 0; r < request_count; ++r)
      if (responses[r] != nullptr && !answered[r])
        respond_error(responses, r, HPS_ERROR(INTERNAL, "hps backend: ", e.what()));
  }
  if (max_exec_end_ns != 0) {
    HPS_LOG_IF_ERROR(TRITONBACKEND_ModelInstanceReportBatchStatistics(instance, total_batch_size, min_exec_start_ns,
                                                                      batch_compute_start_ns, batch_compute_end_ns,
                                                                      max_exec_end_ns),
                     "failed reporting batch request statistics");
  }
  for (uint32_t r = 0; r < request_count; ++r) {
    // a response that was already sent as an error is nullptr here: record the failure (timestamps ignored)
    if (responses[r] == nullptr) {
      HPS_LOG_IF_ERROR(TRITONBACKEND_ModelInstanceReportStatistics(instance, requests[r], false, 0, 0, 0, 0),
                       "failed reporting request statistics");
    }
    HPS_LOG_IF_ERROR(TRITONBACKEND_RequestRelease(requests[r], TRITONSERVER_REQUEST_RELEASE_ALL),
                     "failed releasing request");
  }
  return nullptr;
}

}  // extern "C"
