// fake_triton — TEST INFRASTRUCTURE.  A stand-in for the Triton server process: it defines every
// TRITONSERVER_* / TRITONBACKEND_* function that libtriton_hps.so imports (include/triton_compat.h),
// dlopen()s the backend the way Triton does, drives its seven exported entry points, and records what
// the backend did (responses, parameters, statistics, releases, leaked errors) for the tests to check.
// Controlled from Python through the ft_* C functions at the bottom (tests/fake_triton/__init__.py).
//
// No CUDA here: a GPU output buffer is whatever device pointer the test hands in (a torch tensor);
// without one the harness falls back to CPU memory, as Triton may (hps_backend/src/hps.cc:638-648).
#include <chrono>
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "triton_compat.h"

// ------------------------------------------------------------------------------------------------
// concrete definitions of Triton's opaque types
// ------------------------------------------------------------------------------------------------
struct TRITONSERVER_Error {
  TRITONSERVER_Error_Code code;
  std::string msg;
};
struct TRITONSERVER_Message {
  std::string json;
};
struct TRITONSERVER_Server {
  int dummy;
};

using BackendFn = TRITONSERVER_Error* (*)(TRITONBACKEND_Backend*);
using ModelFn = TRITONSERVER_Error* (*)(TRITONBACKEND_Model*);
using InstanceFn = TRITONSERVER_Error* (*)(TRITONBACKEND_ModelInstance*);
using ExecuteFn = TRITONSERVER_Error* (*)(TRITONBACKEND_ModelInstance*, TRITONBACKEND_Request**, const uint32_t);

struct TRITONBACKEND_Backend {
  std::string name, artifacts;
  TRITONSERVER_Message config;
  void* state = nullptr;
  void* dl = nullptr;
  BackendFn init = nullptr, fini = nullptr;
  ModelFn model_init = nullptr, model_fini = nullptr;
  InstanceFn inst_init = nullptr, inst_fini = nullptr;
  ExecuteFn execute = nullptr;
  TRITONSERVER_Server server{0};
};
struct TRITONBACKEND_Model {
  TRITONBACKEND_Backend* backend = nullptr;
  std::string name, repo, config_json;
  uint64_t version = 1;
  void* state = nullptr;
};
struct TRITONBACKEND_ModelInstance {
  TRITONBACKEND_Model* model = nullptr;
  std::string name;
  TRITONSERVER_InstanceGroupKind kind = TRITONSERVER_INSTANCEGROUPKIND_GPU;
  int32_t device = 0;
  void* state = nullptr;
  // statistics reported by the backend
  uint64_t ok_requests = 0, failed_requests = 0, batch_reports = 0, last_batch_size = 0;
  uint64_t last_exec_start = 0, last_compute_start = 0, last_compute_end = 0, last_exec_end = 0;
};
struct InputBufferRec {
  const void* ptr;
  uint64_t bytes;
  TRITONSERVER_MemoryType mt;
  int64_t mt_id;
};
struct TRITONBACKEND_Input {
  std::string name;
  TRITONSERVER_DataType dtype = TRITONSERVER_TYPE_INVALID;
  std::vector<int64_t> shape;
  uint64_t byte_size = 0;
  std::vector<InputBufferRec> buffers;
};
struct TRITONBACKEND_Output {
  std::string name;
  TRITONSERVER_DataType dtype = TRITONSERVER_TYPE_INVALID;
  std::vector<int64_t> shape;
  void* buffer = nullptr;
  uint64_t bytes = 0;
  TRITONSERVER_MemoryType mt = TRITONSERVER_MEMORY_CPU;
  int64_t mt_id = 0;
  bool owned = false;
  TRITONBACKEND_Request* request = nullptr;
};
struct TRITONBACKEND_Response {
  TRITONBACKEND_Request* request = nullptr;
  std::vector<std::unique_ptr<TRITONBACKEND_Output>> outputs;
  std::map<std::string, int64_t> int_params;
  int sent = 0;
  uint32_t flags = 0;
  bool has_error = false;
  TRITONSERVER_Error_Code err_code = TRITONSERVER_ERROR_UNKNOWN;
  std::string err_msg;
  ~TRITONBACKEND_Response() {
    for (auto& o : outputs)
      if (o->owned) std::free(o->buffer);
  }
};
struct TRITONBACKEND_Request {
  std::string id;
  uint64_t correlation_id = 0;
  std::vector<std::unique_ptr<TRITONBACKEND_Input>> inputs;
  std::vector<std::string> requested_outputs;
  int released = 0;
  std::vector<std::unique_ptr<TRITONBACKEND_Response>> responses;
  // output placement policy
  void* gpu_out = nullptr;
  uint64_t gpu_out_cap = 0;
  int64_t gpu_out_device = 0;
  void* cpu_out = nullptr;  // caller-owned host buffer handed out when the output ends up in CPU memory
  uint64_t cpu_out_cap = 0;
  int cpu_out_mt = 0;       // TRITONSERVER_MEMORY_CPU or TRITONSERVER_MEMORY_CPU_PINNED, as the caller allocated it
  int force_output_memory = -1;  // -1: honour the backend's preference when possible
  bool fail_output_buffer = false;
};

namespace {
std::atomic<long> g_live_errors{0}, g_live_messages{0};
uint32_t g_api_major = TRITONBACKEND_API_VERSION_MAJOR, g_api_minor = TRITONBACKEND_API_VERSION_MINOR;
std::mutex g_log_mu;
long g_log_count[4] = {0, 0, 0, 0};
std::string g_last_log[4];
thread_local std::string g_ft_error;
thread_local int g_ft_error_code = 0;

TRITONSERVER_Error* new_error(TRITONSERVER_Error_Code c, const std::string& m) {
  ++g_live_errors;
  return new TRITONSERVER_Error{c, m};
}
// consumes a backend-returned error: records it for ft_last_error and frees it
int consume(TRITONSERVER_Error* e) {
  if (e == nullptr) {
    g_ft_error.clear();
    g_ft_error_code = 0;
    return 0;
  }
  g_ft_error = e->msg;
  g_ft_error_code = 100 + static_cast<int>(e->code);
  TRITONSERVER_ErrorDelete(e);
  return g_ft_error_code;
}
int ft_fail(const std::string& m) {
  g_ft_error = m;
  g_ft_error_code = 1;
  return 1;
}
}  // namespace

extern "C" {

// ================================================================================================
// TRITONSERVER_*
// ================================================================================================
TRITONSERVER_Error* TRITONSERVER_ErrorNew(TRITONSERVER_Error_Code code, const char* msg) {
  return new_error(code, msg ? msg : "");
}
void TRITONSERVER_ErrorDelete(TRITONSERVER_Error* error) {
  if (error != nullptr) {
    --g_live_errors;
    delete error;
  }
}
TRITONSERVER_Error_Code TRITONSERVER_ErrorCode(TRITONSERVER_Error* error) { return error->code; }
const char* TRITONSERVER_ErrorCodeString(TRITONSERVER_Error* error) {
  switch (error->code) {
    case TRITONSERVER_ERROR_UNKNOWN: return "Unknown";
    case TRITONSERVER_ERROR_INTERNAL: return "Internal";
    case TRITONSERVER_ERROR_NOT_FOUND: return "Not found";
    case TRITONSERVER_ERROR_INVALID_ARG: return "Invalid argument";
    case TRITONSERVER_ERROR_UNAVAILABLE: return "Unavailable";
    case TRITONSERVER_ERROR_UNSUPPORTED: return "Unsupported";
    case TRITONSERVER_ERROR_ALREADY_EXISTS: return "Already exists";
  }
  return "<invalid code>";
}
const char* TRITONSERVER_ErrorMessage(TRITONSERVER_Error* error) { return error->msg.c_str(); }

TRITONSERVER_Error* TRITONSERVER_LogMessage(TRITONSERVER_LogLevel level, const char* filename, const int line,
                                            const char* msg) {
  std::lock_guard<std::mutex> lk(g_log_mu);
  const int l = static_cast<int>(level) & 3;
  ++g_log_count[l];
  g_last_log[l] = msg ? msg : "";
  static const bool echo = std::getenv("FT_LOG") != nullptr;
  if (echo) {
    const char* base = filename ? std::strrchr(filename, '/') : nullptr;
    std::fprintf(stderr, "%c %s:%d] %s\n", "IWEV"[l], base ? base + 1 : (filename ? filename : "?"), line,
                 msg ? msg : "");
  }
  return nullptr;
}
TRITONSERVER_Error* TRITONSERVER_MessageSerializeToJson(TRITONSERVER_Message* message, const char** base,
                                                        size_t* byte_size) {
  if (message == nullptr) return new_error(TRITONSERVER_ERROR_INVALID_ARG, "null message");
  *base = message->json.c_str();
  *byte_size = message->json.size();
  return nullptr;
}
TRITONSERVER_Error* TRITONSERVER_MessageDelete(TRITONSERVER_Message* message) {
  if (message != nullptr) {
    --g_live_messages;
    delete message;
  }
  return nullptr;
}
const char* TRITONSERVER_DataTypeString(TRITONSERVER_DataType datatype) {
  static const char* names[] = {"<invalid>", "BOOL", "UINT8", "UINT16", "UINT32", "UINT64", "INT8", "INT16",
                                "INT32",     "INT64", "FP16", "FP32",   "FP64",   "BYTES",  "BF16"};
  const int i = static_cast<int>(datatype);
  return (i >= 0 && i <= 14) ? names[i] : "<invalid>";
}

// ================================================================================================
// TRITONBACKEND_* backend / model / instance
// ================================================================================================
TRITONSERVER_Error* TRITONBACKEND_ApiVersion(uint32_t* major, uint32_t* minor) {
  *major = g_api_major;
  *minor = g_api_minor;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_BackendName(TRITONBACKEND_Backend* b, const char** name) {
  *name = b->name.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_BackendConfig(TRITONBACKEND_Backend* b, TRITONSERVER_Message** cfg) {
  *cfg = &b->config;  // owned by the backend object, like Triton's
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_BackendArtifacts(TRITONBACKEND_Backend* b, TRITONBACKEND_ArtifactType* t,
                                                   const char** location) {
  *t = TRITONBACKEND_ARTIFACT_FILESYSTEM;
  *location = b->artifacts.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_BackendState(TRITONBACKEND_Backend* b, void** state) {
  *state = b->state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_BackendSetState(TRITONBACKEND_Backend* b, void* state) {
  b->state = state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelName(TRITONBACKEND_Model* m, const char** name) {
  *name = m->name.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelVersion(TRITONBACKEND_Model* m, uint64_t* version) {
  *version = m->version;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelRepository(TRITONBACKEND_Model* m, TRITONBACKEND_ArtifactType* t,
                                                  const char** location) {
  *t = TRITONBACKEND_ARTIFACT_FILESYSTEM;
  *location = m->repo.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelConfig(TRITONBACKEND_Model* m, const uint32_t config_version,
                                              TRITONSERVER_Message** model_config) {
  if (config_version != 1) return new_error(TRITONSERVER_ERROR_UNSUPPORTED, "model config version must be 1");
  ++g_live_messages;  // the caller owns this one and must TRITONSERVER_MessageDelete it
  *model_config = new TRITONSERVER_Message{m->config_json};
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelServer(TRITONBACKEND_Model* m, TRITONSERVER_Server** server) {
  *server = &m->backend->server;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelBackend(TRITONBACKEND_Model* m, TRITONBACKEND_Backend** backend) {
  *backend = m->backend;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelState(TRITONBACKEND_Model* m, void** state) {
  *state = m->state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelSetState(TRITONBACKEND_Model* m, void* state) {
  m->state = state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceName(TRITONBACKEND_ModelInstance* i, const char** name) {
  *name = i->name.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceKind(TRITONBACKEND_ModelInstance* i,
                                                    TRITONSERVER_InstanceGroupKind* kind) {
  *kind = i->kind;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceDeviceId(TRITONBACKEND_ModelInstance* i, int32_t* device_id) {
  *device_id = i->device;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceModel(TRITONBACKEND_ModelInstance* i, TRITONBACKEND_Model** model) {
  *model = i->model;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceState(TRITONBACKEND_ModelInstance* i, void** state) {
  *state = i->state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceSetState(TRITONBACKEND_ModelInstance* i, void* state) {
  i->state = state;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceReportStatistics(TRITONBACKEND_ModelInstance* i,
                                                                TRITONBACKEND_Request* request, const bool success,
                                                                const uint64_t exec_start_ns,
                                                                const uint64_t compute_start_ns,
                                                                const uint64_t compute_end_ns,
                                                                const uint64_t exec_end_ns) {
  (void)request;
  if (success) {
    ++i->ok_requests;
    i->last_exec_start = exec_start_ns;
    i->last_compute_start = compute_start_ns;
    i->last_compute_end = compute_end_ns;
    i->last_exec_end = exec_end_ns;
  } else {
    ++i->failed_requests;
  }
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ModelInstanceReportBatchStatistics(TRITONBACKEND_ModelInstance* i,
                                                                     const uint64_t batch_size, const uint64_t,
                                                                     const uint64_t, const uint64_t, const uint64_t) {
  ++i->batch_reports;
  i->last_batch_size = batch_size;
  return nullptr;
}

// ================================================================================================
// request / input / response / output
// ================================================================================================
TRITONSERVER_Error* TRITONBACKEND_RequestId(TRITONBACKEND_Request* r, const char** id) {
  *id = r->id.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestCorrelationId(TRITONBACKEND_Request* r, uint64_t* id) {
  *id = r->correlation_id;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestInputCount(TRITONBACKEND_Request* r, uint32_t* count) {
  *count = static_cast<uint32_t>(r->inputs.size());
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestInputName(TRITONBACKEND_Request* r, const uint32_t index,
                                                   const char** input_name) {
  if (index >= r->inputs.size())
    return new_error(TRITONSERVER_ERROR_INVALID_ARG, "out of bounds index " + std::to_string(index) +
                                                         ": request has " + std::to_string(r->inputs.size()) +
                                                         " inputs");
  *input_name = r->inputs[index]->name.c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestInput(TRITONBACKEND_Request* r, const char* name,
                                               TRITONBACKEND_Input** input) {
  for (auto& in : r->inputs)
    if (in->name == name) {
      *input = in.get();
      return nullptr;
    }
  return new_error(TRITONSERVER_ERROR_INVALID_ARG, std::string("unknown request input name ") + name);
}
TRITONSERVER_Error* TRITONBACKEND_RequestOutputCount(TRITONBACKEND_Request* r, uint32_t* count) {
  *count = static_cast<uint32_t>(r->requested_outputs.size());
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestOutputName(TRITONBACKEND_Request* r, const uint32_t index,
                                                    const char** output_name) {
  if (index >= r->requested_outputs.size())
    return new_error(TRITONSERVER_ERROR_INVALID_ARG, "out of bounds requested output index");
  *output_name = r->requested_outputs[index].c_str();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_RequestRelease(TRITONBACKEND_Request* r, uint32_t release_flags) {
  if ((release_flags & TRITONSERVER_REQUEST_RELEASE_ALL) != 0) ++r->released;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_InputProperties(TRITONBACKEND_Input* in, const char** name,
                                                  TRITONSERVER_DataType* datatype, const int64_t** shape,
                                                  uint32_t* dims_count, uint64_t* byte_size,
                                                  uint32_t* buffer_count) {
  if (name) *name = in->name.c_str();
  if (datatype) *datatype = in->dtype;
  if (shape) *shape = in->shape.data();
  if (dims_count) *dims_count = static_cast<uint32_t>(in->shape.size());
  if (byte_size) *byte_size = in->byte_size;
  if (buffer_count) *buffer_count = static_cast<uint32_t>(in->buffers.size());
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_InputBuffer(TRITONBACKEND_Input* in, const uint32_t index, const void** buffer,
                                              uint64_t* buffer_byte_size, TRITONSERVER_MemoryType* memory_type,
                                              int64_t* memory_type_id) {
  if (index >= in->buffers.size())
    return new_error(TRITONSERVER_ERROR_INVALID_ARG, "out of bounds index " + std::to_string(index) + ": input " +
                                                         in->name + " has " + std::to_string(in->buffers.size()) +
                                                         " buffers");
  const InputBufferRec& b = in->buffers[index];
  *buffer = b.ptr;
  *buffer_byte_size = b.bytes;
  *memory_type = b.mt;  // the data is where it is, whatever the caller preferred
  *memory_type_id = b.mt_id;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ResponseNew(TRITONBACKEND_Response** response, TRITONBACKEND_Request* request) {
  request->responses.emplace_back(new TRITONBACKEND_Response());
  request->responses.back()->request = request;
  *response = request->responses.back().get();
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ResponseOutput(TRITONBACKEND_Response* response, TRITONBACKEND_Output** output,
                                                 const char* name, const TRITONSERVER_DataType datatype,
                                                 const int64_t* shape, const uint32_t dims_count) {
  response->outputs.emplace_back(new TRITONBACKEND_Output());
  TRITONBACKEND_Output* o = response->outputs.back().get();
  o->name = name ? name : "";
  o->dtype = datatype;
  o->shape.assign(shape, shape + dims_count);
  o->request = response->request;
  *output = o;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_OutputBuffer(TRITONBACKEND_Output* o, void** buffer,
                                               const uint64_t buffer_byte_size, TRITONSERVER_MemoryType* memory_type,
                                               int64_t* memory_type_id) {
  // the output's owning request: the one whose ResponseOutput call created it
  TRITONBACKEND_Request* req = o->request;
  if (req != nullptr && req->fail_output_buffer)
    return new_error(TRITONSERVER_ERROR_INTERNAL, "fake_triton: output buffer allocation failed (injected)");
  TRITONSERVER_MemoryType want = *memory_type;
  if (req != nullptr && req->force_output_memory >= 0)
    want = static_cast<TRITONSERVER_MemoryType>(req->force_output_memory);
  if (want == TRITONSERVER_MEMORY_GPU && req != nullptr && req->gpu_out != nullptr &&
      buffer_byte_size <= req->gpu_out_cap) {
    o->buffer = req->gpu_out;
    o->mt = TRITONSERVER_MEMORY_GPU;
    o->mt_id = req->gpu_out_device;
    o->owned = false;
  } else if (req != nullptr && req->cpu_out != nullptr && buffer_byte_size <= req->cpu_out_cap) {
    // Triton may override the preference: CPU memory, here a buffer of the caller (e.g. from the pinned pool)
    o->buffer = req->cpu_out;
    o->mt = static_cast<TRITONSERVER_MemoryType>(req->cpu_out_mt);
    o->mt_id = 0;
    o->owned = false;
  } else {
    // Triton may override the preference: CPU memory
    o->buffer = buffer_byte_size ? std::malloc(buffer_byte_size) : nullptr;
    if (buffer_byte_size && o->buffer == nullptr)
      return new_error(TRITONSERVER_ERROR_INTERNAL, "fake_triton: out of host memory");
    if (o->buffer) std::memset(o->buffer, 0xFF, buffer_byte_size);  // NaN pattern: unwritten floats are visible
    o->mt = TRITONSERVER_MEMORY_CPU;
    o->mt_id = 0;
    o->owned = true;
  }
  o->bytes = buffer_byte_size;
  *buffer = o->buffer;
  *memory_type = o->mt;
  *memory_type_id = o->mt_id;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ResponseSetIntParameter(TRITONBACKEND_Response* response, const char* name,
                                                          const int64_t value) {
  if (response == nullptr) return new_error(TRITONSERVER_ERROR_INVALID_ARG, "null response");
  response->int_params[name] = value;
  return nullptr;
}
TRITONSERVER_Error* TRITONBACKEND_ResponseSend(TRITONBACKEND_Response* response, const uint32_t send_flags,
                                               TRITONSERVER_Error* error) {
  if (response == nullptr) return new_error(TRITONSERVER_ERROR_INVALID_ARG, "null response");
  ++response->sent;
  response->flags = send_flags;
  if (error != nullptr) {  // not owned by this call: the backend deletes it
    response->has_error = true;
    response->err_code = error->code;
    response->err_msg = error->msg;
  }
  return nullptr;
}

// ================================================================================================
// ft_* control surface
// ================================================================================================
const char* ft_last_error(void) { return g_ft_error.c_str(); }
int ft_last_error_code(void) { return g_ft_error_code; }
long ft_live_errors(void) { return g_live_errors.load(); }
long ft_live_messages(void) { return g_live_messages.load(); }
void ft_set_api_version(uint32_t major, uint32_t minor) {
  g_api_major = major;
  g_api_minor = minor;
}
long ft_log_count(int level) {
  std::lock_guard<std::mutex> lk(g_log_mu);
  return g_log_count[level & 3];
}
const char* ft_last_log(int level) {
  std::lock_guard<std::mutex> lk(g_log_mu);
  thread_local std::string copy;
  copy = g_last_log[level & 3];
  return copy.c_str();
}

// dlopen the backend and run TRITONBACKEND_Initialize.  Returns 0 or an error code (100 + Triton code).
int ft_backend_load(const char* so_path, const char* name, const char* backend_config_json, const char* artifacts,
                    TRITONBACKEND_Backend** out) {
  *out = nullptr;
  std::unique_ptr<TRITONBACKEND_Backend> b(new TRITONBACKEND_Backend());
  b->name = name ? name : "hps";
  b->artifacts = artifacts ? artifacts : "";
  b->config.json = backend_config_json ? backend_config_json : "{}";
  b->dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (b->dl == nullptr) return ft_fail(std::string("dlopen failed: ") + dlerror());
#define FT_SYM(field, type, sym)                                                            \
  b->field = reinterpret_cast<type>(dlsym(b->dl, sym));                                     \
  if (b->field == nullptr) return ft_fail(std::string("backend does not export ") + sym);
  FT_SYM(init, BackendFn, "TRITONBACKEND_Initialize")
  FT_SYM(fini, BackendFn, "TRITONBACKEND_Finalize")
  FT_SYM(model_init, ModelFn, "TRITONBACKEND_ModelInitialize")
  FT_SYM(model_fini, ModelFn, "TRITONBACKEND_ModelFinalize")
  FT_SYM(inst_init, InstanceFn, "TRITONBACKEND_ModelInstanceInitialize")
  FT_SYM(inst_fini, InstanceFn, "TRITONBACKEND_ModelInstanceFinalize")
  FT_SYM(execute, ExecuteFn, "TRITONBACKEND_ModelInstanceExecute")
#undef FT_SYM
  const int rc = consume(b->init(b.get()));
  if (rc != 0) {
    // Triton does not call Finalize when Initialize failed; state (if any was set) is the backend's leak
    return rc;
  }
  *out = b.release();
  return 0;
}
int ft_backend_unload(TRITONBACKEND_Backend* b) {
  if (b == nullptr) return 0;
  const int rc = consume(b->fini(b));
  // keep the library mapped: thread-local destructors of the backend may still run at thread exit
  delete b;
  return rc;
}
int ft_backend_has_state(TRITONBACKEND_Backend* b) { return b->state != nullptr; }

int ft_model_load(TRITONBACKEND_Backend* b, const char* name, uint64_t version, const char* config_json,
                  const char* repo, TRITONBACKEND_Model** out) {
  *out = nullptr;
  std::unique_ptr<TRITONBACKEND_Model> m(new TRITONBACKEND_Model());
  m->backend = b;
  m->name = name;
  m->version = version;
  m->config_json = config_json ? config_json : "{}";
  m->repo = repo ? repo : "";
  const int rc = consume(b->model_init(m.get()));
  if (rc != 0) return rc;
  *out = m.release();
  return 0;
}
int ft_model_unload(TRITONBACKEND_Model* m) {
  if (m == nullptr) return 0;
  const int rc = consume(m->backend->model_fini(m));
  delete m;
  return rc;
}
int ft_instance_create(TRITONBACKEND_Model* m, const char* name, int kind, int device,
                       TRITONBACKEND_ModelInstance** out) {
  *out = nullptr;
  std::unique_ptr<TRITONBACKEND_ModelInstance> i(new TRITONBACKEND_ModelInstance());
  i->model = m;
  i->name = name;
  i->kind = static_cast<TRITONSERVER_InstanceGroupKind>(kind);
  i->device = device;
  const int rc = consume(m->backend->inst_init(i.get()));
  if (rc != 0) return rc;
  *out = i.release();
  return 0;
}
int ft_instance_destroy(TRITONBACKEND_ModelInstance* i) {
  if (i == nullptr) return 0;
  const int rc = consume(i->model->backend->inst_fini(i));
  delete i;
  return rc;
}
void ft_instance_stats(TRITONBACKEND_ModelInstance* i, uint64_t* out8) {
  out8[0] = i->ok_requests;
  out8[1] = i->failed_requests;
  out8[2] = i->batch_reports;
  out8[3] = i->last_batch_size;
  out8[4] = i->last_exec_start;
  out8[5] = i->last_compute_start;
  out8[6] = i->last_compute_end;
  out8[7] = i->last_exec_end;
}

TRITONBACKEND_Request* ft_request_new(const char* id, uint64_t correlation_id) {
  TRITONBACKEND_Request* r = new TRITONBACKEND_Request();
  r->id = id ? id : "";
  r->correlation_id = correlation_id;
  return r;
}
void ft_request_delete(TRITONBACKEND_Request* r) { delete r; }
int ft_request_add_input(TRITONBACKEND_Request* r, const char* name, int dtype, const int64_t* shape, uint32_t dims,
                         uint64_t byte_size) {
  r->inputs.emplace_back(new TRITONBACKEND_Input());
  TRITONBACKEND_Input* in = r->inputs.back().get();
  in->name = name;
  in->dtype = static_cast<TRITONSERVER_DataType>(dtype);
  in->shape.assign(shape, shape + dims);
  in->byte_size = byte_size;
  return static_cast<int>(r->inputs.size()) - 1;
}
void ft_request_input_append_buffer(TRITONBACKEND_Request* r, int input_index, const void* ptr, uint64_t bytes,
                                    int memory_type, int64_t memory_type_id) {
  r->inputs[input_index]->buffers.push_back(
      InputBufferRec{ptr, bytes, static_cast<TRITONSERVER_MemoryType>(memory_type), memory_type_id});
}
void ft_request_add_requested_output(TRITONBACKEND_Request* r, const char* name) {
  r->requested_outputs.emplace_back(name);
}
void ft_request_set_gpu_output(TRITONBACKEND_Request* r, void* d_ptr, uint64_t capacity, int64_t device) {
  r->gpu_out = d_ptr;
  r->gpu_out_cap = capacity;
  r->gpu_out_device = device;
}
void ft_request_set_cpu_output(TRITONBACKEND_Request* r, void* h_ptr, uint64_t capacity, int memory_type) {
  r->cpu_out = h_ptr;
  r->cpu_out_cap = capacity;
  r->cpu_out_mt = memory_type;
}
void ft_request_force_output_memory(TRITONBACKEND_Request* r, int memory_type) { r->force_output_memory = memory_type; }
void ft_request_fail_output_buffer(TRITONBACKEND_Request* r, int fail) { r->fail_output_buffer = fail != 0; }

// Runs TRITONBACKEND_ModelInstanceExecute on `n` requests (one call, like Triton's scheduler).
int ft_execute(TRITONBACKEND_ModelInstance* i, TRITONBACKEND_Request** reqs, uint32_t n) {
  return consume(i->model->backend->execute(i, reqs, n));
}

// The same Execute call `repeat` times over prepared requests (each pass: one call carrying all `n` requests), timed
// here so that a benchmark of small requests measures the backend and not the Python harness.  Responses of earlier
// passes are dropped; the last pass's stay for inspection.
int ft_execute_repeat(TRITONBACKEND_ModelInstance* i, TRITONBACKEND_Request** reqs, uint32_t n, uint32_t repeat,
                      uint64_t* elapsed_ns) {
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t rep = 0; rep < repeat; ++rep) {
    for (uint32_t q = 0; q < n; ++q) reqs[q]->responses.clear();
    const int rc = consume(i->model->backend->execute(i, reqs, n));
    if (rc != 0) return rc;
  }
  if (elapsed_ns != nullptr)
    *elapsed_ns = static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

// A stream of DISTINCT requests on one instance: one Execute call per request, in order, timed here.
int ft_execute_sequence(TRITONBACKEND_ModelInstance* i, TRITONBACKEND_Request** reqs, uint32_t n, uint64_t* elapsed_ns) {
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t q = 0; q < n; ++q) {
    reqs[q]->responses.clear();
    const int rc = consume(i->model->backend->execute(i, reqs + q, 1));
    if (rc != 0) return rc;
  }
  if (elapsed_ns != nullptr)
    *elapsed_ns = static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
  return 0;
}

// The same for `k` instances at once, one thread per instance — a server process driving several GPUs.  `reqs` is the
// concatenation of the instances' sequences (counts[j] requests for instance j).  *elapsed_ns = wall time from the
// common start to the end of the last thread; returns the first non-zero thread result.
int ft_execute_sequences_parallel(TRITONBACKEND_ModelInstance** insts, uint32_t k, TRITONBACKEND_Request** reqs,
                                  const uint32_t* counts, uint64_t* elapsed_ns) {
  std::vector<int> rcs(k, 0);
  std::vector<std::string> errs(k);
  std::vector<std::thread> threads;
  std::atomic<uint32_t> ready{0};
  std::atomic<bool> go{false};
  uint32_t off = 0;
  for (uint32_t j = 0; j < k; ++j) {
    TRITONBACKEND_Request** mine = reqs + off;
    off += counts[j];
    threads.emplace_back([&, j, mine] {
      ready.fetch_add(1);
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
      rcs[j] = ft_execute_sequence(insts[j], mine, counts[j], nullptr);
      if (rcs[j] != 0) errs[j] = g_ft_error;
    });
  }
  while (ready.load() < k) std::this_thread::yield();
  const auto t0 = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (std::thread& t : threads) t.join();
  if (elapsed_ns != nullptr)
    *elapsed_ns = static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
  for (uint32_t j = 0; j < k; ++j)
    if (rcs[j] != 0) {
      g_ft_error = errs[j];
      return rcs[j];
    }
  return 0;
}

int ft_request_released(TRITONBACKEND_Request* r) { return r->released; }
int ft_request_response_count(TRITONBACKEND_Request* r) { return static_cast<int>(r->responses.size()); }
static TRITONBACKEND_Response* last_response(TRITONBACKEND_Request* r) {
  return r->responses.empty() ? nullptr : r->responses.back().get();
}
int ft_response_sent(TRITONBACKEND_Request* r) {
  TRITONBACKEND_Response* p = last_response(r);
  return p ? p->sent : 0;
}
uint32_t ft_response_flags(TRITONBACKEND_Request* r) {
  TRITONBACKEND_Response* p = last_response(r);
  return p ? p->flags : 0;
}
int ft_response_error_code(TRITONBACKEND_Request* r) {  // -1: success
  TRITONBACKEND_Response* p = last_response(r);
  return (p && p->has_error) ? static_cast<int>(p->err_code) : -1;
}
const char* ft_response_error_message(TRITONBACKEND_Request* r) {
  TRITONBACKEND_Response* p = last_response(r);
  return (p && p->has_error) ? p->err_msg.c_str() : "";
}
int ft_response_output_count(TRITONBACKEND_Request* r) {
  TRITONBACKEND_Response* p = last_response(r);
  return p ? static_cast<int>(p->outputs.size()) : 0;
}
int ft_response_output(TRITONBACKEND_Request* r, int index, const char** name, int* dtype, const int64_t** shape,
                       uint32_t* dims, void** buffer, uint64_t* bytes, int* memory_type, int64_t* memory_type_id) {
  TRITONBACKEND_Response* p = last_response(r);
  if (p == nullptr || index < 0 || index >= static_cast<int>(p->outputs.size())) return 1;
  TRITONBACKEND_Output* o = p->outputs[index].get();
  *name = o->name.c_str();
  *dtype = static_cast<int>(o->dtype);
  *shape = o->shape.data();
  *dims = static_cast<uint32_t>(o->shape.size());
  *buffer = o->buffer;
  *bytes = o->bytes;
  *memory_type = static_cast<int>(o->mt);
  *memory_type_id = o->mt_id;
  return 0;
}
int ft_response_int_param(TRITONBACKEND_Request* r, const char* name, int64_t* value) {
  TRITONBACKEND_Response* p = last_response(r);
  if (p == nullptr) return 0;
  auto it = p->int_params.find(name);
  if (it == p->int_params.end()) return 0;
  *value = it->second;
  return 1;
}

}  // extern "C"
