// Helpers shared by the translation units of libhpsx.so (hpsx.cpp: server, cache, session, lookups;
// shard_group.cpp: model-parallel groups; mlp_abi.cpp: dense head).  Not part of any ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <mutex>
#include <shared_mutex>
#include <string>

#include "engine.hpp"

namespace hpsx {
namespace eng {

// thread-local text behind hpsx_last_error()
extern thread_local std::string g_err;
int fail(int code, std::string msg);

double now_ms();
bool trace_on();  // HPSX_TRACE=1: per-call phase timings on stderr

// Is the primary context of `dev` alive in this process?  (A thread that never selected a device reports device 0;
// switching "back" to it would CREATE a context on GPU 0 — ~0.4 s and some HBM — in a process that only serves
// another GPU, e.g. one rank of a one-process-per-GPU deployment calling from a worker thread.)
bool primary_context_active(int dev);

// NVTX range around the phases of a lookup (the reference marks the same places: hps_backend/src/hps.cc:375,671,
// 674,701, src/model_instance_state.cpp:179; opt-in there, free here when no tool is attached).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev && primary_context_active(prev)) cudaSetDevice(prev);
  }
};

// `mb` (nullable) replaces the session's own miss list (model-parallel groups keep a larger one).
struct MissBufs {
  const int64_t* h_keys;
  const int64_t* d_keys;
  const uint32_t* d_pos;
};

// hpsx.cpp
int stream_miss_rows(hpsx_session* s, size_t t, size_t key_off, uint32_t m, float* d_out, bool insert, uint32_t epoch,
                     float* d_all_stage, std::unique_lock<std::shared_mutex>* wlock, const MissBufs* mb = nullptr);
int ensure_pool_stage(hpsx_session* s, size_t m);
int ensure_sort_workspace(hpsx_session* s);
size_t pull_sort_min();
int sync_direct_pull_index(hpsx_cache* c, cudaStream_t stream);

// peer_tier.cpp — callers hold c->async_mu, c->pull_rw and c->rw exclusively (or the cache is not published yet)
int tier_build_locked(hpsx_cache* c, uint32_t rank, uint32_t world);
int tier_attach_local_locked(hpsx_cache* c, size_t table, uint32_t peer, hpsx_cache* peer_cache);
int tier_commit_locked(hpsx_cache* c);
void tier_release(hpsx_cache* c);  // unmaps the peers' shards and frees the own ones
// every cache of `m` in this process becomes one rank of a tier (ranks in ascending device order)
int tier_connect_local_locked(hpsx::Model* m, const std::vector<hpsx_cache*>& caches);

}  // namespace eng
}  // namespace hpsx

#define HPSX_CU(call)                                                                                   \
  do {                                                                                                  \
    const cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                             \
      return ::hpsx::eng::fail(HPSX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)

#define HPSX_GUARD_BEGIN try {
#define HPSX_GUARD_END                                                   \
  }                                                                      \
  catch (const std::bad_alloc&) {                                        \
    return ::hpsx::eng::fail(HPSX_ERR_INTERNAL, "out of host memory");   \
  }                                                                      \
  catch (const std::exception& e) {                                      \
    return ::hpsx::eng::fail(HPSX_ERR_INTERNAL, e.what());               \
  }
