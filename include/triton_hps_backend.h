/*
 * triton_hps_backend.h — the C ABI that libtriton_hps.so EXPORTS: the drop-in boundary.
 *
 * Triton dlopen()s <backend-directory>/hps/libtriton_hps.so (reference: hps_backend/CMakeLists.txt:154,171,
 * README.md:101-109) for models whose config.pbtxt says `backend: "hps"` and binds exactly these seven
 * symbols; nothing else is exported (reference version script: hps_backend/src/libtriton_hps.ldscript:26-30,
 * ours: hugectr_backend_b200/csrc/libtriton_hps.ldscript).  Each entry point cites the reference
 * function it replaces (paths relative to /root/reference/hps_backend).
 *
 * All functions return nullptr on success or a TRITONSERVER_Error* that Triton owns.
 * Types come from tritonserver.h / tritonbackend.h (re-declared in triton_compat.h for this image).
 */
#ifndef TRITON_HPS_BACKEND_H_
#define TRITON_HPS_BACKEND_H_

#include "triton_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* replaces src/hps.cc:57-136.  Checks the backend API version (major equal, minor >=), reads
 * {"cmdline":{"ps":"<path to ps.json>"}} from the backend config (tritonserver
 * --backend-config=hps,ps=<file>, README.md:108), creates the process-wide parameter server
 * (loads every model's sparse files, builds the HBM caches) and attaches it as backend state. */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_Initialize(TRITONBACKEND_Backend* backend);

/* replaces src/hps.cc:142-155.  Destroys the parameter server and every cache. */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_Finalize(TRITONBACKEND_Backend* backend);

/* replaces src/hps.cc:162-246 (+ src/model_state.cpp:66-106,180-432).  Re-reads ps.json when the model
 * is not known yet (online deployment), validates config.pbtxt (2 inputs KEYS:INT64 / NUMKEYS:INT32
 * with dims[0] == -1, 1 FP32 output with dims[0] == -1), parses instance_group / parameters, and makes
 * sure an embedding cache exists on every GPU of the instance group (each must be in
 * deployed_device_list). */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInitialize(TRITONBACKEND_Model* model);

/* replaces src/hps.cc:252-274 (+ src/model_state.cpp:108-122). */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelFinalize(TRITONBACKEND_Model* model);

/* replaces src/hps.cc:280-324 (+ src/model_instance_state.cpp:73-174).  One lookup session (stream +
 * pinned/device workspaces) per instance; instances of one model on one device share its cache. */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceInitialize(TRITONBACKEND_ModelInstance* instance);

/* replaces src/hps.cc:330-344. */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceFinalize(TRITONBACKEND_ModelInstance* instance);

/* replaces src/hps.cc:348-788 (+ src/model_instance_state.cpp:176-197).  Per request: KEYS int64
 * table-major + NUMKEYS int32 [1,T] -> one FP32 output [sum_t NUMKEYS[t] * vecsize[t]] =
 * concat_t rows_t(KEYS_t), written by the lookup kernels straight into the Triton output buffer;
 * response parameters NumSample and DeviceID; exactly one FINAL response and one release per
 * request; per-request errors (oversize batch, malformed NUMKEYS, CUDA failures) become error
 * responses instead of crashes / C++ exceptions. */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceExecute(TRITONBACKEND_ModelInstance* instance,
                                                                           TRITONBACKEND_Request** requests,
                                                                           const uint32_t request_count);

#ifdef __cplusplus
}
#endif
#endif /* TRITON_HPS_BACKEND_H_ */
