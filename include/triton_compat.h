/*
 * triton_compat.h — the slice of the Triton Inference Server C API (tritonserver.h / tritonbackend.h)
 * that an `hps` backend needs, re-declared by hand because no Triton headers exist in this image.
 *
 * Everything here is the public, stable Triton ABI: opaque handle types, enum values and the
 * prototypes of the functions a backend IMPORTS from the server process.  The list is exactly the
 * set the reference glue uses (grep 'TRITON(BACKEND|SERVER)_\w+' over
 * /root/reference/hps_backend/{src,include}; call sites in SURVEY.md Appendix D).  When real Triton
 * headers are available, include them instead of this file (define HPS_USE_REAL_TRITON_HEADERS) —
 * nothing else in the backend changes.
 */
#ifndef HPS_TRITON_COMPAT_H_
#define HPS_TRITON_COMPAT_H_

#ifdef HPS_USE_REAL_TRITON_HEADERS
#include "triton/core/tritonbackend.h"
#include "triton/core/tritonserver.h"
#else

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HPS_TRITON_DECLSPEC __attribute__((__visibility__("default")))
#else
#define HPS_TRITON_DECLSPEC
#endif

/* The backend API version this backend is written against.  Load succeeds iff the server's major
 * equals this major and its minor is >= this minor (reference check: src/hps.cc:66-82). */
#define TRITONBACKEND_API_VERSION_MAJOR 1
#define TRITONBACKEND_API_VERSION_MINOR 10

/* ---- opaque handles ------------------------------------------------------------------------ */
struct TRITONSERVER_Error;
struct TRITONSERVER_Message;
struct TRITONSERVER_Server;
struct TRITONBACKEND_Backend;
struct TRITONBACKEND_Model;
struct TRITONBACKEND_ModelInstance;
struct TRITONBACKEND_Request;
struct TRITONBACKEND_Response;
struct TRITONBACKEND_Input;
struct TRITONBACKEND_Output;
#ifndef __cplusplus
typedef struct TRITONSERVER_Error TRITONSERVER_Error;
typedef struct TRITONSERVER_Message TRITONSERVER_Message;
typedef struct TRITONSERVER_Server TRITONSERVER_Server;
typedef struct TRITONBACKEND_Backend TRITONBACKEND_Backend;
typedef struct TRITONBACKEND_Model TRITONBACKEND_Model;
typedef struct TRITONBACKEND_ModelInstance TRITONBACKEND_ModelInstance;
typedef struct TRITONBACKEND_Request TRITONBACKEND_Request;
typedef struct TRITONBACKEND_Response TRITONBACKEND_Response;
typedef struct TRITONBACKEND_Input TRITONBACKEND_Input;
typedef struct TRITONBACKEND_Output TRITONBACKEND_Output;
#endif

/* ---- enums (values are ABI) ---------------------------------------------------------------- */
typedef enum TRITONSERVER_datatype_enum {
  TRITONSERVER_TYPE_INVALID = 0,
  TRITONSERVER_TYPE_BOOL = 1,
  TRITONSERVER_TYPE_UINT8 = 2,
  TRITONSERVER_TYPE_UINT16 = 3,
  TRITONSERVER_TYPE_UINT32 = 4,
  TRITONSERVER_TYPE_UINT64 = 5,
  TRITONSERVER_TYPE_INT8 = 6,
  TRITONSERVER_TYPE_INT16 = 7,
  TRITONSERVER_TYPE_INT32 = 8,
  TRITONSERVER_TYPE_INT64 = 9,
  TRITONSERVER_TYPE_FP16 = 10,
  TRITONSERVER_TYPE_FP32 = 11,
  TRITONSERVER_TYPE_FP64 = 12,
  TRITONSERVER_TYPE_BYTES = 13,
  TRITONSERVER_TYPE_BF16 = 14
} TRITONSERVER_DataType;

typedef enum TRITONSERVER_memorytype_enum {
  TRITONSERVER_MEMORY_CPU = 0,
  TRITONSERVER_MEMORY_CPU_PINNED = 1,
  TRITONSERVER_MEMORY_GPU = 2
} TRITONSERVER_MemoryType;

typedef enum TRITONSERVER_errorcode_enum {
  TRITONSERVER_ERROR_UNKNOWN = 0,
  TRITONSERVER_ERROR_INTERNAL = 1,
  TRITONSERVER_ERROR_NOT_FOUND = 2,
  TRITONSERVER_ERROR_INVALID_ARG = 3,
  TRITONSERVER_ERROR_UNAVAILABLE = 4,
  TRITONSERVER_ERROR_UNSUPPORTED = 5,
  TRITONSERVER_ERROR_ALREADY_EXISTS = 6
} TRITONSERVER_Error_Code;

typedef enum TRITONSERVER_loglevel_enum {
  TRITONSERVER_LOG_INFO = 0,
  TRITONSERVER_LOG_WARN = 1,
  TRITONSERVER_LOG_ERROR = 2,
  TRITONSERVER_LOG_VERBOSE = 3
} TRITONSERVER_LogLevel;

typedef enum TRITONSERVER_instancegroupkind_enum {
  TRITONSERVER_INSTANCEGROUPKIND_AUTO = 0,
  TRITONSERVER_INSTANCEGROUPKIND_CPU = 1,
  TRITONSERVER_INSTANCEGROUPKIND_GPU = 2,
  TRITONSERVER_INSTANCEGROUPKIND_MODEL = 3
} TRITONSERVER_InstanceGroupKind;

typedef enum TRITONBACKEND_artifacttype_enum {
  TRITONBACKEND_ARTIFACT_FILESYSTEM = 0
} TRITONBACKEND_ArtifactType;

typedef enum tritonserver_responsecompleteflag_enum {
  TRITONSERVER_RESPONSE_COMPLETE_FINAL = 1
} TRITONSERVER_ResponseCompleteFlag;

typedef enum tritonserver_requestreleaseflag_enum {
  TRITONSERVER_REQUEST_RELEASE_ALL = 1
} TRITONSERVER_RequestReleaseFlag;

/* ---- TRITONSERVER_* imports ------------------------------------------------------------------ */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONSERVER_ErrorNew(TRITONSERVER_Error_Code code, const char* msg);
HPS_TRITON_DECLSPEC void TRITONSERVER_ErrorDelete(TRITONSERVER_Error* error);
HPS_TRITON_DECLSPEC TRITONSERVER_Error_Code TRITONSERVER_ErrorCode(TRITONSERVER_Error* error);
HPS_TRITON_DECLSPEC const char* TRITONSERVER_ErrorCodeString(TRITONSERVER_Error* error);
HPS_TRITON_DECLSPEC const char* TRITONSERVER_ErrorMessage(TRITONSERVER_Error* error);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONSERVER_LogMessage(TRITONSERVER_LogLevel level, const char* filename,
                                                                const int line, const char* msg);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONSERVER_MessageSerializeToJson(TRITONSERVER_Message* message,
                                                                            const char** base, size_t* byte_size);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONSERVER_MessageDelete(TRITONSERVER_Message* message);
HPS_TRITON_DECLSPEC const char* TRITONSERVER_DataTypeString(TRITONSERVER_DataType datatype);

/* ---- TRITONBACKEND_* imports: backend ---------------------------------------------------------- */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ApiVersion(uint32_t* major, uint32_t* minor);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_BackendName(TRITONBACKEND_Backend* backend, const char** name);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_BackendConfig(TRITONBACKEND_Backend* backend,
                                                                    TRITONSERVER_Message** backend_config);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_BackendArtifacts(TRITONBACKEND_Backend* backend,
                                                                       TRITONBACKEND_ArtifactType* artifact_type,
                                                                       const char** location);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_BackendState(TRITONBACKEND_Backend* backend, void** state);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_BackendSetState(TRITONBACKEND_Backend* backend, void* state);

/* ---- model ----------------------------------------------------------------------------------- */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelName(TRITONBACKEND_Model* model, const char** name);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelVersion(TRITONBACKEND_Model* model, uint64_t* version);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelRepository(TRITONBACKEND_Model* model,
                                                                      TRITONBACKEND_ArtifactType* artifact_type,
                                                                      const char** location);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelConfig(TRITONBACKEND_Model* model,
                                                                  const uint32_t config_version,
                                                                  TRITONSERVER_Message** model_config);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelServer(TRITONBACKEND_Model* model,
                                                                  TRITONSERVER_Server** server);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelBackend(TRITONBACKEND_Model* model,
                                                                   TRITONBACKEND_Backend** backend);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelState(TRITONBACKEND_Model* model, void** state);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelSetState(TRITONBACKEND_Model* model, void* state);

/* ---- model instance ---------------------------------------------------------------------------- */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceName(TRITONBACKEND_ModelInstance* instance,
                                                                        const char** name);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceKind(TRITONBACKEND_ModelInstance* instance,
                                                                        TRITONSERVER_InstanceGroupKind* kind);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceDeviceId(TRITONBACKEND_ModelInstance* instance,
                                                                            int32_t* device_id);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceModel(TRITONBACKEND_ModelInstance* instance,
                                                                         TRITONBACKEND_Model** model);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceState(TRITONBACKEND_ModelInstance* instance,
                                                                         void** state);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceSetState(TRITONBACKEND_ModelInstance* instance,
                                                                            void* state);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceReportStatistics(
    TRITONBACKEND_ModelInstance* instance, TRITONBACKEND_Request* request, const bool success,
    const uint64_t exec_start_ns, const uint64_t compute_start_ns, const uint64_t compute_end_ns,
    const uint64_t exec_end_ns);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ModelInstanceReportBatchStatistics(
    TRITONBACKEND_ModelInstance* instance, const uint64_t batch_size, const uint64_t exec_start_ns,
    const uint64_t compute_start_ns, const uint64_t compute_end_ns, const uint64_t exec_end_ns);

/* ---- request / input --------------------------------------------------------------------------- */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestId(TRITONBACKEND_Request* request, const char** id);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestCorrelationId(TRITONBACKEND_Request* request,
                                                                           uint64_t* id);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestInputCount(TRITONBACKEND_Request* request,
                                                                        uint32_t* count);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestInputName(TRITONBACKEND_Request* request,
                                                                       const uint32_t index,
                                                                       const char** input_name);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestInput(TRITONBACKEND_Request* request, const char* name,
                                                                   TRITONBACKEND_Input** input);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestOutputCount(TRITONBACKEND_Request* request,
                                                                         uint32_t* count);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestOutputName(TRITONBACKEND_Request* request,
                                                                        const uint32_t index,
                                                                        const char** output_name);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_RequestRelease(TRITONBACKEND_Request* request,
                                                                     uint32_t release_flags);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_InputProperties(TRITONBACKEND_Input* input, const char** name,
                                                                      TRITONSERVER_DataType* datatype,
                                                                      const int64_t** shape, uint32_t* dims_count,
                                                                      uint64_t* byte_size, uint32_t* buffer_count);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_InputBuffer(TRITONBACKEND_Input* input, const uint32_t index,
                                                                  const void** buffer, uint64_t* buffer_byte_size,
                                                                  TRITONSERVER_MemoryType* memory_type,
                                                                  int64_t* memory_type_id);

/* ---- response / output ------------------------------------------------------------------------- */
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ResponseNew(TRITONBACKEND_Response** response,
                                                                  TRITONBACKEND_Request* request);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ResponseOutput(TRITONBACKEND_Response* response,
                                                                     TRITONBACKEND_Output** output, const char* name,
                                                                     const TRITONSERVER_DataType datatype,
                                                                     const int64_t* shape, const uint32_t dims_count);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_OutputBuffer(TRITONBACKEND_Output* output, void** buffer,
                                                                   const uint64_t buffer_byte_size,
                                                                   TRITONSERVER_MemoryType* memory_type,
                                                                   int64_t* memory_type_id);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ResponseSetIntParameter(TRITONBACKEND_Response* response,
                                                                              const char* name, const int64_t value);
HPS_TRITON_DECLSPEC TRITONSERVER_Error* TRITONBACKEND_ResponseSend(TRITONBACKEND_Response* response,
                                                                   const uint32_t send_flags,
                                                                   TRITONSERVER_Error* error);

#ifdef __cplusplus
}
#endif
#endif /* HPS_USE_REAL_TRITON_HEADERS */
#endif /* HPS_TRITON_COMPAT_H_ */
