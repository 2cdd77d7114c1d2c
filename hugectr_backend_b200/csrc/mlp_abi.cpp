// C ABI of the dense MLP head (include/hpsx.h: hpsx_mlp_*) over dense_mlp.cu.
#include "dense_mlp.h"
#include "engine_internal.hpp"

using namespace hpsx::eng;

extern "C" {

// ------------------------------------------------------------------------------------------------
// dense MLP head
// ------------------------------------------------------------------------------------------------
struct hpsx_mlp {
  hpsx::DenseMlp* impl = nullptr;
  int device = 0;
};

int hpsx_mlp_create(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                    const float* const* biases, const int* relu, hpsx_mlp** out) {
  return hpsx_mlp_create_ex(device, num_layers, dims, weights, biases, relu, HPSX_MLP_BF16, out);
}

int hpsx_mlp_create_ex(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                       const float* const* biases, const int* relu, int precision, hpsx_mlp** out) {
  HPSX_GUARD_BEGIN
  if (!out) return fail(HPSX_ERR_INVALID_ARG, "null output handle");
  if (precision != HPSX_MLP_BF16 && precision != HPSX_MLP_TF32) return fail(HPSX_ERR_INVALID_ARG, "unknown MLP precision");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  hpsx::DenseMlp* impl = nullptr;
  const cudaError_t e = hpsx::mlp_create(device, num_layers, dims, weights, biases, relu, &impl, precision);
  if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? HPSX_ERR_INVALID_ARG : HPSX_ERR_CUDA, hpsx::mlp_last_error());
  hpsx_mlp* m = new hpsx_mlp();
  m->impl = impl;
  m->device = device;
  *out = m;
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_mlp_forward(hpsx_mlp* m, const float* d_in, size_t batch, float* d_out, void* stream) {
  HPSX_GUARD_BEGIN
  if (!m) return fail(HPSX_ERR_INVALID_ARG, "null mlp");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  const cudaError_t e = hpsx::mlp_forward(m->impl, d_in, batch, d_out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? HPSX_ERR_INVALID_ARG : HPSX_ERR_CUDA, hpsx::mlp_last_error());
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_mlp_forward_bf16(hpsx_mlp* m, const void* d_in_bf16, size_t batch, float* d_out, void* stream) {
  HPSX_GUARD_BEGIN
  if (!m) return fail(HPSX_ERR_INVALID_ARG, "null mlp");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(HPSX_ERR_CUDA, "cudaSetDevice failed");
  const cudaError_t e = hpsx::mlp_forward(m->impl, nullptr, batch, d_out, static_cast<cudaStream_t>(stream), d_in_bf16);
  if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? HPSX_ERR_INVALID_ARG : HPSX_ERR_CUDA, hpsx::mlp_last_error());
  return HPSX_OK;
  HPSX_GUARD_END
}

int hpsx_mlp_destroy(hpsx_mlp* m) {
  HPSX_GUARD_BEGIN
  if (!m) return HPSX_OK;
  DeviceGuard guard(m->device);
  hpsx::mlp_destroy(m->impl);
  delete m;
  return HPSX_OK;
  HPSX_GUARD_END
}

}  // extern "C"
