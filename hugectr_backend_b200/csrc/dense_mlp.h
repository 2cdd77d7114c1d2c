// Dense MLP head on the tcgen05 tensor cores (dense_mlp.cu); host C++ only sees this header.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace hpsx {

struct DenseMlp;

// dims[L+1]: layer l maps dims[l] -> dims[l+1]; weights[l] is host fp32 [dims[l+1], dims[l]] row-major (the
// layout of a Keras Dense kernel transposed / a torch Linear weight); biases[l] host fp32 [dims[l+1]] or nullptr;
// relu[l] != 0 applies max(x, 0).  Every dims[l] must be a multiple of 8.
// precision 0: bf16 operands; 1: TF32 — weights and activations stay fp32 in memory, the input is read in place.
cudaError_t mlp_create(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                       const float* const* biases, const int* relu, DenseMlp** out, int precision = 0);
void mlp_destroy(DenseMlp* m);
// d_in: device fp32 [batch, dims[0]] (e.g. the lookup's OUTPUT0 viewed as [batch, slots * dim]); d_out: device fp32
// [batch, dims[L]].  Asynchronous on `stream`.
cudaError_t mlp_forward(DenseMlp* m, const float* d_in, size_t batch, float* d_out, cudaStream_t stream,
                        const void* d_in_bf16 = nullptr);
// With d_in_bf16 (device bf16 [batch, dims[0]], 16-B aligned: the lookup's bf16 mirror) the conversion pass is
// skipped and d_in is ignored.
const char* mlp_last_error();

}  // namespace hpsx
