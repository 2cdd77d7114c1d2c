// Dense MLP head fed by the embedding lookup: the step that follows the HPS model in the
// reference's ensembles (samples/hps-triton-ensemble/01_model_training.ipynb cells 7,11: fc_1 -> fc_2 -> fc_3 on the
// reshaped LOOKUP_VECTORS; 02_model_inference_hps_tf_ensemble.ipynb:336-395).  It is the only true contraction next to
// the lookup path, so it runs on the 5th-generation tensor cores:
//
//   Y[M,N] = act(X[M,K] . W[N,K]^T + b)      bf16 operands (default) or TF32 operands, fp32 accumulation in TMEM
//
// TF32 mode (mlp_create precision 1; the reference's dense model is fp32, 01_model_training.ipynb cells 7,11): X and W stay
// fp32 in memory — the lookup's OUTPUT0 is read by TMA as it is, no conversion pass and no bf16 mirror — the tensor
// cores round the operands to TF32 (10-bit mantissa) and accumulate in fp32, hidden activations are stored as fp32.
// Same tile shape: a 128-byte swizzle row holds 32 fp32 instead of 64 bf16, K per instruction is 8 instead of 16.
//
//   warp 0  (one elected lane)  TMA producer: cp.async.bulk.tensor 2-D tiles of X and W into a 128B-swizzled
//                               shared-memory ring, completion on mbarriers
//   warp 1  (one elected lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, 128 x 256 x 16 per instruction,
//                               accumulator in 256 TMEM columns; tcgen05.commit releases ring slots / signals the epilogue
//   warps 2-5                   epilogue: tcgen05.ld (32 lanes x 32 columns per warp and step) -> bias + ReLU ->
//                               bf16 (hidden layers) or fp32 (last GEMM layer) -> global
//
// Persistent kernel, one CTA per SM: a 4-stage ring of 48 KB stages (192 KB of shared memory) and two accumulators
// (all 512 TMEM columns), so the epilogue of tile i overlaps the MMAs of tile i+1.  128 x 256 tiles halve the operand
// bytes per flop of 128 x 128 ones (the first version: L2-bandwidth bound at 0.49 of the tensor peak).  A final layer
// with N == 1 (the samples' logit) is a dot product per row and runs as a SIMT kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dense_mlp.h"

namespace hpsx {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;  // one tcgen05.mma covers the whole 128 x 256 tile: half the operand bytes per flop of 128 x 128
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle row
// K per tcgen05.mma: 16 for 16-bit operands, 8 for TF32 — 32 bytes of a swizzle row either way (kUmmaKBytes)
constexpr int kBlockK32 = 32; // TF32 mode: 32 fp32 = 128 B = one swizzle row (same tile bytes)
constexpr int kUmmaKBytes = 32;  // K bytes per tcgen05.mma inside a swizzle row: 16 x 2 B = 8 x 4 B
constexpr int kStages = 4;
constexpr int kAccStages = 2;  // accumulators in TMEM: the epilogue of tile i overlaps the MMAs of tile i+1
constexpr int kThreads = 192;  // 6 warps
constexpr uint32_t kTileABytes = kBlockM * kBlockK * 2;
constexpr uint32_t kTileBBytes = kBlockN * kBlockK * 2;
constexpr uint32_t kStageBytes = kTileABytes + kTileBBytes;
constexpr uint32_t kTmemCols = kAccStages * kBlockN;  // 512: all of an SM's tensor memory, one CTA per SM
constexpr size_t kSmemBytes = 1024 /* alignment slack */ + kStages * kStageBytes + 256 /* barriers + tmem slot */;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MLP_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MLP_DONE_%=;\n"
      "bra MLP_WAIT_%=;\n"
      "MLP_DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_addr(smem_dst)),
      "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major operand tile [rows][64 bf16], 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= static_cast<uint64_t>(0) << 16;                        // leading byte offset: unused for swizzled K-major
  d |= static_cast<uint64_t>(1024u >> 4) << 32;               // stride byte offset (8 rows x 128 B), bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                        // layout: SWIZZLE_128B
  return d;
}

// tcgen05 instruction descriptor: D fp32, A/B bf16 (format 1; kind::f16) or TF32 (format 2; kind::tf32), both K-major.
__device__ __forceinline__ uint32_t make_instr_desc(uint32_t m, uint32_t n, bool tf32) {
  const uint32_t fmt = tf32 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 -> nearest TF32 value (10 mantissa bits, ties away from zero), still stored as fp32.  The tensor cores TRUNCATE
// the low 13 bits of an fp32 operand (measured: a three-layer linear model drifts by 1.6e-3 of its output scale, a
// systematic shrink); operands rounded here first carry an unbiased error of half the size.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__global__ void __launch_bounds__(256) round_tf32_kernel(float* w, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) w[i] = round_tf32(w[i]);
}

struct GemmArgs {
  int round_out;  // TF32 mode: the output feeds another TF32 GEMM, store it rounded to TF32
  int M, N, K;
  const float* bias;  // [N] or nullptr
  int relu;
  __nv_bfloat16* out_bf16;  // [M, N] when the next layer is another GEMM
  float* out_f32;           // [M, N] otherwise
};

// Persistent: CTA b works on tiles b, b + gridDim.x, ... (n fastest, so concurrently running CTAs share X rows in L2).
template <bool kTf32>
__global__ void __launch_bounds__(kThreads, 1)
mlp_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const GemmArgs g) {
  constexpr int kBK = kTf32 ? kBlockK32 : kBlockK;  // elements of K per ring stage (128 bytes either way)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full_bar = empty_bar + kStages;
  uint64_t* acc_empty_bar = acc_full_bar + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + kAccStages);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const int num_k_blocks = (g.K + kBK - 1) / kBK;
  const int num_n = (g.N + kBlockN - 1) / kBlockN;
  const int num_m = (g.M + kBlockM - 1) / kBlockM;
  const int num_tiles = num_n * num_m;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full_bar[s], 1);
      mbar_init(&acc_empty_bar[s], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // one warp allocates the accumulator columns and owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;  // k-blocks issued so far, over all tiles of this CTA
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * kBlockM, n0 = (tile % num_n) * kBlockN;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const uint32_t s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);  // passes immediately on the first round
          unsigned char* a_tile = tiles + s * kStageBytes;
          unsigned char* b_tile = a_tile + kTileABytes;
          mbar_expect_tx(&full_bar[s], kStageBytes);
          tma_load_2d(a_tile, &map_x, kb * kBK, m0, &full_bar[s]);
          tma_load_2d(b_tile, &map_w, kb * kBK, n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_instr_desc(kBlockM, kBlockN, kTf32);
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t % kAccStages;
        const uint32_t acc_ph = (t / kAccStages) & 1u;
        mbar_wait(&acc_empty_bar[acc], acc_ph ^ 1u);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const uint32_t s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_addr(tiles + s * kStageBytes);
          const uint32_t b_addr = a_addr + kTileABytes;
#pragma unroll
          for (int k = 0; k < 128 / kUmmaKBytes; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * kUmmaKBytes);
            const uint64_t db = make_smem_desc(b_addr + k * kUmmaKBytes);
            if constexpr (kTf32)
              umma_tf32(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // the ring slot is free once these MMAs have read it
        }
        umma_commit(&acc_full_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const uint32_t quad = warp & 3u;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m0 = (tile / num_n) * kBlockM, n0 = (tile % num_n) * kBlockN;
      const uint32_t acc = t % kAccStages;
      const uint32_t acc_ph = (t / kAccStages) & 1u;
      mbar_wait(&acc_full_bar[acc], acc_ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = m0 + static_cast<int>(quad * 32u + lane);
#pragma unroll 1
      for (int c = 0; c < kBlockN; c += 32) {
        if (n0 + c >= g.N) break;  // warp-uniform: nothing of this column block is inside the matrix
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((quad * 32u) << 16) + acc * kBlockN + static_cast<uint32_t>(c), v);
        if (m < g.M) {
          const int n_base = n0 + c;
          if (g.out_bf16 != nullptr) {
            __nv_bfloat16* dst = g.out_bf16 + static_cast<size_t>(m) * g.N + n_base;
            if (n_base + 32 <= g.N && (g.N & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint32_t packed[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float x0 = __uint_as_float(v[j + 2 * q]), x1 = __uint_as_float(v[j + 2 * q + 1]);
                  if (g.bias != nullptr) {
                    x0 += __ldg(g.bias + n_base + j + 2 * q);
                    x1 += __ldg(g.bias + n_base + j + 2 * q + 1);
                  }
                  if (g.relu) {
                    x0 = fmaxf(x0, 0.f);
                    x1 = fmaxf(x1, 0.f);
                  }
                  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                  packed[q] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (n_base + j < g.N) {
                  float x = __uint_as_float(v[j]);
                  if (g.bias != nullptr) x += __ldg(g.bias + n_base + j);
                  if (g.relu) x = fmaxf(x, 0.f);
                  dst[j] = __float2bfloat16_rn(x);
                }
              }
            }
          } else {
            float* dst = g.out_f32 + static_cast<size_t>(m) * g.N + n_base;
            if (n_base + 32 <= g.N && (g.N & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  x[q] = __uint_as_float(v[j + q]);
                  if (g.bias != nullptr) x[q] += __ldg(g.bias + n_base + j + q);
                  if (g.relu) x[q] = fmaxf(x[q], 0.f);
                  if (kTf32 && g.round_out) x[q] = round_tf32(x[q]);
                }
                *reinterpret_cast<float4*>(dst + j) = make_float4(x[0], x[1], x[2], x[3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (n_base + j < g.N) {
                  float x = __uint_as_float(v[j]);
                  if (g.bias != nullptr) x += __ldg(g.bias + n_base + j);
                  if (g.relu) x = fmaxf(x, 0.f);
                  if (kTf32 && g.round_out) x = round_tf32(x);
                  dst[j] = x;
                }
              }
            }
          }
        }
      }
      // this warp has read its share of the accumulator: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
  const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
  reinterpret_cast<uint4*>(out)[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
}
__global__ void to_bf16_tail_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t begin, size_t n) {
  const size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

// Layer with a single output unit (the samples' logit): one warp per row, fp32 accumulation.
__global__ void __launch_bounds__(256) mlp_dot_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                      const float* __restrict__ bias, int M, int K, int relu,
                                                      float* __restrict__ out) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc += __bfloat162float(xr[k]) * __bfloat162float(w[k]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    if (bias != nullptr) acc += bias[0];
    if (relu) acc = fmaxf(acc, 0.f);
    out[row] = acc;
  }
}

__global__ void __launch_bounds__(256) mlp_dot_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int M, int K, int relu,
                                                          float* __restrict__ out) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<size_t>(row) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc += xr[k] * w[k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    if (bias != nullptr) acc += bias[0];
    if (relu) acc = fmaxf(acc, 0.f);
    out[row] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 (or, TF32 mode, fp32) matrix [rows, cols] row-major (cols contiguous): box = 128 bytes of columns x box_rows rows,
// 128-byte swizzle.
bool make_map(CUtensorMap* map, const void* base, size_t rows, size_t cols, uint32_t box_rows, bool f32 = false) {
  EncodeTiledFn fn = encode_tiled();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * (f32 ? sizeof(float) : sizeof(__nv_bfloat16))};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(f32 ? kBlockK32 : kBlockK), box_rows};
  const cuuint32_t elem[2] = {1, 1};
  return fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

struct DenseMlp {
  int device = 0;
  std::vector<size_t> dims;  // [L+1]
  int tf32 = 0;                   // precision 1: fp32 weights and activations, TF32 tensor-core arithmetic
  std::vector<float*> w32;        // [L] device, [out, in] fp32 (TF32 mode)
  float* act32[2] = {nullptr, nullptr};
  std::vector<__nv_bfloat16*> w;  // [L] device, [out, in] bf16
  std::vector<float*> b;          // [L] device or nullptr
  std::vector<int> relu;
  // activations: ping-pong bf16 buffers sized on first use
  __nv_bfloat16* act[2] = {nullptr, nullptr};
  size_t act_rows = 0;
  size_t max_dim = 0;
  int num_sms = 148;
};

static thread_local std::string g_mlp_err;
const char* mlp_last_error() { return g_mlp_err.c_str(); }

static cudaError_t fail_cuda(cudaError_t e, const char* what) {
  g_mlp_err = std::string(what) + ": " + cudaGetErrorString(e);
  return e;
}

cudaError_t mlp_create(int device, size_t num_layers, const size_t* dims, const float* const* weights,
                       const float* const* biases, const int* relu, DenseMlp** out, int precision) {
  if (num_layers == 0 || !dims || !weights || !out) {
    g_mlp_err = "null argument";
    return cudaErrorInvalidValue;
  }
  for (size_t l = 0; l < num_layers; ++l)
    if (dims[l] == 0 || dims[l + 1] == 0 || dims[l] % 8 != 0) {
      g_mlp_err = "every layer input width must be a positive multiple of 8 (16-byte rows for the TMA tensor maps)";
      return cudaErrorInvalidValue;
    }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(mlp_gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp_gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
    attr_set = true;
  }
  DenseMlp* m = new DenseMlp();
  m->device = device;
  m->tf32 = precision == 1 ? 1 : 0;
  cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (m->num_sms <= 0) m->num_sms = 148;
  m->dims.assign(dims, dims + num_layers + 1);
  for (size_t d : m->dims) m->max_dim = d > m->max_dim ? d : m->max_dim;
  for (size_t l = 0; l < num_layers; ++l) {
    const size_t n = dims[l + 1], k = dims[l];
    float* tmp = nullptr;
    __nv_bfloat16* w = nullptr;
    e = cudaMalloc(&tmp, n * k * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(tmp, weights[l], n * k * sizeof(float), cudaMemcpyHostToDevice);
    if (m->tf32) {
      if (e == cudaSuccess && n > 1) {  // GEMM layers: weights rounded to TF32 once, here (the dot layer keeps fp32 weights)
        round_tf32_kernel<<<static_cast<unsigned>((n * k + 255) / 256), 256>>>(tmp, n * k);
        e = cudaDeviceSynchronize();
      }
      m->w32.push_back(tmp);  // kept as fp32
      m->w.push_back(nullptr);
    } else {
      if (e == cudaSuccess) e = cudaMalloc(&w, n * k * sizeof(__nv_bfloat16));
      if (e == cudaSuccess) {
        const size_t n8 = n * k / 8;
        to_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256>>>(tmp, w, n8);
        e = cudaDeviceSynchronize();
      }
      cudaFree(tmp);
      m->w.push_back(w);
      m->w32.push_back(nullptr);
    }
    float* b = nullptr;
    if (e == cudaSuccess && biases && biases[l]) {
      e = cudaMalloc(&b, n * sizeof(float));
      if (e == cudaSuccess) e = cudaMemcpy(b, biases[l], n * sizeof(float), cudaMemcpyHostToDevice);
    }
    m->b.push_back(b);
    m->relu.push_back(relu ? relu[l] : 0);
    if (e != cudaSuccess) {
      fail_cuda(e, "uploading layer weights");
      mlp_destroy(m);
      return e;
    }
  }
  *out = m;
  return cudaSuccess;
}

void mlp_destroy(DenseMlp* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (auto* p : m->w) cudaFree(p);
  for (auto* p : m->w32) cudaFree(p);
  for (auto* p : m->b) cudaFree(p);
  cudaFree(m->act[0]);
  cudaFree(m->act[1]);
  cudaFree(m->act32[0]);
  cudaFree(m->act32[1]);
  delete m;
}

cudaError_t mlp_forward(DenseMlp* m, const float* d_in, size_t batch, float* d_out, cudaStream_t stream,
                        const void* d_in_bf16) {
  if (!m || (!d_in && !d_in_bf16) || !d_out) {
    g_mlp_err = "null argument";
    return cudaErrorInvalidValue;
  }
  if (batch == 0) return cudaSuccess;
  cudaError_t e = cudaSetDevice(m->device);
  if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
  const size_t L = m->w.size();
  if (m->tf32) {
    // TF32 mode: fp32 end to end in memory; layer 0 reads d_in (the lookup's OUTPUT0) in place
    if (d_in == nullptr) {
      g_mlp_err = "a TF32 head takes the fp32 vectors (hpsx_mlp_forward), not the bf16 mirror";
      return cudaErrorInvalidValue;
    }
    size_t hidden = 0;
    for (size_t l = 1; l < L; ++l) hidden = m->dims[l] > hidden ? m->dims[l] : hidden;
    if (m->act_rows < batch && hidden != 0) {
      for (int i = 0; i < 2; ++i) {
        cudaFree(m->act32[i]);
        m->act32[i] = nullptr;
        e = cudaMalloc(&m->act32[i], batch * hidden * sizeof(float));
        if (e != cudaSuccess) return fail_cuda(e, "allocating activations");
      }
      m->act_rows = batch;
    }
    int cur32 = 0;
    const float* x_in = d_in;
    for (size_t l = 0; l < L; ++l) {
      const size_t K = m->dims[l], N = m->dims[l + 1];
      const bool last = l + 1 == L;
      if (N == 1) {
        if (!last) {
          g_mlp_err = "a layer with one output unit must be the last layer";
          return cudaErrorInvalidValue;
        }
        mlp_dot_f32_kernel<<<static_cast<unsigned>((batch * 32 + 255) / 256), 256, 0, stream>>>(
            x_in, m->w32[l], m->b[l], static_cast<int>(batch), static_cast<int>(K), m->relu[l], d_out);
        continue;
      }
      if ((reinterpret_cast<uintptr_t>(x_in) & 15u) != 0) {
        g_mlp_err = "the fp32 input must be 16-byte aligned";
        return cudaErrorInvalidValue;
      }
      CUtensorMap map_x, map_w;
      if (!make_map(&map_x, x_in, batch, K, kBlockM, true) || !make_map(&map_w, m->w32[l], N, K, kBlockN, true)) {
        g_mlp_err = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
      }
      GemmArgs g{};
      g.M = static_cast<int>(batch);
      g.N = static_cast<int>(N);
      g.K = static_cast<int>(K);
      g.bias = m->b[l];
      g.relu = m->relu[l];
      g.out_bf16 = nullptr;
      g.out_f32 = last ? d_out : m->act32[cur32];
      g.round_out = (!last && m->dims[l + 2] > 1) ? 1 : 0;
      const size_t tiles = ((N + kBlockN - 1) / kBlockN) * ((batch + kBlockM - 1) / kBlockM);
      const unsigned grid = static_cast<unsigned>(tiles < static_cast<size_t>(m->num_sms) ? tiles : m->num_sms);
      mlp_gemm_tcgen05_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(map_x, map_w, g);
      x_in = g.out_f32;
      cur32 ^= 1;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "launching the MLP kernels");
    return cudaSuccess;
  }
  if (m->act_rows < batch) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(m->act[i]);
      m->act[i] = nullptr;
      e = cudaMalloc(&m->act[i], batch * m->max_dim * sizeof(__nv_bfloat16));
      if (e != cudaSuccess) return fail_cuda(e, "allocating activations");
    }
    m->act_rows = batch;
  }
  // input fp32 (the lookup's output) -> bf16, unless the lookup already wrote a bf16 mirror
  if (d_in_bf16 == nullptr) {
    const size_t n = batch * m->dims[0], n8 = n / 8;
    if (n8) to_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(d_in, m->act[0], n8);
    if (n8 * 8 < n) to_bf16_tail_kernel<<<1, 8, 0, stream>>>(d_in, m->act[0], n8 * 8, n);
  } else if ((reinterpret_cast<uintptr_t>(d_in_bf16) & 15u) != 0) {
    g_mlp_err = "the bf16 input must be 16-byte aligned";
    return cudaErrorInvalidValue;
  }
  int cur = 0;
  for (size_t l = 0; l < L; ++l) {
    const __nv_bfloat16* x_in = (l == 0 && d_in_bf16 != nullptr) ? static_cast<const __nv_bfloat16*>(d_in_bf16) : m->act[cur];
    const size_t K = m->dims[l], N = m->dims[l + 1];
    const bool last = l + 1 == L;
    if (N == 1) {
      // single output unit: dot product per row; its result is fp32 whatever follows
      float* dst = last ? d_out : nullptr;
      if (!last) {
        g_mlp_err = "a layer with one output unit must be the last layer";
        return cudaErrorInvalidValue;
      }
      mlp_dot_kernel<<<static_cast<unsigned>((batch * 32 + 255) / 256), 256, 0, stream>>>(x_in, m->w[l], m->b[l],
                                                                                          static_cast<int>(batch), static_cast<int>(K),
                                                                                          m->relu[l], dst);
      continue;
    }
    CUtensorMap map_x, map_w;
    if (!make_map(&map_x, x_in, batch, K, kBlockM) || !make_map(&map_w, m->w[l], N, K, kBlockN)) {
      g_mlp_err = "cuTensorMapEncodeTiled failed";
      return cudaErrorInvalidValue;
    }
    GemmArgs g{};
    g.M = static_cast<int>(batch);
    g.N = static_cast<int>(N);
    g.K = static_cast<int>(K);
    g.bias = m->b[l];
    g.relu = m->relu[l];
    g.out_bf16 = last ? nullptr : m->act[cur ^ 1];
    g.out_f32 = last ? d_out : nullptr;
    const size_t tiles = ((N + kBlockN - 1) / kBlockN) * ((batch + kBlockM - 1) / kBlockM);
    const unsigned grid = static_cast<unsigned>(tiles < static_cast<size_t>(m->num_sms) ? tiles : m->num_sms);
    mlp_gemm_tcgen05_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(map_x, map_w, g);
    cur ^= 1;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "launching the MLP kernels");
  return cudaSuccess;
}

}  // namespace hpsx
