// NVLink peer-READ micro-benchmark (2 GPUs, one process): how fast can SMs of GPU0 pull 512-B rows OUT of GPU1's memory?
// The NVLink tier reads cache misses from the owner's HBM one-sidedly; this probe separates what bounds it:
//   region   size of the remote table the random rows are spread over (TLB / page-table reach of peer mappings)
//   order    random rows vs ascending rows
//   depth    rows in flight per warp (1, 4, 8) and CTAs per SM
//   width    16-B (LDG.128) vs 32-B (LDG.256) loads, and cp.async.bulk (TMA engine) row reads
// Compare with tools/nvlink_probe.cu (peer STORES: 709 GB/s whatever the pattern) and the copy-engine memcpy.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int V = 32;  // float4 per row (512 B)

// warp-per-row-group gather: kRows rows in flight per warp, 16 B per lane per row; rows -> sequential local output
template <int kRows>
__global__ void __launch_bounds__(256) rows_ld(const float4* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n,
                                               float4* out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * 256u + threadIdx.x) >> 5, nwarps = (gridDim.x * 256u) >> 5;
  for (uint32_t i0 = warp * kRows; i0 < n; i0 += nwarps * kRows) {
    float4 x[kRows];
    uint32_t r[kRows];
#pragma unroll
    for (int e = 0; e < kRows; ++e) r[e] = i0 + e < n ? idx[i0 + e] : 0u;
#pragma unroll
    for (int e = 0; e < kRows; ++e) x[e] = table[(size_t)r[e] * V + lane];
#pragma unroll
    for (int e = 0; e < kRows; ++e)
      if (i0 + e < n) __stcs(out + (size_t)(i0 + e) * V + lane, x[e]);
  }
}

struct alignas(32) V8 { float v[8]; };
__device__ __forceinline__ V8 ld256(const V8* p) {
  V8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
// 32-B loads: a row is 16 lanes, a warp instruction covers two rows; kPairs row pairs in flight per warp
template <int kPairs>
__global__ void __launch_bounds__(256) rows_ld256(const V8* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n, V8* out) {
  const uint32_t lane = threadIdx.x & 31u, half = lane >> 4, sub = lane & 15u;
  const uint32_t warp = (blockIdx.x * 256u + threadIdx.x) >> 5, nwarps = (gridDim.x * 256u) >> 5;
  for (uint32_t i0 = warp * 2 * kPairs; i0 < n; i0 += nwarps * 2 * kPairs) {
    V8 x[kPairs];
#pragma unroll
    for (int e = 0; e < kPairs; ++e) {
      const uint32_t i = i0 + 2 * e + half;
      const uint32_t r = i < n ? idx[i] : 0u;
      x[e] = ld256(table + (size_t)r * 16 + sub);
    }
#pragma unroll
    for (int e = 0; e < kPairs; ++e) {
      const uint32_t i = i0 + 2 * e + half;
      if (i < n) out[(size_t)i * 16 + sub] = x[e];
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// TMA-engine reads: every warp keeps kRows bulk copies of 512 B in flight (global -> shared), then stores them out
template <int kRows>
__global__ void __launch_bounds__(256) rows_bulk_ld(const float4* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n,
                                                    float4* out) {
  __shared__ __align__(128) float4 stage[8][kRows][V];
  __shared__ uint64_t bars[8];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t warp = (blockIdx.x * 256u + threadIdx.x) >> 5, nwarps = (gridDim.x * 256u) >> 5;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[w])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;
  for (uint32_t i0 = warp * kRows; i0 < n; i0 += nwarps * kRows) {
    const uint32_t cnt = min((uint32_t)kRows, n - i0);
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[w])), "r"(cnt * 512u) : "memory");
    __syncwarp();
    if (lane < cnt) {
      const uint32_t r = idx[i0 + lane];
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(&stage[w][lane][0])),
                   "l"(table + (size_t)r * V), "r"(512u), "r"(smem_u32(&bars[w]))
                   : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bars[w])),
        "r"(phase)
        : "memory");
    phase ^= 1u;
    for (uint32_t e = 0; e < cnt; ++e) __stcs(out + (size_t)(i0 + e) * V + lane, stage[w][e][lane]);
    __syncwarp();
  }
}

static float time_ms(cudaStream_t s, void (*fn)(void*), void* ctx, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  fn(ctx);
  CK(cudaStreamSynchronize(s));
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0, s));
    fn(ctx);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  return best;
}

struct Ctx {
  const float4* table;
  const uint32_t* idx;
  uint32_t n;
  float4* out;
  int grid;
  int kind;
  cudaStream_t s;
};
static void launch(void* p) {
  Ctx* c = static_cast<Ctx*>(p);
  switch (c->kind) {
    case 1: rows_ld<1><<<c->grid, 256, 0, c->s>>>(c->table, c->idx, c->n, c->out); break;
    case 4: rows_ld<4><<<c->grid, 256, 0, c->s>>>(c->table, c->idx, c->n, c->out); break;
    case 8: rows_ld<8><<<c->grid, 256, 0, c->s>>>(c->table, c->idx, c->n, c->out); break;
    case 16: rows_ld<16><<<c->grid, 256, 0, c->s>>>(c->table, c->idx, c->n, c->out); break;
    case 256: rows_ld256<4><<<c->grid, 256, 0, c->s>>>((const V8*)c->table, c->idx, c->n, (V8*)c->out); break;
    case 100: rows_bulk_ld<8><<<c->grid, 256, 0, c->s>>>(c->table, c->idx, c->n, c->out); break;
  }
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const uint32_t n = 1u << 20;  // rows per pass: 512 MiB read
  size_t max_rows = (argc > 1 ? (size_t)atoll(argv[1]) : 64ull << 30) / 512;
  CK(cudaSetDevice(1));
  float4* remote = nullptr;
  CK(cudaMalloc(&remote, max_rows * 512));
  CK(cudaMemset(remote, 1, max_rows * 512));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  float4* local = nullptr;
  CK(cudaMalloc(&local, (8ull << 30)));
  CK(cudaMemset(local, 1, (8ull << 30)));
  float4* out = nullptr;
  CK(cudaMalloc(&out, (size_t)n * 512));
  uint32_t* d_idx = nullptr;
  CK(cudaMalloc(&d_idx, n * 4));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemcpyPeerAsync(out, 0, remote, 1, (size_t)n * 512, s));
    CK(cudaEventRecord(e0, s));
    CK(cudaMemcpyPeerAsync(out, 0, remote, 1, (size_t)n * 512, s));
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("copy engine, 512 MiB remote -> local: %.3f ms  %.1f GB/s\n", ms, n * 512.0 / ms / 1e6);
  }
  std::vector<uint32_t> h(n);
  auto fill = [&](size_t region_rows, bool ascending) {
    uint64_t x = 88172645463325252ull;
    for (uint32_t i = 0; i < n; ++i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      h[i] = ascending ? (uint32_t)(((uint64_t)i * region_rows) / n) : (uint32_t)(x % region_rows);
    }
    CK(cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice));
  };
  const size_t regions[] = {256ull << 20, 2ull << 30, 8ull << 30, 64ull << 30};
  for (size_t reg : regions) {
    if (reg / 512 > max_rows || argc > 2) continue;
    for (int asc = 0; asc < 2; ++asc) {
      fill(reg / 512, asc != 0);
      printf("== remote region %5.1f GiB, %s rows\n", reg / 1073741824.0, asc ? "ascending" : "random");
      const int kinds[] = {1, 4, 8, 16, 256, 100};
      const char* names[] = {"LDG.128 1 row/warp", "LDG.128 4 rows/warp", "LDG.128 8 rows/warp", "LDG.128 16 rows/warp",
                             "LDG.256 8 rows/warp", "bulk (TMA) 8 rows/warp"};
      for (int k = 0; k < 6; ++k) {
        printf("   %-26s", names[k]);
        for (int ctas_per_sm : {1, 2, 4, 8}) {
          if (kinds[k] >= 100 && ctas_per_sm > 4) { printf("            "); continue; }
          Ctx c{remote, d_idx, n, out, 148 * ctas_per_sm, kinds[k], s};
          const float ms = time_ms(s, launch, &c);
          printf("  x%d %6.1f", ctas_per_sm, n * 512.0 / ms / 1e6);
        }
        printf("  GB/s\n");
      }
    }
  }
  // BOTH directions at once: GPU1 gathers from GPU0's memory while GPU0 gathers from GPU1's (what replicas with the
  // tier, and the model-parallel table, do all the time)
  {
    CK(cudaSetDevice(1));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    float4* out1 = nullptr;
    uint32_t* d_idx1 = nullptr;
    CK(cudaMalloc(&out1, (size_t)n * 512));
    CK(cudaMalloc(&d_idx1, n * 4));
    fill((8ull << 30) / 512, false);
    CK(cudaMemcpy(d_idx1, h.data(), n * 4, cudaMemcpyHostToDevice));
    cudaStream_t s1;
    CK(cudaStreamCreate(&s1));
    CK(cudaSetDevice(0));
    printf("== both directions at once, 8 GiB regions, random rows, LDG.128 4 rows/warp\n");
    for (int ctas_per_sm : {1, 2, 4, 8}) {
      for (int reps_1 : {0, 8}) {  // 0: GPU1 idle (one direction), 8: GPU1 runs 8 passes meanwhile
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaSetDevice(1));
        for (int r = 0; r < reps_1; ++r) rows_ld<4><<<148 * ctas_per_sm, 256, 0, s1>>>(local, d_idx1, n, out1);
        CK(cudaSetDevice(0));
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < 4; ++r) rows_ld<4><<<148 * ctas_per_sm, 256, 0, s>>>(remote, d_idx, n, out);
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaSetDevice(1)); CK(cudaStreamSynchronize(s1)); CK(cudaSetDevice(0));
        printf("   x%d  GPU0 <- GPU1 %6.1f GB/s  (%s)\n", ctas_per_sm, 4.0 * n * 512.0 / ms / 1e6,
               reps_1 ? "GPU1 <- GPU0 running at the same time" : "GPU1 idle");
      }
    }
  }
  // the same gather from LOCAL HBM for reference (8 GiB region, random)
  fill((8ull << 30) / 512, false);
  printf("== LOCAL region 8 GiB, random rows\n");
  for (int kind : {1, 4, 8}) {
    Ctx c{local, d_idx, n, out, 148 * 8, kind, s};
    printf("   LDG.128 %d rows/warp x8: %.1f GB/s\n", kind, n * 512.0 / time_ms(s, launch, &c) / 1e6);
  }
  return 0;
}
