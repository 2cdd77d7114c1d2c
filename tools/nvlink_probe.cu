// NVLink peer-store micro-benchmark (2 GPUs, one process): how fast can SMs of GPU0 push 512-B rows into GPU1's
// memory?  Decides how the model-parallel gather kernel returns rows to their requesters.
//   memcpy      cudaMemcpyPeerAsync (copy engines): the link ceiling
//   st.v4       random local 512-B rows -> scattered remote rows, 16-B st.global.cs per lane (the shipped kernel's shape)
//   st.v4 seq   same, sequential source and destination (pure streaming copy through SMs)
//   bulk        rows staged in shared memory with cp.async.bulk g2s, then ONE cp.async.bulk s2g of 512 B per row
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int V = 32;  // float4 per row

template <int kUnroll>
__global__ void __launch_bounds__(256) rows_st(const float4* __restrict__ table, const uint32_t* __restrict__ src_idx,
                                               const uint32_t* __restrict__ dst_idx, size_t n, float4* out) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t tile_base = ((blockIdx.x * (size_t)256 + threadIdx.x) >> 5) * 32;
  if (tile_base >= n) return;
  const uint32_t s = src_idx[tile_base + lane], d = dst_idx[tile_base + lane];
  for (uint32_t i0 = 0; i0 < 32 * V; i0 += 32 * kUnroll) {
    float4 buf[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32 + lane, kk = i / V;
      buf[u] = __ldg(table + (size_t)__shfl_sync(0xffffffffu, s, kk) * V + (i - kk * V));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32 + lane, kk = i / V;
      __stcs(out + (size_t)__shfl_sync(0xffffffffu, d, kk) * V + (i - kk * V), buf[u]);
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one warp = one tile of 32 rows; kStages tiles in flight per warp
template <int kWarps, int kStages>
__global__ void __launch_bounds__(kWarps * 32) rows_bulk(const float4* __restrict__ table, const uint32_t* __restrict__ src_idx,
                                                         const uint32_t* __restrict__ dst_idx, size_t n, float4* out,
                                                         uint32_t tiles_per_warp) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  constexpr uint32_t kTile = 32 * V * 16;
  unsigned char* ring = smem + (size_t)warp * kStages * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarps * kStages * kTile) + warp * kStages;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const size_t first = ((size_t)blockIdx.x * kWarps + warp) * tiles_per_warp;
  const size_t num_tiles = n / 32;
  uint32_t phase = 0;
  uint32_t pend_d = 0;
  int pend = -1;
  for (uint32_t t = 0; t <= tiles_per_warp; ++t) {
    const size_t tile = first + t;
    const bool have = t < tiles_per_warp && tile < num_tiles;
    const int stage = t % kStages;
    if (have) {
      const uint32_t s = src_idx[tile * 32 + lane];
      // bulk groups are per thread: every lane waits for its own store that last read this stage
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStages - 2) : "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[stage])), "r"(kTile) : "memory");
      }
      __syncwarp();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(ring + (size_t)stage * kTile + lane * V * 16)),
                   "l"(table + (size_t)s * V), "r"(V * 16), "r"(smem_u32(&bars[stage]))
                   : "memory");
    }
    if (pend >= 0) {
      const uint32_t par = (phase >> pend) & 1u;
      asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(
                       smem_u32(&bars[pend])),
                   "r"(par)
                   : "memory");
      phase ^= 1u << pend;
      // every lane stores its own row: 512-B bulk store to the (peer) destination
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)pend_d * V),
                   "r"(smem_u32(ring + (size_t)pend * kTile + lane * V * 16)), "r"(V * 16)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      pend = -1;
    }
    if (have) {
      pend = stage;
      pend_d = dst_idx[tile * 32 + lane];
    }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const size_t rows = 4u << 20;                          // 2 GiB source table on GPU0
  const size_t n = argc > 1 ? atol(argv[1]) : (1u << 20);  // rows moved per launch (512 MiB)
  const int dst_dev = ndev > 1 ? 1 : 0;
  if (ndev < 2) printf("only one GPU visible: destination is local HBM (no NVLink)\n");
  CK(cudaSetDevice(0));
  if (dst_dev != 0) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, dst_dev));
    if (!can) { printf("no peer access\n"); return 1; }
    CK(cudaDeviceEnablePeerAccess(dst_dev, 0));
  }
  float4 *table, *out, *local_out;
  CK(cudaMalloc(&table, rows * V * 16));
  CK(cudaMemset(table, 1, rows * V * 16));
  CK(cudaMalloc(&local_out, n * V * 16));
  CK(cudaSetDevice(dst_dev));
  CK(cudaMalloc(&out, n * V * 16));
  CK(cudaSetDevice(0));
  std::vector<uint32_t> si(n), di(n), seq(n);
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; si[i] = (uint32_t)(s % rows); seq[i] = (uint32_t)i; di[i] = (uint32_t)i; }
  for (size_t i = n - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(di[i], di[s % (i + 1)]); }
  uint32_t *d_si, *d_di, *d_seq;
  CK(cudaMalloc(&d_si, n * 4)); CK(cudaMalloc(&d_di, n * 4)); CK(cudaMalloc(&d_seq, n * 4));
  CK(cudaMemcpy(d_si, si.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_di, di.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_seq, seq.data(), n * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = (double)n * V * 16;
  auto report = [&](const char* name, auto&& fn) {
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
      CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-44s %8.3f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
  };
  const unsigned grid = (unsigned)(n * 32 / 256 / 32 * 1);  // one warp per 32-row tile
  report("memcpy peer (copy engine)", [&] { CK(cudaMemcpyPeerAsync(out, dst_dev, table, 0, (size_t)bytes, 0)); });
  report("st.v4 unroll 8: random rows -> scattered rows", [&] { rows_st<8><<<grid, 256>>>(table, d_si, d_di, n, out); });
  report("st.v4 unroll 8: random rows -> sequential rows", [&] { rows_st<8><<<grid, 256>>>(table, d_si, d_seq, n, out); });
  report("st.v4 unroll 8: sequential -> sequential", [&] { rows_st<8><<<grid, 256>>>(table, d_seq, d_seq, n, out); });
  report("st.v4 unroll 4: random rows -> scattered rows", [&] { rows_st<4><<<grid, 256>>>(table, d_si, d_di, n, out); });
  report("st.v4 unroll 8: random -> scattered LOCAL HBM", [&] { rows_st<8><<<grid, 256>>>(table, d_si, d_di, n, local_out); });
  {
    constexpr int W = 4, S = 3;
    const size_t smem = (size_t)W * S * 32 * V * 16 + W * S * 8;
    CK(cudaFuncSetAttribute(rows_bulk<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (uint32_t tpw : {4u, 16u}) {
      const size_t warps = (n / 32 + tpw - 1) / tpw;
      const unsigned g = (unsigned)((warps + W - 1) / W);
      char name[96];
      snprintf(name, sizeof name, "bulk g2s+s2g 4 warps x 3 stages, %u tiles/warp", tpw);
      report(name, [&] { rows_bulk<W, S><<<g, W * 32, smem>>>(table, d_si, d_di, n, out, tpw); });
      snprintf(name, sizeof name, "  same -> LOCAL HBM");
      report(name, [&] { rows_bulk<W, S><<<g, W * 32, smem>>>(table, d_si, d_di, n, local_out, tpw); });
    }
  }
  {
    constexpr int W = 8, S = 2;
    const size_t smem = (size_t)W * S * 32 * V * 16 + W * S * 8;
    CK(cudaFuncSetAttribute(rows_bulk<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t tpw = 8;
    const size_t warps = (n / 32 + tpw - 1) / tpw;
    const unsigned g = (unsigned)((warps + W - 1) / W);
    report("bulk g2s+s2g 8 warps x 2 stages, 8 tiles/warp", [&] { rows_bulk<W, S><<<g, W * 32, smem>>>(table, d_si, d_di, n, out, tpw); });
  }
  return 0;
}
