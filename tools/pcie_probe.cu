// Host-link micro-benchmark: cudaMemcpyAsync H2D/D2H from pinned memory vs a kernel gathering random
// 512-B rows straight out of mapped pinned host memory (zero-copy).  Decides how the miss path moves rows.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <sys/mman.h>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void gather_rows(const float4* __restrict__ table, const uint32_t* __restrict__ idx, size_t n, int V,
                            float4* __restrict__ out) {
  // one warp per row of V float4 (V = 32 -> 512 B)
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < n; r += nwarps) {
    const float4* src = table + (size_t)idx[r] * V;
    for (int v = lane; v < V; v += 32) out[r * V + v] = src[v];
  }
}

int main(int argc, char** argv) {
  const size_t rows = 1 << 21, V = 32;           // 2 Mi rows x 512 B = 1 GiB pinned table
  const size_t n = argc > 1 ? atol(argv[1]) : 262144;  // rows gathered per launch (128 MiB)
  float4 *h_table, *hd_table, *d_out, *h_stage;
  CK(cudaHostAlloc(&h_table, rows * V * sizeof(float4), cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&hd_table, h_table, 0));
  for (size_t i = 0; i < rows * V; i += 1024) h_table[i] = make_float4(i, 0, 0, 0);
  CK(cudaMalloc(&d_out, n * V * sizeof(float4)));
  CK(cudaMallocHost(&h_stage, n * V * sizeof(float4)));
  std::vector<uint32_t> idx(n);
  uint64_t s = 88172645463325252ull;
  for (auto& x : idx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % rows); }
  uint32_t* d_idx; CK(cudaMalloc(&d_idx, n * 4)); CK(cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = (double)n * V * 16;
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d_out, h_stage, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy H2D pinned %.0f MiB: %.3f ms  %.1f GB/s\n", bytes / 1048576, ms, bytes / ms / 1e6);
  }
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(h_stage, d_out, bytes, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy D2H pinned %.0f MiB: %.3f ms  %.1f GB/s\n", bytes / 1048576, ms, bytes / ms / 1e6);
  }
  for (int blocks : {148, 296, 592, 1184, 2368}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0)); gather_rows<<<blocks, 256>>>(hd_table, d_idx, n, V, d_out); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("zero-copy gather %zu x 512 B, %d blocks: %.3f ms  %.1f GB/s\n", n, blocks, ms, bytes / ms / 1e6);
  }
  // registered (not cudaHostAlloc'd) tables, 4 GiB, with and without transparent huge pages
  {
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    char buf[128] = {0};
    if (f) { fgets(buf, 127, f); fclose(f); }
    printf("THP: %s", buf[0] ? buf : "unknown\n");
  }
  for (int mode = 0; mode < 2; ++mode) {
    const size_t big_rows = 1 << 23;  // 8 Mi rows x 512 B = 4 GiB
    void* mem = nullptr;
    if (posix_memalign(&mem, mode ? (2u << 20) : 4096, big_rows * V * 16) != 0) { printf("alloc failed\n"); break; }
    if (mode) madvise(mem, big_rows * V * 16, MADV_HUGEPAGE);
    memset(mem, 1, big_rows * V * 16);
    CK(cudaHostRegister(mem, big_rows * V * 16, cudaHostRegisterMapped | cudaHostRegisterPortable));
    float4* dev; CK(cudaHostGetDevicePointer((void**)&dev, mem, 0));
    for (auto& x : idx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % big_rows); }
    CK(cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0)); gather_rows<<<592, 256>>>(dev, d_idx, n, V, d_out); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("zero-copy gather from cudaHostRegister'd 4 GiB (%s): %.3f ms  %.1f GB/s\n", mode ? "2M-aligned + MADV_HUGEPAGE" : "4K pages", ms, bytes / ms / 1e6);
    CK(cudaHostUnregister(mem)); free(mem);
  }
  {  // locality: sorted row order over a cudaHostAlloc'd 8 GiB table (TLB / IOTLB reach?)
    const size_t big_rows = 1 << 24;
    float4 *hb, *hbd;
    CK(cudaHostAlloc(&hb, big_rows * V * 16, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&hbd, hb, 0));
    memset(hb, 1, big_rows * V * 16);
    for (auto& x : idx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % big_rows); }
    for (int sorted = 0; sorted < 2; ++sorted) {
      if (sorted) std::sort(idx.begin(), idx.end());
      CK(cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice));
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); gather_rows<<<1184, 256>>>(hbd, d_idx, n, V, d_out); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("zero-copy gather from cudaHostAlloc'd 8 GiB, %s rows: %.3f ms  %.1f GB/s\n", sorted ? "SORTED" : "random", ms, bytes / ms / 1e6);
    }
    CK(cudaFreeHost(hb));
    FILE* f = fopen("/proc/meminfo", "r"); char line[256];
    while (f && fgets(line, 255, f)) if (strstr(line, "Huge")) printf("%s", line);
    if (f) fclose(f);
  }
  {  // cudaHostAlloc'd 4 GiB: does the allocation method or the table size set the zero-copy rate?
    const size_t big_rows = 1 << 23;
    float4 *hb, *hbd;
    CK(cudaHostAlloc(&hb, big_rows * V * 16, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&hbd, hb, 0));
    memset(hb, 1, big_rows * V * 16);
    for (auto& x : idx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % big_rows); }
    CK(cudaMemcpy(d_idx, idx.data(), n * 4, cudaMemcpyHostToDevice));
    for (int blocks : {592, 2368}) {
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); gather_rows<<<blocks, 256>>>(hbd, d_idx, n, V, d_out); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("zero-copy gather from cudaHostAlloc'd 4 GiB, %d blocks: %.3f ms  %.1f GB/s\n", blocks, ms, bytes / ms / 1e6);
    }
    CK(cudaFreeHost(hb));
  }
  // small transfers: latency of a 16 KiB and a 1 MiB H2D
  for (size_t b : {16384ul, 1048576ul, 16777216ul}) {
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d_out, h_stage, b, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy H2D %zu B: %.3f ms  %.1f GB/s\n", b, ms, b / ms / 1e6);
  }
  return 0;
}
