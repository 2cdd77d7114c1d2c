// Host-link micro-benchmark, round 2 (b): can a PCIe-bound gather run BESIDE an HBM-bound kernel without either
// losing its rate?  The chunked binned pull of round 2 showed both kernels collapsing when they share SMs (probe
// kernel 5x slower, pull 2.5x slower: profiles/r02_pipeline_timeline.txt).  This probe measures
//   (1) how many SMs a zero-copy gather of 512-B rows needs for the full link rate (k CTAs of 1024 threads,
//       1/2/4 rows in flight per warp), and
//   (2) gather and an HBM copy kernel side by side, with the gather CTAs either sharing SMs with the copy CTAs
//       or owning their SMs (each gather CTA asks for 227 KB of shared memory, so no other CTA fits beside it).
// usage: pcie_probe3 [table_GiB=5] [rows_per_launch=172800]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x)                                                     \
  do {                                                            \
    cudaError_t e = (x);                                          \
    if (e != cudaSuccess) {                                       \
      printf("%s: %s\n", #x, cudaGetErrorString(e));              \
      exit(1);                                                    \
    }                                                             \
  } while (0)

constexpr int kRowBytes = 512;

template <int kRows>
__global__ void __launch_bounds__(1024) gather_k(const unsigned long long* __restrict__ addr, size_t n, float4* __restrict__ out) {
  extern __shared__ unsigned char smem[];  // only reserved (exclusive-SM mode), never touched
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r0 = warp * kRows; r0 < n; r0 += nwarps * kRows) {
    float4 x[kRows];
#pragma unroll
    for (int j = 0; j < kRows; ++j)
      if (r0 + j < n) x[j] = reinterpret_cast<const float4*>(addr[r0 + j])[lane];
#pragma unroll
    for (int j = 0; j < kRows; ++j)
      if (r0 + j < n) __stcs(&out[(r0 + j) * 32 + lane], x[j]);
  }
}

__global__ void hbm_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int reps) {
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// HBM-side co-runner shaped like the probe+gather kernel: random 512-B rows of a 2 GiB table -> contiguous output
__global__ void hbm_random_gather(const float4* __restrict__ table, const uint32_t* __restrict__ idx, size_t n, float4* __restrict__ out,
                                  int reps) {
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (int rep = 0; rep < reps; ++rep)
    for (size_t r = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps)
      __stcs(&out[r * 32 + lane], table[(size_t)idx[(r * 7 + rep * 13) % n] * 32 + lane]);
}

int main(int argc, char** argv) {
  const size_t gib = argc > 1 ? atol(argv[1]) : 5;
  const size_t n = argc > 2 ? atol(argv[2]) : 172800;
  const double bytes = (double)n * kRowBytes;
  cudaEvent_t e0, e1, c0, c1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&c0));
  CK(cudaEventCreate(&c1));
  // table: registered 256-MiB slabs, rows visited in ascending order (what the binned pull produces)
  const size_t slab_bytes = 256ull << 20, slabs = (gib << 30) / slab_bytes, rows_per_slab = slab_bytes / kRowBytes;
  std::vector<char*> dev(slabs);
  for (size_t i = 0; i < slabs; ++i) {
    void* mem = nullptr;
    if (posix_memalign(&mem, 4096, slab_bytes) != 0) return 1;
    memset(mem, 1, slab_bytes);
    CK(cudaHostRegister(mem, slab_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    void* d = nullptr;
    CK(cudaHostGetDevicePointer(&d, mem, 0));
    dev[i] = static_cast<char*>(d);
  }
  std::vector<size_t> rows(n);
  uint64_t s = 88172645463325252ull;
  for (auto& r : rows) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    r = s % (slabs * rows_per_slab);
  }
  std::sort(rows.begin(), rows.end());
  std::vector<unsigned long long> addr(n);
  for (size_t i = 0; i < n; ++i)
    addr[i] = reinterpret_cast<unsigned long long>(dev[rows[i] / rows_per_slab]) + (rows[i] % rows_per_slab) * kRowBytes;
  unsigned long long* d_addr;
  float4* d_out;
  CK(cudaMalloc(&d_addr, n * 8));
  CK(cudaMalloc(&d_out, n * kRowBytes));
  CK(cudaMemcpy(d_addr, addr.data(), n * 8, cudaMemcpyHostToDevice));
  float4 *a, *b;
  const size_t cn = (1ull << 30) / 16;
  CK(cudaMalloc(&a, cn * 16));
  CK(cudaMalloc(&b, cn * 16));
  CK(cudaMemset(a, 1, cn * 16));
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  const int kExcl = 227 * 1024;
  CK(cudaFuncSetAttribute(gather_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExcl));
  CK(cudaFuncSetAttribute(gather_k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExcl));
  CK(cudaFuncSetAttribute(gather_k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExcl));

  auto launch_gather = [&](int k, int rows_in_flight, size_t smem, cudaStream_t st) {
    if (rows_in_flight == 1) gather_k<1><<<k, 1024, smem, st>>>(d_addr, n, d_out);
    else if (rows_in_flight == 2) gather_k<2><<<k, 1024, smem, st>>>(d_addr, n, d_out);
    else gather_k<4><<<k, 1024, smem, st>>>(d_addr, n, d_out);
  };
  float ms = 0, cms = 0;
  printf("== (1) gather alone: k CTAs x 1024 threads, R rows in flight per warp; %zu rows x 512 B from a %zu GiB table, ascending\n", n, gib);
  for (int k : {4, 8, 12, 16, 24, 32, 48, 74, 148}) {
    printf("   k=%3d:", k);
    for (int r : {1, 2, 4}) {
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0, s1));
        launch_gather(k, r, 0, s1);
        CK(cudaEventRecord(e1, s1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("  R=%d %5.1f GB/s", r, bytes / ms / 1e6);
    }
    printf("\n");
  }
  // HBM copy alone: 8 x (1 GiB read + 1 GiB write)
  const int reps = 8;
  const double copy_bytes = 2.0 * cn * 16 * reps;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(c0, s2));
    hbm_copy<<<148 * 8, 256, 0, s2>>>(a, b, cn, reps);
    CK(cudaEventRecord(c1, s2));
    CK(cudaEventSynchronize(c1));
    CK(cudaEventElapsedTime(&cms, c0, c1));
  }
  printf("== HBM copy alone: %.2f ms, %.0f GB/s\n", cms, copy_bytes / cms / 1e6);
  printf("== (2) side by side: the gather is launched first (idle GPU), the copy right behind it on another stream\n");
  for (int excl = 0; excl < 2; ++excl) {
    for (int k : {8, 16, 24, 32, 48}) {
      for (int r : {2, 4}) {
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaDeviceSynchronize());
          CK(cudaEventRecord(e0, s1));
          launch_gather(k, r, excl ? kExcl : 0, s1);
          CK(cudaEventRecord(e1, s1));
          CK(cudaEventRecord(c0, s2));
          hbm_copy<<<148 * 8, 256, 0, s2>>>(a, b, cn, reps);
          CK(cudaEventRecord(c1, s2));
          CK(cudaDeviceSynchronize());
          CK(cudaGetLastError());
          CK(cudaEventElapsedTime(&ms, e0, e1));
          CK(cudaEventElapsedTime(&cms, c0, c1));
        }
        printf("   %s SMs, k=%2d R=%d: gather %5.1f GB/s (%.2f ms) | copy %5.0f GB/s (%.2f ms)\n", excl ? "OWN   " : "shared", k, r,
               bytes / ms / 1e6, ms, copy_bytes / cms / 1e6, cms);
      }
    }
  }
  // (3) the same with a RANDOM-gather co-runner (the probe kernel's access pattern): 2 GiB table, 1.7 M rows per pass
  {
    const size_t trows = (2ull << 30) / kRowBytes, gn = 1703936;
    float4 *table, *gout;
    uint32_t* gidx;
    CK(cudaMalloc(&table, trows * kRowBytes));
    CK(cudaMalloc(&gout, gn * kRowBytes));
    CK(cudaMalloc(&gidx, gn * 4));
    std::vector<uint32_t> hi(gn);
    for (auto& x : hi) {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (uint32_t)(s % trows);
    }
    CK(cudaMemcpy(gidx, hi.data(), gn * 4, cudaMemcpyHostToDevice));
    const int greps = 6;
    const double gbytes = 2.0 * gn * kRowBytes * greps;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(c0, s2));
      hbm_random_gather<<<148 * 8, 256, 0, s2>>>(table, gidx, gn, gout, greps);
      CK(cudaEventRecord(c1, s2));
      CK(cudaEventSynchronize(c1));
      CK(cudaEventElapsedTime(&cms, c0, c1));
    }
    printf("== HBM random gather alone: %.2f ms, %.0f GB/s\n", cms, gbytes / cms / 1e6);
    printf("== (3) PCIe gather beside the HBM random gather\n");
    for (int excl = 0; excl < 2; ++excl) {
      for (int k : {8, 16, 32, 148}) {
        if (excl && k == 148) continue;
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaDeviceSynchronize());
          CK(cudaEventRecord(e0, s1));
          launch_gather(k, 2, excl ? kExcl : 0, s1);
          CK(cudaEventRecord(e1, s1));
          CK(cudaEventRecord(c0, s2));
          hbm_random_gather<<<148 * 8, 256, 0, s2>>>(table, gidx, gn, gout, greps);
          CK(cudaEventRecord(c1, s2));
          CK(cudaDeviceSynchronize());
          CK(cudaGetLastError());
          CK(cudaEventElapsedTime(&ms, e0, e1));
          CK(cudaEventElapsedTime(&cms, c0, c1));
        }
        printf("   %s SMs, k=%3d R=2: PCIe gather %5.1f GB/s (%.2f ms) | HBM random gather %5.0f GB/s (%.2f ms)\n", excl ? "OWN   " : "shared", k,
               bytes / ms / 1e6, ms, gbytes / cms / 1e6, cms);
      }
    }
  }
  return 0;
}
