// Host-link micro-benchmark, round 2: what sets the zero-copy gather rate of 512-B rows out of a multi-GB
// page-locked host table?  Decides the shape of the miss path (DESIGN.md §3 "Miss path"):
//   order    random | sorted | coarse bins (rows random INSIDE a window of 64 MiB .. 1 GiB, windows ascending)
//   loads    ld.global 16 B/lane | ld.global.nc 32 B/lane | cp.async.bulk 512 B (TMA engine) -> smem -> bulk store
//   grid     1..8 CTAs of 256 threads per SM
//   memory   cudaHostRegister'd 256-MiB slabs (production layout) | + MADV_HUGEPAGE | one cudaHostAlloc
//   overlap  the same gather while an HBM-bound copy kernel runs on a second stream (does sharing SMs cost link rate?)
// usage: pcie_probe2 [table_GiB=5] [rows_per_launch=172800]
#include <cuda_runtime.h>
#include <sys/mman.h>

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

__global__ void gather_ld16(const unsigned long long* __restrict__ addr, size_t n, float4* __restrict__ out) {
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < n; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(addr[r]);
    out[r * 32 + lane] = src[lane];
  }
}

struct alignas(32) V8 {
  float v[8];
};
__global__ void gather_ld32(const unsigned long long* __restrict__ addr, size_t n, V8* __restrict__ out) {
  // 16 lanes per row, two rows per warp instruction
  const size_t half = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
  const int l = threadIdx.x & 15;
  const size_t nhalf = ((size_t)gridDim.x * blockDim.x) >> 4;
  for (size_t r = half; r < n; r += nhalf) {
    const V8* src = reinterpret_cast<const V8*>(addr[r]) + l;
    V8 x;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(x.v[0]), "=f"(x.v[1]), "=f"(x.v[2]), "=f"(x.v[3]), "=f"(x.v[4]), "=f"(x.v[5]), "=f"(x.v[6]),
                   "=f"(x.v[7])
                 : "l"(src));
    out[r * 16 + l] = x;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// one warp owns a ring of kSlots 512-B rows in shared memory; lane 0 issues bulk loads (host -> smem) and bulk
// stores (smem -> HBM)
template <int kSlots>
__global__ void gather_bulk(const unsigned long long* __restrict__ addr, size_t n, unsigned char* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31, warps_per_cta = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp_in_cta * kSlots * kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps_per_cta * kSlots * kRowBytes) + warp_in_cta * kSlots;
  if (lane == 0) {
    for (int s = 0; s < kSlots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane != 0) return;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  // rows warp, warp + nwarps, ...; software pipeline of depth kSlots
  size_t issue = warp, retire = warp;
  uint32_t phase = 0;
  int is = 0, rs = 0, inflight = 0;
  while (retire < n) {
    while (inflight < kSlots && issue < n) {
      // the bulk store that last read this slot must be done reading shared memory
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSlots - 1) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[is])), "r"(kRowBytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(ring + is * kRowBytes)),
                   "l"(addr[issue]), "r"(kRowBytes), "r"(smem_u32(&bars[is]))
                   : "memory");
      issue += nwarps;
      is = (is + 1) % kSlots;
      ++inflight;
    }
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(
            smem_u32(&bars[rs])),
        "r"((phase >> rs) & 1u)
        : "memory");
    phase ^= 1u << rs;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + retire * kRowBytes),
                 "r"(smem_u32(ring + rs * kRowBytes)), "r"(kRowBytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    retire += nwarps;
    rs = (rs + 1) % kSlots;
    --inflight;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__global__ void hbm_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n, int reps) {
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

struct Table {
  std::vector<char*> host, dev;
  size_t slab_bytes = 0, rows_per_slab = 0, rows = 0;
  unsigned long long addr(size_t row) const {
    return reinterpret_cast<unsigned long long>(dev[row / rows_per_slab]) + (row % rows_per_slab) * kRowBytes;
  }
};

int main(int argc, char** argv) {
  const size_t gib = argc > 1 ? atol(argv[1]) : 5;
  const size_t n = argc > 2 ? atol(argv[2]) : 172800;
  const double bytes = (double)n * kRowBytes;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  unsigned long long* d_addr;
  unsigned char* d_out;
  CK(cudaMalloc(&d_addr, n * 8));
  CK(cudaMalloc(&d_out, n * kRowBytes));
  std::vector<unsigned long long> addr(n);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
  };
  {
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    char buf[128] = {0};
    if (f) {
      if (!fgets(buf, 127, f)) buf[0] = 0;
      fclose(f);
    }
    printf("THP: %s", buf[0] ? buf : "unknown\n");
  }
  float ms = 0;
  auto time_it = [&](auto&& launch) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    return bytes / ms / 1e6;
  };
  CK(cudaFuncSetAttribute(gather_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(gather_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));

  const int mem_modes = argc > 3 ? atoi(argv[3]) : 3;
  for (int mem_mode = 0; mem_mode < mem_modes; ++mem_mode) {
    const char* mem_name[] = {"registered 256-MiB slabs (4K-aligned malloc)", "registered 256-MiB slabs, 2M-aligned + MADV_HUGEPAGE",
                              "one cudaHostAlloc"};
    Table t;
    t.slab_bytes = mem_mode == 2 ? gib << 30 : 256ull << 20;
    t.rows_per_slab = t.slab_bytes / kRowBytes;
    const size_t slabs = (gib << 30) / t.slab_bytes;
    t.rows = slabs * t.rows_per_slab;
    for (size_t i = 0; i < slabs; ++i) {
      void* mem = nullptr;
      void* dev = nullptr;
      if (mem_mode == 2) {
        CK(cudaHostAlloc(&mem, t.slab_bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(mem, 1, t.slab_bytes);
      } else {
        if (posix_memalign(&mem, mem_mode == 1 ? (2u << 20) : 4096, t.slab_bytes) != 0) {
          printf("alloc failed\n");
          return 1;
        }
        if (mem_mode == 1) madvise(mem, t.slab_bytes, MADV_HUGEPAGE);
        memset(mem, 1, t.slab_bytes);
        CK(cudaHostRegister(mem, t.slab_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
      }
      CK(cudaHostGetDevicePointer(&dev, mem, 0));
      t.host.push_back(static_cast<char*>(mem));
      t.dev.push_back(static_cast<char*>(dev));
    }
    printf("== %zu GiB table, %s; %zu rows x 512 B per launch (%.1f MB)\n", gib, mem_name[mem_mode], n, bytes / 1e6);
    if (mem_mode == 1) {
      FILE* f = fopen("/proc/meminfo", "r");
      char line[256];
      while (f && fgets(line, 255, f))
        if (strstr(line, "AnonHugePages")) printf("   %s", line);
      if (f) fclose(f);
    }
    std::vector<size_t> rows(n);
    for (auto& r : rows) r = rnd() % t.rows;
    // orders: -1 random, 0 fully sorted, else bin shift (rows random inside a window of 2^shift bytes)
    for (int shift : {-1, 0, 23, 24, 25, 26, 28, 30}) {
      if (mem_mode != 0 && shift > 0) continue;
      std::vector<size_t> ord = rows;
      if (shift == 0) {
        std::sort(ord.begin(), ord.end());
      } else if (shift > 0) {
        const size_t rows_per_bin = (1ull << shift) / kRowBytes;
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return a / rows_per_bin < b / rows_per_bin; });
      }
      for (size_t i = 0; i < n; ++i) addr[i] = t.addr(ord[i]);
      CK(cudaMemcpy(d_addr, addr.data(), n * 8, cudaMemcpyHostToDevice));
      char oname[64];
      if (shift < 0) snprintf(oname, 64, "random");
      else if (shift == 0) snprintf(oname, 64, "sorted");
      else snprintf(oname, 64, "bins of %d MiB", 1 << (shift - 20));
      for (int cps : {1, 2, 4, 8}) {
        if ((shift > 0 || mem_mode != 0) && cps != 1 && cps != 2 && cps != 8) continue;
        const double g16 = time_it([&] { gather_ld16<<<148 * cps, 256>>>(d_addr, n, reinterpret_cast<float4*>(d_out)); });
        const double g32 = time_it([&] { gather_ld32<<<148 * cps, 256>>>(d_addr, n, reinterpret_cast<V8*>(d_out)); });
        printf("   %-16s %d CTA/SM: ld16 %5.1f GB/s | ld32 %5.1f GB/s", oname, cps, g16, g32);
        if (cps <= 2) {
          const size_t sm4 = 8 * 4 * kRowBytes + 8 * 4 * 8, sm8 = 8 * 8 * kRowBytes + 8 * 8 * 8;
          const double b4 = time_it([&] { gather_bulk<4><<<148 * cps, 256, sm4>>>(d_addr, n, d_out); });
          const double b8 = time_it([&] { gather_bulk<8><<<148 * cps, 256, sm8>>>(d_addr, n, d_out); });
          printf(" | bulk x4 %5.1f | bulk x8 %5.1f", b4, b8);
        }
        printf("\n");
      }
    }
    if (mem_mode == 0) {
      // sorted gather on 1 CTA/SM while an HBM copy saturates the other SM resources
      std::vector<size_t> ord = rows;
      std::sort(ord.begin(), ord.end());
      for (size_t i = 0; i < n; ++i) addr[i] = t.addr(ord[i]);
      CK(cudaMemcpy(d_addr, addr.data(), n * 8, cudaMemcpyHostToDevice));
      float4 *a, *b;
      const size_t cn = (1ull << 30) / 16;
      CK(cudaMalloc(&a, cn * 16));
      CK(cudaMalloc(&b, cn * 16));
      cudaStream_t s2;
      CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      cudaStream_t s1;
      CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
      for (int cps : {1, 2}) {
        for (int rep = 0; rep < 2; ++rep) {
          hbm_copy<<<148 * 4, 256, 0, s2>>>(a, b, cn, 8);  // ~8 x 2 GiB of traffic = ~2.7 ms at 6.5 TB/s
          CK(cudaEventRecord(e0, s1));
          gather_ld16<<<148 * cps, 256, 0, s1>>>(d_addr, n, reinterpret_cast<float4*>(d_out));
          CK(cudaEventRecord(e1, s1));
          CK(cudaDeviceSynchronize());
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("   sorted, %d CTA/SM, beside an HBM copy kernel on another stream: ld16 %5.1f GB/s\n", cps, bytes / ms / 1e6);
      }
      CK(cudaFree(a));
      CK(cudaFree(b));
    }
    for (size_t i = 0; i < slabs; ++i) {
      if (mem_mode == 2) {
        CK(cudaFreeHost(t.host[i]));
      } else {
        CK(cudaHostUnregister(t.host[i]));
        free(t.host[i]);
      }
    }
  }
  return 0;
}
