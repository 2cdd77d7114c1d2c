// Host-link probe for N GPUs at once: every process (one per GPU, started together by scripts/pcie_conc.sh) page-locks
// its own 4 GiB table and, from a common wall-clock start time, runs for ~3 s each: (a) pinned cudaMemcpyAsync H2D of
// 128 MiB blocks, (b) the zero-copy gather of 512-B rows in ascending address order.  Comparing the per-GPU rates at
// N = 1, 4, 8 tells whether the 1->8 scaling loss of the miss path is the host's (DRAM / root complex) or ours.
// usage: pcie_probe_conc <device> <start_unix_seconds> [shared_table=0]
#include <cuda_runtime.h>
#include <sys/time.h>
#include <unistd.h>

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

__global__ void gather_ld16(const unsigned long long* __restrict__ addr, size_t n, float4* __restrict__ out) {
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < n; r += nwarps) out[r * 32 + lane] = reinterpret_cast<const float4*>(addr[r])[lane];
}

static double now_s() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

int main(int argc, char** argv) {
  const int dev = argc > 1 ? atoi(argv[1]) : 0;
  const double start = argc > 2 ? atof(argv[2]) : now_s() + 1;
  CK(cudaSetDevice(dev));
  const size_t n = 172800, slab_bytes = 256ull << 20, slabs = 16, rows_per_slab = slab_bytes / 512;
  std::vector<char*> d(slabs);
  for (size_t i = 0; i < slabs; ++i) {
    void* mem = nullptr;
    if (posix_memalign(&mem, 4096, slab_bytes) != 0) return 1;
    memset(mem, 1, slab_bytes);
    CK(cudaHostRegister(mem, slab_bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    void* p = nullptr;
    CK(cudaHostGetDevicePointer(&p, mem, 0));
    d[i] = static_cast<char*>(p);
  }
  std::vector<size_t> rows(n);
  uint64_t s = 88172645463325252ull + dev;
  for (auto& r : rows) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    r = s % (slabs * rows_per_slab);
  }
  std::sort(rows.begin(), rows.end());
  std::vector<unsigned long long> addr(n);
  for (size_t i = 0; i < n; ++i) addr[i] = reinterpret_cast<unsigned long long>(d[rows[i] / rows_per_slab]) + (rows[i] % rows_per_slab) * 512;
  unsigned long long* d_addr;
  float4* d_out;
  void *h_stage, *d_stage;
  CK(cudaMalloc(&d_addr, n * 8));
  CK(cudaMalloc(&d_out, n * 512));
  CK(cudaMemcpy(d_addr, addr.data(), n * 8, cudaMemcpyHostToDevice));
  CK(cudaMallocHost(&h_stage, 128 << 20));
  CK(cudaMalloc(&d_stage, 128 << 20));
  gather_ld16<<<148, 256>>>(d_addr, n, d_out);
  CK(cudaDeviceSynchronize());
  while (now_s() < start) usleep(200);
  double t0 = now_s();
  int it = 0;
  while (now_s() - t0 < 3.0) {
    CK(cudaMemcpyAsync(d_stage, h_stage, 128 << 20, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    ++it;
  }
  const double memcpy_gbs = it * (double)(128 << 20) / (now_s() - t0) / 1e9;
  while (now_s() < start + 4.0) usleep(200);
  t0 = now_s();
  it = 0;
  while (now_s() - t0 < 3.0) {
    gather_ld16<<<148, 256>>>(d_addr, n, d_out);
    CK(cudaDeviceSynchronize());
    ++it;
  }
  const double gather_gbs = it * (double)n * 512 / (now_s() - t0) / 1e9;
  printf("device %d: pinned memcpy H2D %.1f GB/s | zero-copy sorted gather %.1f GB/s\n", dev, memcpy_gbs, gather_gbs);
  return 0;
}
