// Peer reads through a CUDA IPC mapping (two PROCESSES, two GPUs): is there a per-launch cost, or a per-byte one, that the
// same reads through cudaDeviceEnablePeerAccess inside one process (tools/nvlink_read_probe.cu) do not pay?
// The child owns the table on GPU1 and exports it; the parent maps it on GPU0 and gathers n random 512-B rows for
// n = 1 Ki .. 1 Mi, then repeats with the child's GPU busy (the child gathers from its own table meanwhile).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <sys/wait.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int V = 32;
template <int kRows>
__global__ void __launch_bounds__(256) rows_ld(const float4* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n, float4* out) {
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
int main() {
  const size_t rows = (4ull << 30) / 512;
  int to_parent[2], to_child[2];
  if (pipe(to_parent) || pipe(to_child)) return 1;
  const pid_t pid = fork();
  if (pid == 0) {  // child: GPU1 owns the table
    CK(cudaSetDevice(1));
    float4* t = nullptr;
    CK(cudaMalloc(&t, rows * 512));
    CK(cudaMemset(t, 1, rows * 512));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, t));
    if (write(to_parent[1], &h, sizeof(h)) != (ssize_t)sizeof(h)) return 1;
    char cmd = 0;
    const uint32_t n = 1u << 20;
    uint32_t* idx = nullptr; float4* out = nullptr;
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&out, (size_t)n * 512));
    CK(cudaMemset(idx, 0, n * 4));
    while (read(to_child[0], &cmd, 1) == 1 && cmd != 'q') {
      if (cmd == 'o') {  // map the parent's table too: IPC peer mappings in BOTH directions, like two replicas with the tier
        cudaIpcMemHandle_t ph;
        if (read(to_child[0], &ph, sizeof(ph)) != (ssize_t)sizeof(ph)) return 1;
        void* pp = nullptr;
        CK(cudaIpcOpenMemHandle(&pp, ph, cudaIpcMemLazyEnablePeerAccess));
        rows_ld<4><<<148, 256>>>(static_cast<const float4*>(pp), idx, 1024, out);  // touch it once
        CK(cudaDeviceSynchronize());
      }
      if (cmd == 'b') {  // keep this GPU busy for a while
        for (int r = 0; r < 200; ++r) rows_ld<4><<<148 * 4, 256>>>(t, idx, n, out);
        if (write(to_parent[1], &cmd, 1) != 1) return 1;
        CK(cudaDeviceSynchronize());
      }
      if (write(to_parent[1], &cmd, 1) != 1) return 1;
    }
    return 0;
  }
  CK(cudaSetDevice(0));
  cudaIpcMemHandle_t h;
  if (read(to_parent[0], &h, sizeof(h)) != (ssize_t)sizeof(h)) return 1;
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  const float4* remote = static_cast<const float4*>(p);
  const uint32_t nmax = 1u << 20;
  std::vector<uint32_t> hidx(nmax);
  uint64_t x = 88172645463325252ull;
  for (auto& v : hidx) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (uint32_t)(x % rows); }
  uint32_t* idx = nullptr; float4* out = nullptr;
  CK(cudaMalloc(&idx, nmax * 4)); CK(cudaMalloc(&out, (size_t)nmax * 512));
  CK(cudaMemcpy(idx, hidx.data(), nmax * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto sweep = [&](const char* what) {
    printf("== %s\n", what);
    for (uint32_t n : {1u << 10, 1u << 13, 1u << 16, 173000u, 1u << 20}) {
      float best = 1e30f, first = 0;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        rows_ld<4><<<148 * 4, 256>>>(remote, idx, n, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 0) first = ms;
        best = ms < best ? ms : best;
        usleep(300);  // the link idles between launches, like between the pulls of two requests
      }
      printf("   n = %7u rows: first %.3f ms, best %.3f ms  (%.1f GB/s)\n", n, first, best, n * 512.0 / best / 1e6);
    }
  };
  sweep("IPC-mapped remote table, owner GPU idle");
  char c = 'b';
  if (write(to_child[1], &c, 1) != 1 || read(to_parent[0], &c, 1) != 1) return 1;
  sweep("IPC-mapped remote table, owner GPU busy with its own gathers");
  if (read(to_parent[0], &c, 1) != 1) return 1;
  {
    float4* mine = nullptr;
    CK(cudaMalloc(&mine, 1ull << 30));
    CK(cudaMemset(mine, 2, 1ull << 30));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, mine));
    c = 'o';
    if (write(to_child[1], &c, 1) != 1 || write(to_child[1], &mh, sizeof(mh)) != (ssize_t)sizeof(mh)) return 1;
    if (read(to_parent[0], &c, 1) != 1) return 1;
    sweep("IPC-mapped remote table, after the owner process mapped one of OUR allocations too (both directions)");
  }
  c = 'q';
  if (write(to_child[1], &c, 1) != 1) return 1;
  CK(cudaIpcCloseMemHandle(p));
  int st; waitpid(pid, &st, 0);
  return 0;
}
