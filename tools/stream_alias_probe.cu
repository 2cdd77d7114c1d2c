// Can two kernels of ONE process on ONE device always run side by side when they sit on different streams?
// The world-2-on-one-device arrangement of tests/test_shard_group_gpu.py needs it: rank A's flag-wait kernel spins
// until rank B's kernel (another stream) has published.  Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS
// hardware work queues (default 8); two streams that share a queue are serialised, and then the waiter never sees
// its flag.  This probe creates `nstreams` non-blocking streams and, for every ordered pair (i, j), launches a waiter
// on i and then a setter on j; it prints the pairs that timed out.
// usage: stream_alias_probe [nstreams=24]      (run with CUDA_DEVICE_MAX_CONNECTIONS unset, then =32)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x)                                                     \
  do {                                                            \
    cudaError_t e = (x);                                          \
    if (e != cudaSuccess) {                                       \
      printf("%s: %s\n", #x, cudaGetErrorString(e));              \
      exit(1);                                                    \
    }                                                             \
  } while (0)

__global__ void waiter(volatile unsigned* flag, unsigned want, unsigned* timed_out, unsigned long long timeout_ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  while (*flag != want) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      *timed_out = 1;
      return;
    }
    __nanosleep(64);
  }
}
__global__ void setter(volatile unsigned* flag, unsigned v) { *flag = v; }

int main(int argc, char** argv) {
  const int ns = argc > 1 ? atoi(argv[1]) : 24;
  const char* env = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
  printf("CUDA_DEVICE_MAX_CONNECTIONS=%s, %d streams\n", env ? env : "(unset: 8)", ns);
  std::vector<cudaStream_t> st(ns);
  for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  unsigned *flag, *to;
  CK(cudaMalloc(&flag, 4));
  CK(cudaMalloc(&to, 4));
  CK(cudaMemset(flag, 0, 4));
  unsigned seq = 0;
  int bad = 0, total = 0;
  for (int i = 0; i < ns; ++i) {
    for (int j = 0; j < ns; ++j) {
      if (i == j) continue;
      ++seq;
      CK(cudaMemset(to, 0, 4));
      waiter<<<1, 32, 0, st[i]>>>(flag, seq, to, 20ull * 1000 * 1000);  // 20 ms
      setter<<<1, 1, 0, st[j]>>>(flag, seq);
      CK(cudaDeviceSynchronize());
      unsigned h = 0;
      CK(cudaMemcpy(&h, to, 4, cudaMemcpyDeviceToHost));
      ++total;
      if (h) {
        ++bad;
        if (bad <= 40) printf("  waiter on stream %d never saw the setter on stream %d (serialised)\n", i, j);
      }
    }
  }
  printf("%d of %d ordered stream pairs were serialised\n", bad, total);
  return 0;
}
