// Shared host/device definitions of the HPS engine: hashing, HBM bucket layout, synthetic rows.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HPSX_HD __host__ __device__ __forceinline__
#else
#define HPSX_HD inline
#endif

namespace hpsx {

// Slot key meaning "empty".  INT64_MIN can therefore never be cached; such a key is always
// answered by the host parameter server (the reference gpu_cache reserves the max key instead).
constexpr int64_t kEmptyKey = INT64_MIN;

constexpr int kWays = 8;  // keys per bucket: 8 x int64 = 64 B = two 32-B DRAM sectors per probe

// One set of the set-associative HBM cache.  128-B aligned so that the probe (which reads only
// `keys`) touches exactly the first two sectors of one L2 line; the LRU stamps and the insert lock
// live in the other half of the same line (same DRAM page as the keys they describe).
struct alignas(128) Bucket {
  int64_t keys[kWays];     // 64 B, read by every probe
  uint32_t stamp[kWays];   // 32 B, lookup epoch of the last hit / insert (LRU)
  uint32_t lock;           // spin lock taken by the insert kernel only
  uint32_t pad[7];
};
static_assert(sizeof(Bucket) == 128, "bucket must be one 128-B line");

// One slot of the HBM-resident index of a page-locked host table ("direct pull" mode): key -> address
// of the row in mapped pinned host memory.  Open addressing, linear probing, power-of-two capacity.
struct alignas(16) IndexSlot {
  int64_t key;           // kEmptyKey = free
  const float* row;      // device-visible address of the row in host memory
};
static_assert(sizeof(IndexSlot) == 16, "index slot must be 16 B");

// murmur3 64-bit finalizer.
HPSX_HD uint64_t mix64(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 33;
  return h;
}

// Bucket of a key: high 32 hash bits range-reduced by multiply-shift (no power-of-two constraint).
HPSX_HD uint32_t bucket_of(int64_t key, uint32_t num_buckets) {
  const uint64_t h = mix64(static_cast<uint64_t>(key));
  return static_cast<uint32_t>(((h >> 32) * static_cast<uint64_t>(num_buckets)) >> 32);
}

// Second-choice bucket (two-choice hashing): an independent hash of the key.  A key lives in bucket_of() unless
// that bucket was full when it was inserted; ways are never emptied, so a probe that misses in a primary
// bucket with a free way knows the key is not cached and reads nothing else.
HPSX_HD uint32_t bucket2_of(int64_t key, uint32_t num_buckets) {
  const uint64_t h = mix64(static_cast<uint64_t>(key) ^ 0x9E3779B97F4A7C15ULL);
  return static_cast<uint32_t>(((h >> 32) * static_cast<uint64_t>(num_buckets)) >> 32);
}

// Partition of the HOST table a key lives in (host_ps.hpp): high 32 hash bits, multiply-shift.  Shared with the
// device because a partition doubles as the locality bin of the direct-pull miss lists.
HPSX_HD uint32_t host_partition_of_hash(uint64_t h, uint32_t num_partitions) {
  return static_cast<uint32_t>(((h >> 32) * static_cast<uint64_t>(num_partitions)) >> 32);
}
HPSX_HD uint32_t host_partition_of(int64_t key, uint32_t num_partitions) {
  return host_partition_of_hash(mix64(static_cast<uint64_t>(key)), num_partitions);
}

// Owning shard of a key in the model-parallel mode: low 32 hash bits, so that the bucket index
// (high bits) stays uniform inside one shard.
HPSX_HD uint32_t owner_of(int64_t key, uint32_t num_shards) {
  const uint64_t h = mix64(static_cast<uint64_t>(key));
  return static_cast<uint32_t>(((h & 0xffffffffULL) * static_cast<uint64_t>(num_shards)) >> 32);
}

HPSX_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

// Synthetic row element: uniform in [-0.5, 0.5), exact in fp32, identical on
// CPU and GPU because it is built from integer ops and one exact subtraction.
HPSX_HD float synth_value(int64_t key, uint32_t j, uint64_t seed) {
  const uint64_t r = splitmix64(static_cast<uint64_t>(key) * 131ULL + j + seed);
  const uint32_t bits = 0x3F800000u | static_cast<uint32_t>(r >> 41);
#if defined(__CUDA_ARCH__)
  return __uint_as_float(bits) - 1.5f;
#else
  union {
    uint32_t u;
    float f;
  } c;
  c.u = bits;
  return c.f - 1.5f;
#endif
}

}  // namespace hpsx
