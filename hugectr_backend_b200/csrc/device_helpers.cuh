// Device helpers shared by the kernel translation units (kernels.cu, shard_kernels.cu): vector loads/stores with
// cache hints, the two-choice bucket probe, the warp-aggregated miss list append.  Everything is
// __device__ __forceinline__ inside an anonymous namespace, so including it in several .cu files is safe.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace hpsx {
namespace {

constexpr int kRouteChunk = 2048;  // keys per CTA of the routing / dispatch kernels

constexpr unsigned kFull = 0xffffffffu;
constexpr int kBlock = 256;  // 8 warps; one warp owns one tile of 32 keys

// ------------------------------------------------------------------------------------------------
// vector load/store helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, const float2& v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, const float& v) { __stcs(p, v); }

template <typename VecT>
__device__ __forceinline__ VecT splat(float x);
template <>
__device__ __forceinline__ float4 splat<float4>(float x) {
  return make_float4(x, x, x, x);
}
template <>
__device__ __forceinline__ float2 splat<float2>(float x) {
  return make_float2(x, x);
}
template <>
__device__ __forceinline__ float splat<float>(float x) {
  return x;
}

__device__ __forceinline__ void vadd(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}
__device__ __forceinline__ void vadd(float2& a, const float2& b) {
  a.x += b.x;
  a.y += b.y;
}
__device__ __forceinline__ void vadd(float& a, const float& b) { a += b; }
__device__ __forceinline__ void vdiv(float4& a, float d) {
  a.x = __fdiv_rn(a.x, d);
  a.y = __fdiv_rn(a.y, d);
  a.z = __fdiv_rn(a.z, d);
  a.w = __fdiv_rn(a.w, d);
}
__device__ __forceinline__ void vdiv(float2& a, float d) {
  a.x = __fdiv_rn(a.x, d);
  a.y = __fdiv_rn(a.y, d);
}
__device__ __forceinline__ void vdiv(float& a, float d) { a = __fdiv_rn(a, d); }

// ------------------------------------------------------------------------------------------------
// probe: one thread, one key, one 64-B bucket (two DRAM sectors, four LDG.128 through L2 only); the
// second-choice bucket is read only when the key is not in a FULL primary bucket
// ------------------------------------------------------------------------------------------------
struct BucketKeys {
  longlong2 k01, k23, k45, k67;
};

__device__ __forceinline__ BucketKeys load_bucket_keys(const Bucket* __restrict__ buckets, uint32_t b) {
  const longlong2* kp = reinterpret_cast<const longlong2*>(buckets[b].keys);
  BucketKeys r;
  r.k01 = __ldcg(kp + 0);
  r.k23 = __ldcg(kp + 1);
  r.k45 = __ldcg(kp + 2);
  r.k67 = __ldcg(kp + 3);
  return r;
}

__device__ __forceinline__ int match_way(const BucketKeys& k, int64_t key) {
  int way = -1;
  way = (k.k01.x == key) ? 0 : way;
  way = (k.k01.y == key) ? 1 : way;
  way = (k.k23.x == key) ? 2 : way;
  way = (k.k23.y == key) ? 3 : way;
  way = (k.k45.x == key) ? 4 : way;
  way = (k.k45.y == key) ? 5 : way;
  way = (k.k67.x == key) ? 6 : way;
  way = (k.k67.y == key) ? 7 : way;
  return way;
}

// LRU touch: the stamp lives in the third sector of the bucket's line.  It is read first and written only when it
// changes: under skewed keys (Zipf 1.05: one key is 10 % of a request) tens of thousands of warps would otherwise
// store to the same word — measured 1.01 ms vs 0.26 ms for the probe kernel of such a request — while a read of a
// hot line is served by L2 without serialising.
__device__ __forceinline__ void touch_stamp(uint32_t* stamp, uint32_t epoch) {
  if (__ldcg(stamp) != epoch) *stamp = epoch;
}

__device__ __forceinline__ bool bucket_full(const BucketKeys& k) {
  return k.k01.x != kEmptyKey && k.k01.y != kEmptyKey && k.k23.x != kEmptyKey && k.k23.y != kEmptyKey &&
         k.k45.x != kEmptyKey && k.k45.y != kEmptyKey && k.k67.x != kEmptyKey && k.k67.y != kEmptyKey;
}

// Second half of a probe whose primary bucket `bk` (index b) is already in registers.
__device__ __forceinline__ uint32_t resolve_slot(Bucket* __restrict__ buckets, uint32_t num_buckets, int64_t key,
                                                 uint32_t b, const BucketKeys& bk, uint32_t epoch, bool touch) {
  int way = match_way(bk, key);
  if (way < 0) {
    if (!bucket_full(bk)) return kMissSlot;
    b = bucket2_of(key, num_buckets);
    way = match_way(load_bucket_keys(buckets, b), key);
    if (way < 0) return kMissSlot;
  }
  if (touch) touch_stamp(&buckets[b].stamp[way], epoch);
  return b * kWays + static_cast<uint32_t>(way);
}

__device__ __forceinline__ uint32_t probe_bucket(Bucket* __restrict__ buckets, uint32_t num_buckets,
                                                 int64_t key, uint32_t epoch, bool touch) {
  if (key == kEmptyKey) return kMissSlot;
  const uint32_t b = bucket_of(key, num_buckets);
  return resolve_slot(buckets, num_buckets, key, b, load_bucket_keys(buckets, b), epoch, touch);
}

struct ProbeArgs {
  Bucket* buckets;
  const float* values;
  uint32_t num_buckets;
  uint32_t dim;
  float default_value;
  const int64_t* keys;
  size_t n;
  float* out;
  uint32_t epoch;
  int touch;
  uint32_t* miss_count;
  uint32_t* miss_pos;
  int64_t* miss_keys;
  int64_t* miss_keys_host;  // optional mirror in mapped pinned host memory (zero-copy PCIe writes)
  uint32_t* src;  // probe_index only
  const uint32_t* pos;  // optional: key i is delivered to row pos[i] of `out` (which may be peer memory)
  uint32_t pos_base;    // added to the positions recorded in the miss list (this launch covers keys [pos_base, pos_base + n))
  __nv_bfloat16* out_bf16;  // optional mirror of `out` in bf16 (same row order), feeds the dense head without a conversion pass
  MissBins bins;        // bins.count != nullptr: misses go to the binned lists instead of miss_count/miss_pos/miss_keys
  int skip_miss_rows;   // the rows of missed keys are left unwritten (a pull kernel delivers them: synchronous insertion)
};

// Append the misses of one warp tile to the global miss list: ballot -> popc prefix -> one atomic.
__device__ __forceinline__ uint32_t warp_claim_misses(bool is_miss, uint32_t lane,
                                                      uint32_t* miss_count, unsigned* mask_out) {
  const unsigned mask = __ballot_sync(kFull, is_miss);
  *mask_out = mask;
  uint32_t base = 0;
  if (mask != 0u) {
    if (lane == 0) base = atomicAdd(miss_count, static_cast<uint32_t>(__popc(mask)));
    base = __shfl_sync(kFull, base, 0);
  }
  return base;
}

// Binned form: the miss goes to the list of its host-table partition (one atomic per missing lane; a bin that is
// full spills into the shared overflow list).
__device__ __forceinline__ void append_miss_binned(const MissBins& bins, int64_t key, uint32_t pos) {
  const uint32_t b = host_partition_of(key, bins.num_bins);
  const uint32_t j = atomicAdd(&bins.count[b], 1u);
  size_t r;
  if (j < bins.bin_cap)
    r = static_cast<size_t>(b) * bins.bin_cap + j;
  else
    r = static_cast<size_t>(bins.num_bins) * bins.bin_cap + atomicAdd(&bins.count[bins.num_bins], 1u);
  bins.keys[r] = key;
  bins.pos[r] = pos;
}

inline unsigned grid_for(size_t threads) {
  return static_cast<unsigned>((threads + kBlock - 1) / kBlock);
}

}  // namespace
}  // namespace hpsx
