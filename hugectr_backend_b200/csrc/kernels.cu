// sm_100a kernels of the HPS lookup hot path.
//
// None of this is a contraction: every kernel is HBM-bound integer hashing + row copies, so the
// design rules are coalescing, 16-B vector accesses, many independent loads in flight per warp,
// bulk-async (TMA engine) row staging where rows are >= 16 B, and warp ballot/prefix compaction.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"
#include "device_helpers.cuh"

namespace hpsx {
namespace {

// ------------------------------------------------------------------------------------------------
// K2 (+K3,K6) LDG variant.  Warp = tile of 32 keys.  Phase 1: lane-per-key probe.  Phase 2: the
// warp copies the 32 rows cooperatively, kUnroll independent 16-B loads per lane in flight, output
// addresses contiguous across the whole tile (rows of consecutive keys are adjacent in `out`).
// kV = vectors per row when known at compile time (0: runtime).
// ------------------------------------------------------------------------------------------------
// kScatter: every key carries its destination row (`pos`), and `out` may be another GPU's buffer mapped
// through NVLink peer access — the owner's gather kernel then IS the return leg of the model-parallel
// exchange (rows leave as 512-B peer stores; no all-to-all of vectors, no scatter pass).
template <typename VecT, int kV, int kUnroll, bool kScatter = false>
__global__ void __launch_bounds__(kBlock) probe_gather_ldg_kernel(const ProbeArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t tile = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t tile_base = tile * 32;
  if (tile_base >= a.n) return;
  const uint32_t nk = static_cast<uint32_t>(min(static_cast<size_t>(32), a.n - tile_base));
  const uint32_t V = kV > 0 ? static_cast<uint32_t>(kV)
                            : a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));

  const bool valid = lane < nk;
  const int64_t key = valid ? a.keys[tile_base + lane] : kEmptyKey;
  uint32_t slot = kMissSlot;
  if (valid) slot = probe_bucket(a.buckets, a.num_buckets, key, a.epoch, a.touch != 0);
  const bool is_miss = valid && slot == kMissSlot;
  unsigned miss_mask = 0;
  uint32_t miss_base = 0;
  if (a.bins.count == nullptr) miss_base = warp_claim_misses(is_miss, lane, a.miss_count, &miss_mask);

  const VecT* __restrict__ vals = reinterpret_cast<const VecT*>(a.values);
  VecT* __restrict__ outv = reinterpret_cast<VecT*>(a.out) + (kScatter ? 0 : tile_base * V);
  const uint32_t dst = (kScatter && valid) ? a.pos[tile_base + lane] : static_cast<uint32_t>(tile_base + lane);
  const VecT defv = splat<VecT>(a.default_value);
  const uint32_t total = nk * V;
  const bool skip = a.skip_miss_rows != 0;
  for (uint32_t i0 = 0; i0 < total; i0 += 32u * kUnroll) {
    VecT buf[kUnroll];
    bool live[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      const uint32_t kk = min(i / V, 31u);
      const uint32_t s = __shfl_sync(kFull, slot, kk);
      const uint32_t v = i - kk * V;
      buf[u] = defv;
      live[u] = i < total && (s != kMissSlot || !skip);
      if (i < total && s != kMissSlot) buf[u] = ld_stream(vals + static_cast<size_t>(s) * V + v);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      if (kScatter) {
        const uint32_t kk = min(i / V, 31u);
        const uint32_t d = __shfl_sync(kFull, dst, kk);
        if (live[u]) st_stream(outv + static_cast<size_t>(d) * V + (i - kk * V), buf[u]);
      } else {
        if (live[u]) st_stream(outv + i, buf[u]);
      }
    }
  }

  if (is_miss) {
    if (a.bins.count != nullptr) {
      append_miss_binned(a.bins, key, kScatter ? dst : dst + a.pos_base);
    } else {
      const uint32_t r = miss_base + __popc(miss_mask & ((1u << lane) - 1u));
      a.miss_pos[r] = kScatter ? dst : dst + a.pos_base;
      a.miss_keys[r] = key;
      if (a.miss_keys_host != nullptr) a.miss_keys_host[r] = key;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2 256-bit variant (sm_100 LDG/STG.256 with inline L2 eviction priorities).  Same tile shape as the LDG
// variant, but (a) rows move as 32-B vectors marked L2::evict_first on both the load and the store, (b) the
// 64-B key block of a bucket is two 32-B loads marked L2::evict_last.  The bucket arrays of a hot table
// (tens of MB) then stay resident in the 126 MB L2 while ~1.7 GB of rows stream through it per request, so a
// probe no longer costs a DRAM activation of its own.
// ------------------------------------------------------------------------------------------------
struct alignas(32) Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 ld_stream256(const Vec8* p) {
  Vec8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                 "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream256(Vec8* p, const Vec8& r) {
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
               "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 8 floats -> 8 bf16 (16 B), streaming store
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const Vec8& r) {
  __stcs(reinterpret_cast<uint4*>(p), make_uint4(pack_bf16x2(r.v[0], r.v[1]), pack_bf16x2(r.v[2], r.v[3]),
                                                 pack_bf16x2(r.v[4], r.v[5]), pack_bf16x2(r.v[6], r.v[7])));
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, const float4& r) {
  __stcs(reinterpret_cast<uint2*>(p), make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w)));
}
__device__ __forceinline__ BucketKeys load_bucket_keys_keep(const Bucket* __restrict__ buckets, uint32_t b) {
  BucketKeys r;
  const long long* kp = reinterpret_cast<const long long*>(buckets[b].keys);
  asm volatile("ld.global.L1::no_allocate.L2::evict_last.v4.s64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.k01.x), "=l"(r.k01.y), "=l"(r.k23.x), "=l"(r.k23.y)
               : "l"(kp));
  asm volatile("ld.global.L1::no_allocate.L2::evict_last.v4.s64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.k45.x), "=l"(r.k45.y), "=l"(r.k67.x), "=l"(r.k67.y)
               : "l"(kp + 4));
  return r;
}

// Measured alternatives that are NOT in the library any more (profiles/probe_experiments_r02.jsonl; the code is in the
// history: commit "Probe kernel experiments (L2 row prefetch, deeper unroll)"): asking L2 for the whole row right after
// the probe (`prefetch.global.L2` per line: 0.71 of the HBM peak; one `cp.async.bulk.prefetch.L2` per row: 0.73) and
// 8 x 32 B per lane in flight (0.64) all lose to this form (0.82).
template <int kV8, int kUnroll, bool kMirror = false>
__global__ void __launch_bounds__(kBlock) probe_gather_v8_kernel(const ProbeArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t tile = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t tile_base = tile * 32;
  if (tile_base >= a.n) return;
  const uint32_t nk = static_cast<uint32_t>(min(static_cast<size_t>(32), a.n - tile_base));
  const uint32_t V = kV8 > 0 ? static_cast<uint32_t>(kV8) : a.dim / 8u;

  const bool valid = lane < nk;
  const int64_t key = valid ? a.keys[tile_base + lane] : kEmptyKey;
  uint32_t slot = kMissSlot;
  if (valid && key != kEmptyKey) {
    uint32_t b = bucket_of(key, a.num_buckets);
    BucketKeys bk = load_bucket_keys_keep(a.buckets, b);
    int way = match_way(bk, key);
    if (way < 0 && bucket_full(bk)) {
      b = bucket2_of(key, a.num_buckets);
      bk = load_bucket_keys_keep(a.buckets, b);
      way = match_way(bk, key);
    }
    if (way >= 0) slot = b * kWays + static_cast<uint32_t>(way);
  }
  const bool is_miss = valid && slot == kMissSlot;
  unsigned miss_mask = 0;
  uint32_t miss_base = 0;
  if (a.bins.count == nullptr) miss_base = warp_claim_misses(is_miss, lane, a.miss_count, &miss_mask);
  const bool skip = a.skip_miss_rows != 0;

  const Vec8* __restrict__ vals = reinterpret_cast<const Vec8*>(a.values);
  Vec8* __restrict__ outv = reinterpret_cast<Vec8*>(a.out) + tile_base * V;
  Vec8 defv;
#pragma unroll
  for (int j = 0; j < 8; ++j) defv.v[j] = a.default_value;
  const uint32_t total = nk * V;
  for (uint32_t i0 = 0; i0 < total; i0 += 32u * kUnroll) {
    Vec8 buf[kUnroll];
    bool live[kUnroll];  // false: the row of a missed key that the pull kernel will deliver (nothing is stored here)
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      const uint32_t kk = min(i / V, 31u);
      const uint32_t s = __shfl_sync(kFull, slot, kk);
      buf[u] = defv;
      live[u] = i < total && (s != kMissSlot || !skip);
      if (i < total && s != kMissSlot) buf[u] = ld_stream256(vals + static_cast<size_t>(s) * V + (i - kk * V));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      if (live[u]) st_stream256(outv + i, buf[u]);
    }
    if (kMirror) {
      __nv_bfloat16* __restrict__ ob = a.out_bf16 + tile_base * static_cast<size_t>(V) * 8u;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = i0 + u * 32u + lane;
        if (live[u]) st_bf16x8(ob + static_cast<size_t>(i) * 8u, buf[u]);
      }
    }
  }
  // LRU touch last: its read-compare-store chain (device_helpers.cuh touch_stamp) must not sit between the probe and
  // the row loads of the tile
  if (a.touch && valid && slot != kMissSlot) touch_stamp(&a.buckets[slot / kWays].stamp[slot % kWays], a.epoch);
  if (is_miss) {
    const uint32_t p = static_cast<uint32_t>(tile_base + lane) + a.pos_base;
    if (a.bins.count != nullptr) {
      append_miss_binned(a.bins, key, p);
    } else {
      const uint32_t r = miss_base + __popc(miss_mask & ((1u << lane) - 1u));
      a.miss_pos[r] = p;
      a.miss_keys[r] = key;
      if (a.miss_keys_host != nullptr) a.miss_keys_host[r] = key;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2 TMA variant: rows are staged global -> shared with cp.async.bulk (UBLKCP, the TMA engine's
// 1-D path) and leave with ONE bulk store per tile, because the 32 output rows of a tile are
// contiguous.  No row data passes through registers.  Ring of kStages tiles per warp-group-free
// CTA: every warp owns its own ring slot sequence, so there is no CTA-wide barrier on the hot path.
// Requires dim*4 % 16 == 0 and 16-B aligned `out`.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr int kTmaTileKeys = 32;

// Dynamic smem: [warp][stage][32 rows * row_bytes] then one mbarrier per (warp, stage).
// Schedule per warp: iteration t issues the loads of tile t into stage t % kStages, then retires
// tile t-1 (wait for its bytes, one bulk store).  The previous user of stage t % kStages is tile
// t-kStages, whose store was committed in iteration t-kStages+1, i.e. it is the (kStages-1)-th most
// recent bulk group at the top of iteration t: wait_group.read <kStages-2> frees it.
template <int kWarps, int kStages>
__global__ void __launch_bounds__(kWarps * 32) probe_gather_tma_kernel(const ProbeArgs a,
                                                                       uint32_t tiles_per_warp) {
  static_assert(kStages >= 2, "need a ring");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t row_bytes = a.dim * 4u;
  const uint32_t tile_bytes = kTmaTileKeys * row_bytes;
  unsigned char* ring = smem_raw + static_cast<size_t>(warp) * kStages * tile_bytes;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(kWarps) * kStages * tile_bytes) +
      warp * kStages;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const size_t warp_global = static_cast<size_t>(blockIdx.x) * kWarps + warp;
  const size_t first_tile = warp_global * tiles_per_warp;
  const size_t num_tiles = (a.n + kTmaTileKeys - 1) / kTmaTileKeys;
  const float4 defv = make_float4(a.default_value, a.default_value, a.default_value, a.default_value);
  const unsigned char* vals = reinterpret_cast<const unsigned char*>(a.values);
  unsigned char* outb = reinterpret_cast<unsigned char*>(a.out);

  uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s
  uint32_t pend_nk = 0;
  size_t pend_base = 0;
  int pend_stage = -1;
  for (uint32_t t = 0; t <= tiles_per_warp; ++t) {
    const size_t tile = first_tile + t;
    const bool have = t < tiles_per_warp && tile < num_tiles;
    const int stage = static_cast<int>(t % kStages);
    uint32_t nk = 0;
    const size_t tile_base = tile * kTmaTileKeys;
    if (have) {
      nk = static_cast<uint32_t>(min(static_cast<size_t>(kTmaTileKeys), a.n - tile_base));
      const bool valid = lane < nk;
      const int64_t key = valid ? a.keys[tile_base + lane] : kEmptyKey;
      uint32_t slot = kMissSlot;
      if (valid) slot = probe_bucket(a.buckets, a.num_buckets, key, a.epoch, a.touch != 0);
      const bool is_miss = valid && slot == kMissSlot;
      unsigned miss_mask;
      const uint32_t miss_base = warp_claim_misses(is_miss, lane, a.miss_count, &miss_mask);
      unsigned char* stage_buf = ring + static_cast<size_t>(stage) * tile_bytes;
      const unsigned hit_mask = __ballot_sync(kFull, valid && !is_miss);
      // the bulk store that last read this stage must have finished reading shared memory
      if (lane == 0) {
        bulk_wait_read<kStages - 2>();
        mbar_expect_tx(&bars[stage], static_cast<uint32_t>(__popc(hit_mask)) * row_bytes);
      }
      __syncwarp();
      if (valid && !is_miss) {
        bulk_g2s(stage_buf + lane * row_bytes, vals + static_cast<size_t>(slot) * row_bytes,
                 row_bytes, &bars[stage]);
      }
      if (miss_mask != 0u) {
        // default vectors for the missing rows, written by the whole warp through the generic proxy
        const uint32_t vec_per_row = row_bytes / 16u;
        for (uint32_t m = miss_mask; m != 0u; m &= m - 1u) {
          const uint32_t kk = __ffs(m) - 1u;
          float4* dst = reinterpret_cast<float4*>(stage_buf + kk * row_bytes);
          for (uint32_t v = lane; v < vec_per_row; v += 32u) dst[v] = defv;
        }
        if (is_miss) {
          const uint32_t r = miss_base + __popc(miss_mask & ((1u << lane) - 1u));
          a.miss_pos[r] = static_cast<uint32_t>(tile_base + lane) + a.pos_base;
          a.miss_keys[r] = key;
          if (a.miss_keys_host != nullptr) a.miss_keys_host[r] = key;
        }
        fence_proxy_async_smem();  // generic-proxy default rows -> visible to the later bulk store
      }
    }
    // retire the previous tile
    if (pend_stage >= 0) {
      mbar_wait(&bars[pend_stage], (phase_bits >> pend_stage) & 1u);
      phase_bits ^= 1u << pend_stage;
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(outb + pend_base * row_bytes, ring + static_cast<size_t>(pend_stage) * tile_bytes,
                 pend_nk * row_bytes);
        bulk_commit();
      }
      pend_stage = -1;
    }
    if (have) {
      pend_stage = stage;
      pend_nk = nk;
      pend_base = tile_base;
    }
  }
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the last bulk store's reads
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// probe only (pooled path): record where each key's row lives.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) probe_index_kernel(const ProbeArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t idx = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  const bool valid = idx < a.n;  // warp tiles are aligned, the whole warp runs the ballot
  const int64_t key = valid ? a.keys[idx] : kEmptyKey;
  uint32_t slot = kMissSlot;
  if (valid) slot = probe_bucket(a.buckets, a.num_buckets, key, a.epoch, a.touch != 0);
  const bool is_miss = valid && slot == kMissSlot;
  unsigned miss_mask;
  const uint32_t miss_base = warp_claim_misses(is_miss, lane, a.miss_count, &miss_mask);
  if (is_miss) {
    const uint32_t r = miss_base + __popc(miss_mask & ((1u << lane) - 1u));
    a.miss_pos[r] = static_cast<uint32_t>(idx) + a.pos_base;
    a.miss_keys[r] = key;
    if (a.miss_keys_host != nullptr) a.miss_keys_host[r] = key;
    slot = kSrcMissBit | r;
  }
  if (valid) a.src[idx] = slot;
}

// ------------------------------------------------------------------------------------------------
// K4+K5 fused merge + insert.  One warp per missing key.
// ------------------------------------------------------------------------------------------------
struct InsertArgs {
  Bucket* buckets;
  float* values;
  uint32_t num_buckets;
  uint32_t dim;
  const int64_t* miss_keys;
  const uint32_t* miss_pos;
  const float* stage;
  size_t m;
  float* out;
  int insert;
  uint32_t epoch;
  uint32_t* inserted;
  __nv_bfloat16* out_bf16;  // optional bf16 mirror of `out` (float4 instantiation only)
};

// Claims the cache slot for `key`.  Called by a whole warp.  Both candidate buckets (primary and second
// choice) are locked in ascending index order, so concurrent inserts can neither deadlock nor place one key
// twice.  Result: slot index to write, or kMissSlot when the key is already resident (LRU refreshed) or nothing
// may be evicted (all 16 ways were touched in this very epoch).  Placement: a free way of the primary bucket,
// else a free way of the second choice, else the way with the oldest stamp of the 16 (primary wins ties).
//
// The critical section covers the bucket's keys and stamps only, and the locks are RELEASED before this returns:
// the caller then writes the row without any lock.  That is safe because a way claimed in this epoch carries
// stamp == epoch and is never evicted by another claim of the same kernel, a second claim of the same key finds
// it present and writes nothing, and rows are read only by probes, which never run beside an inserting kernel
// (host lock of the cache).  The chain per insert is lock -> load -> store + release, without waiting for the
// row (which may still be on its way over PCIe or NVLink) — a per-row fence used to cost several microseconds.
__device__ __forceinline__ void lock_bucket(Bucket* B) {
  uint32_t old;
  do {
    asm volatile("atom.acquire.gpu.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(&B->lock) : "memory");
    if (old != 0u) __nanosleep(32);
  } while (old != 0u);
}
// Both locks of a claim with ONE fence (a release store is MEMBAR + store: two of them were a quarter of the quad pull's
// stall cycles); also gives back a lock that was taken by a failed try (nothing was written under it: no fence at all).
__device__ __forceinline__ void unlock_bucket_relaxed(Bucket* B) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&B->lock), "r"(0u) : "memory");
}
__device__ __forceinline__ void unlock_pair(Bucket* lo, Bucket* hi) {
  __threadfence();
  if (hi != nullptr) unlock_bucket_relaxed(hi);
  unlock_bucket_relaxed(lo);
}

__device__ __forceinline__ uint32_t claim_slot(Bucket* buckets, uint32_t num_buckets, int64_t key, uint32_t epoch,
                                               uint32_t lane) {
  const uint32_t b1 = bucket_of(key, num_buckets);
  const uint32_t b2 = bucket2_of(key, num_buckets);
  Bucket* lo = &buckets[min(b1, b2)];
  Bucket* hi = b1 == b2 ? nullptr : &buckets[max(b1, b2)];
  if (lane == 0) {
    lock_bucket(lo);
    if (hi != nullptr) lock_bucket(hi);
  }
  __syncwarp();
  // lanes 0-7: ways of the primary bucket, lanes 8-15: ways of the second choice
  const uint32_t nways = b1 == b2 ? kWays : 2 * kWays;
  const uint32_t my_b = lane < kWays ? b1 : b2;
  const uint32_t my_w = lane & (kWays - 1);
  int64_t k = kEmptyKey;
  uint32_t st = epoch;
  if (lane < nways) {
    k = __ldcg(reinterpret_cast<const long long*>(&buckets[my_b].keys[my_w]));
    st = __ldcg(&buckets[my_b].stamp[my_w]);
  }
  const unsigned present = __ballot_sync(kFull, lane < nways && k == key);
  int pick = -1;  // lane index of the chosen way
  bool fresh = false;
  if (present != 0u) {
    pick = __ffs(present) - 1;  // already cached: refresh LRU only
  } else {
    const unsigned empties = __ballot_sync(kFull, lane < nways && k == kEmptyKey);
    if (empties != 0u) {
      pick = __ffs(empties) - 1;  // primary ways come first
      fresh = true;
    } else {
      // oldest stamp wins; ways touched in this very epoch (age 0) are never evicted.  Instances that share the
      // cache run with interleaved epochs: a stamp NEWER than this call's epoch is age 0 too (signed difference),
      // not a huge unsigned age that would make the most recently used rows the first victims.
      const int32_t diff = static_cast<int32_t>(epoch - st);
      const uint32_t age = (lane < nways && diff > 0) ? static_cast<uint32_t>(diff) : 0u;
      unsigned long long packed = (static_cast<unsigned long long>(age) << 8) | (255u - lane);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFull, packed, off);
        packed = o > packed ? o : packed;
      }
      packed = __shfl_sync(kFull, packed, 0);
      if ((packed >> 8) != 0ull) {
        pick = 255 - static_cast<int>(packed & 255ull);
        fresh = true;
      }
    }
  }
  const uint32_t pb = pick < kWays ? b1 : b2;
  const uint32_t pw = static_cast<uint32_t>(pick < 0 ? 0 : pick) & (kWays - 1);
  if (lane == 0) {
    if (pick >= 0) {
      if (fresh) buckets[pb].keys[pw] = key;
      buckets[pb].stamp[pw] = epoch;
    }
    unlock_pair(lo, hi);
  }
  return fresh ? pb * kWays + pw : kMissSlot;
}

// Four claims at once (quad form of the tier pull): lanes 8e..8e+7 work on key[e] — lane 8e takes the locks, lane
// 8e+w looks at way w of both candidate buckets.  The lock / load / store chain, a handful of dependent L2 round trips,
// is paid once per quad instead of once per row.  Entries of one quad that share a bucket (the same key twice, or a
// bucket collision) cannot hold their locks side by side: every entry that shares a bucket with a LOWER entry of the
// quad is left to the caller (returned in *deferred as a bit mask) to claim one by one with claim_slot.  The locks of
// a quad are only TRIED: a warp that waited for one lock while its other lanes hold theirs would break the ascending
// lock order that keeps the blocking form deadlock-free (warp A holds 100 and waits for 50, warp B holds 50 and waits
// for 100); an entry whose try fails gives back what it took and is deferred as well.
// `want`: this lane's entry is to be claimed (uniform within its 8-lane group).  Returns the slot of this lane's entry.
__device__ __forceinline__ uint32_t claim_slot4(Bucket* buckets, uint32_t num_buckets, int64_t key, bool want, uint32_t epoch,
                                                uint32_t lane, unsigned* deferred) {
  const uint32_t sub = lane & 7u, shift = lane & ~7u, e = lane >> 3;
  const uint32_t b1 = bucket_of(key, num_buckets);
  const uint32_t b2 = bucket2_of(key, num_buckets);
  // conflicts with lower entries of the quad
  bool conflict = false;
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    const uint32_t ob1 = __shfl_sync(kFull, b1, 8 * o), ob2 = __shfl_sync(kFull, b2, 8 * o);
    const bool owant = __shfl_sync(kFull, want ? 1 : 0, 8 * o) != 0;
    if (static_cast<uint32_t>(o) < e && owant && (ob1 == b1 || ob1 == b2 || ob2 == b1 || ob2 == b2)) conflict = true;
  }
  const unsigned conflict_lanes = __ballot_sync(kFull, want && conflict);
  *deferred = ((conflict_lanes >> 0) & 1u) | (((conflict_lanes >> 8) & 1u) << 1) | (((conflict_lanes >> 16) & 1u) << 2) |
              (((conflict_lanes >> 24) & 1u) << 3);
  Bucket* lo = &buckets[min(b1, b2)];
  Bucket* hi = b1 == b2 ? nullptr : &buckets[max(b1, b2)];
  bool got = false;
  if (want && !conflict && sub == 0) {
    // both tries are issued before either result is looked at: one L2 round trip instead of two
    uint32_t old_lo = 0, old_hi = 0;
    asm volatile("atom.acquire.gpu.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old_lo) : "l"(&lo->lock) : "memory");
    if (hi != nullptr) asm volatile("atom.acquire.gpu.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old_hi) : "l"(&hi->lock) : "memory");
    got = old_lo == 0u && old_hi == 0u;
    if (!got) {  // give back whichever was taken; nothing was written under it
      if (old_lo == 0u) unlock_bucket_relaxed(lo);
      if (hi != nullptr && old_hi == 0u) unlock_bucket_relaxed(hi);
    }
  }
  got = __shfl_sync(kFull, got ? 1 : 0, shift) != 0;  // the lead lane's result, for its whole group
  __syncwarp();
  const unsigned busy_lanes = __ballot_sync(kFull, want && !conflict && !got);
  *deferred |= ((busy_lanes >> 0) & 1u) | (((busy_lanes >> 8) & 1u) << 1) | (((busy_lanes >> 16) & 1u) << 2) |
               (((busy_lanes >> 24) & 1u) << 3);
  const bool active = want && !conflict && got;
  int64_t k1 = kEmptyKey, k2 = kEmptyKey;
  uint32_t s1 = epoch, s2 = epoch;
  const bool two = b1 != b2;
  if (active) {
    k1 = __ldcg(reinterpret_cast<const long long*>(&buckets[b1].keys[sub]));
    s1 = __ldcg(&buckets[b1].stamp[sub]);
    if (two) {
      k2 = __ldcg(reinterpret_cast<const long long*>(&buckets[b2].keys[sub]));
      s2 = __ldcg(&buckets[b2].stamp[sub]);
    }
  }
  const unsigned p1 = (__ballot_sync(kFull, active && k1 == key) >> shift) & 0xffu;
  const unsigned p2 = (__ballot_sync(kFull, active && two && k2 == key) >> shift) & 0xffu;
  const unsigned e1 = (__ballot_sync(kFull, active && k1 == kEmptyKey) >> shift) & 0xffu;
  const unsigned e2 = (__ballot_sync(kFull, active && two && k2 == kEmptyKey) >> shift) & 0xffu;
  // oldest stamp of the 16 ways (primary wins ties, lower way wins ties): same order as claim_slot
  const int32_t d1 = static_cast<int32_t>(epoch - s1), d2 = static_cast<int32_t>(epoch - s2);
  const uint32_t a1 = (active && d1 > 0) ? static_cast<uint32_t>(d1) : 0u;
  const uint32_t a2 = (active && two && d2 > 0) ? static_cast<uint32_t>(d2) : 0u;
  unsigned long long packed1 = (static_cast<unsigned long long>(a1) << 8) | (255u - sub);
  unsigned long long packed2 = (static_cast<unsigned long long>(a2) << 8) | (255u - (8u + sub));
  unsigned long long packed = packed1 > packed2 ? packed1 : packed2;
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(kFull, packed, off);
    packed = o > packed ? o : packed;
  }
  int pick = -1;  // 0-7: way of the primary, 8-15: way of the second choice
  bool fresh = false;
  if (p1 != 0u) {
    pick = __ffs(p1) - 1;
  } else if (p2 != 0u) {
    pick = 8 + __ffs(p2) - 1;
  } else if (e1 != 0u) {
    pick = __ffs(e1) - 1;
    fresh = true;
  } else if (e2 != 0u) {
    pick = 8 + __ffs(e2) - 1;
    fresh = true;
  } else if ((packed >> 8) != 0ull) {
    pick = 255 - static_cast<int>(packed & 255ull);
    fresh = true;
  }
  if (!active) {
    pick = -1;
    fresh = false;
  }
  const uint32_t pb = pick < kWays ? b1 : b2;
  const uint32_t pw = static_cast<uint32_t>(pick < 0 ? 0 : pick) & (kWays - 1);
  if (active && sub == 0) {
    if (pick >= 0) {
      if (fresh) buckets[pb].keys[pw] = key;
      buckets[pb].stamp[pw] = epoch;
    }
    unlock_pair(lo, hi);
  }
  return fresh ? pb * kWays + pw : kMissSlot;
}

template <typename VecT>
__global__ void __launch_bounds__(kBlock) insert_merge_kernel(const InsertArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t nwarps = (static_cast<size_t>(gridDim.x) * kBlock) >> 5;
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  uint32_t n_inserted = 0;
  for (size_t i = warp; i < a.m; i += nwarps) {
    const int64_t key = a.miss_keys[i];
    // stage == nullptr: the row already sits in out[pos[i]] (pipelined direct pull) and is only inserted
    const VecT* src = a.stage ? reinterpret_cast<const VecT*>(a.stage) + i * V
                              : reinterpret_cast<const VecT*>(a.out) + static_cast<size_t>(a.miss_pos[i]) * V;
    VecT* dst_out = (a.out && a.stage)
                        ? reinterpret_cast<VecT*>(a.out) + static_cast<size_t>(a.miss_pos[i]) * V
                        : nullptr;
    VecT* dst_slab = nullptr;
    if (a.insert && key != kEmptyKey) {
      const uint32_t slot = claim_slot(a.buckets, a.num_buckets, key, a.epoch, lane);
      if (slot != kMissSlot) dst_slab = reinterpret_cast<VecT*>(a.values) + static_cast<size_t>(slot) * V;
    }
    for (uint32_t v = lane; v < V; v += 32u) {
      const VecT x = src[v];
      if (dst_out) dst_out[v] = x;
      if (dst_slab) dst_slab[v] = x;
      if constexpr (sizeof(VecT) == 16) {
        if (dst_out && a.out_bf16) st_bf16x4(a.out_bf16 + (static_cast<size_t>(a.miss_pos[i]) * V + v) * 4u, x);
      }
    }
    n_inserted += dst_slab != nullptr ? 1u : 0u;
  }
  // one atomic per warp: a counter bumped once per row serialises in its L2 slice (~2 ns per same-address atomic,
  // 0.35 ms for the 173 k misses of a DCN request — invisible behind PCIe, the whole cost of a pull from the NVLink tier)
  if (lane == 0 && n_inserted != 0u && a.inserted != nullptr) atomicAdd(a.inserted, n_inserted);
}

// K9 value update (cache refresh): overwrite the cached row of every key that is still
// resident; keys that left the cache meanwhile are skipped.  One warp per key; lanes 0-15 compare the 16
// candidate ways (primary + second-choice bucket).  Runs with the cache's host lock held exclusively, so no
// probe copies a row while it is rewritten.
template <typename VecT>
__global__ void __launch_bounds__(kBlock) update_values_kernel(const Bucket* __restrict__ buckets, float* values,
                                                               uint32_t num_buckets, uint32_t dim,
                                                               const int64_t* __restrict__ keys,
                                                               const float* __restrict__ stage, size_t n,
                                                               uint32_t* updated) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t nwarps = (static_cast<size_t>(gridDim.x) * kBlock) >> 5;
  const uint32_t V = dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  for (size_t i = warp; i < n; i += nwarps) {
    const int64_t key = keys[i];
    if (key == kEmptyKey) continue;
    const uint32_t b1 = bucket_of(key, num_buckets), b2 = bucket2_of(key, num_buckets);
    const uint32_t my_b = lane < kWays ? b1 : b2;
    const uint32_t my_w = lane & (kWays - 1);
    int64_t k = kEmptyKey;
    if (lane < 2 * kWays) k = __ldcg(reinterpret_cast<const long long*>(&buckets[my_b].keys[my_w]));
    const unsigned hit = __ballot_sync(kFull, lane < 2 * kWays && k == key);
    if (hit == 0u) continue;
    const uint32_t src_lane = __ffs(hit) - 1;
    const uint32_t slot = __shfl_sync(kFull, my_b * kWays + my_w, src_lane);
    const VecT* src = reinterpret_cast<const VecT*>(stage) + i * V;
    VecT* dst = reinterpret_cast<VecT*>(values) + static_cast<size_t>(slot) * V;
    for (uint32_t v = lane; v < V; v += 32u) dst[v] = src[v];
    if (lane == 0 && updated != nullptr) atomicAdd(updated, 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// direct pull: index of the page-locked host table in HBM, rows read zero-copy over PCIe
// ------------------------------------------------------------------------------------------------
__global__ void index_clear_kernel(IndexSlot* slots, uint64_t capacity) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < capacity) {
    // one 16-B store per slot
    const unsigned long long e = static_cast<unsigned long long>(kEmptyKey);
    *reinterpret_cast<uint4*>(&slots[i]) =
        make_uint4(static_cast<uint32_t>(e), static_cast<uint32_t>(e >> 32), 0u, 0u);
  }
}

__global__ void __launch_bounds__(kBlock) index_build_kernel(IndexSlot* slots, uint64_t mask,
                                                             const int64_t* __restrict__ keys,
                                                             const uint64_t* __restrict__ rows, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= n) return;
  const int64_t key = keys[i];
  if (key == kEmptyKey) return;  // carried separately (DeviceTable::sentinel_row)
  uint64_t slot = mix64(static_cast<uint64_t>(key)) & mask;
  while (true) {
    const unsigned long long prev =
        atomicCAS(reinterpret_cast<unsigned long long*>(&slots[slot].key),
                  static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
    if (prev == static_cast<unsigned long long>(kEmptyKey) || prev == static_cast<unsigned long long>(key)) {
      slots[slot].row = reinterpret_cast<const float*>(rows[i]);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------------------
// NVLink tier (kernels.h): the rows of the page-locked host table sharded over the HBM of the GPUs of one box.
// ------------------------------------------------------------------------------------------------
// One warp per (key, host row address) pair: the pairs this rank owns are appended to its shard, the row is copied
// out of mapped host memory (zero-copy PCIe reads, once, at build time).  Pairs past `cap` stay host-only.
__global__ void __launch_bounds__(kBlock) tier_fill_kernel(const int64_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ addrs, size_t n, uint32_t rank,
                                                           uint32_t world, uint32_t dim, int64_t* shard_keys,
                                                           float* shard_rows, unsigned long long cap,
                                                           unsigned long long* count) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t nwarps = (static_cast<size_t>(gridDim.x) * kBlock) >> 5;
  for (size_t i = warp; i < n; i += nwarps) {
    const int64_t key = keys[i];
    if (key == kEmptyKey || owner_of(key, world) != rank) continue;
    unsigned long long j = 0;
    if (lane == 0) j = atomicAdd(count, 1ull);
    j = __shfl_sync(kFull, j, 0);
    if (j >= cap) continue;
    const float* src = reinterpret_cast<const float*>(addrs[i]);
    float* dst = shard_rows + j * dim;
    if ((dim & 3u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (uint32_t v = lane; v < dim / 4u; v += 32u) d4[v] = s4[v];
    } else {
      for (uint32_t v = lane; v < dim; v += 32u) dst[v] = src[v];
    }
    if (lane == 0) shard_keys[j] = key;
  }
}

// Points the direct-pull index at a shard: entry i of the shard (key shard_keys[i], possibly read over NVLink from
// the owner's HBM) lives at shard_rows + i * dim.  Keys that are not in this rank's index are left alone (the index
// is the host table's: a key the host table does not have keeps answering with the default vector).
__global__ void __launch_bounds__(kBlock) index_repoint_kernel(IndexSlot* slots, uint64_t mask,
                                                               const int64_t* __restrict__ shard_keys,
                                                               const float* shard_rows, unsigned long long n,
                                                               uint32_t dim, unsigned long long* repointed) {
  const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= n) return;
  const int64_t key = shard_keys[i];
  if (key == kEmptyKey) return;
  uint64_t slot = mix64(static_cast<uint64_t>(key)) & mask;
  while (true) {
    const int64_t k = __ldcg(reinterpret_cast<const long long*>(&slots[slot].key));
    if (k == key) {
      slots[slot].row = shard_rows + i * dim;
      if (repointed != nullptr) atomicAdd(repointed, 1ull);
      return;
    }
    if (k == kEmptyKey) return;
    slot = (slot + 1) & mask;
  }
}

// Warp-cooperative index lookup: 8 lanes read 8 consecutive slots (one 128-B line), the first match
// before the first empty slot wins.  Returns the host row address or nullptr (key not in the table).
__device__ __forceinline__ const float* index_find(const IndexSlot* __restrict__ index, uint64_t mask,
                                                   int64_t key, uint32_t lane) {
  uint64_t base = mix64(static_cast<uint64_t>(key)) & mask;
  while (true) {
    int64_t k = kEmptyKey;
    unsigned long long row = 0;
    if (lane < 8) {
      const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(&index[(base + lane) & mask]));
      k = static_cast<int64_t>((static_cast<unsigned long long>(raw.y) << 32) | raw.x);
      row = (static_cast<unsigned long long>(raw.w) << 32) | raw.z;
    }
    const unsigned hit = __ballot_sync(kFull, lane < 8 && k == key);
    const unsigned empty = __ballot_sync(kFull, lane < 8 && k == kEmptyKey);
    if (hit != 0u && (empty == 0u || __ffs(hit) < __ffs(empty)))
      return reinterpret_cast<const float*>(__shfl_sync(kFull, row, __ffs(hit) - 1));
    if (empty != 0u) return nullptr;
    base = (base + 8) & mask;
  }
}

struct PullArgs {
  Bucket* buckets;
  float* values;
  uint32_t num_buckets;
  uint32_t dim;
  float default_value;
  const IndexSlot* index;
  uint64_t index_mask;
  const float* sentinel_row;
  const int64_t* miss_keys;
  const uint32_t* miss_pos;
  const uint32_t* miss_count;
  uint32_t n_keys;
  float* out;
  float* stage;
  int insert;
  int insert_mode;
  float hit_rate_threshold;
  uint32_t epoch;
  uint32_t* inserted;
  uint32_t* absent;
  int load_mode;
  // optional: the misses in ascending host-address order (resolve_rows_kernel + radix sort); then entry i
  // is miss sorted_idx[i] and its row lives at sorted_addr[i] (0: key absent from the host table)
  const unsigned long long* sorted_addr;
  const uint32_t* sorted_idx;
  __nv_bfloat16* out_bf16;  // optional bf16 mirror of `out` (float4 instantiation only)
  int64_t* mark_absent;     // optional (= miss_keys): keys that are not in the host table are overwritten with kEmptyKey,
                            // so that a later insert-from-output pass skips them
  // merged miss list of a batch of requests: miss_pos = (request << kShardPosBits) | row, one output base per request
  int batch;
  float* batch_out[kMaxBatchOuts];
};

// Step 1 of the sorted pull: host address of every missed key (8 lanes per key, one 128-B index line each)
// plus the identity permutation the radix sort carries along.
__global__ void __launch_bounds__(kBlock) resolve_rows_kernel(const IndexSlot* __restrict__ index, uint64_t mask,
                                                              const float* sentinel_row,
                                                              const int64_t* __restrict__ miss_keys, uint32_t m,
                                                              unsigned long long* addr, uint32_t* idx) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t sub = lane & 7u;           // lane within the 8-lane group
  const uint32_t group_shift = lane & ~7u;  // first lane of the group
  const unsigned group_mask = 0xffu << group_shift;
  const size_t g = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 3;
  if (g >= m) return;  // whole 8-lane groups retire together
  const int64_t key = miss_keys[g];
  unsigned long long row = 0;
  if (key == kEmptyKey) {
    row = reinterpret_cast<unsigned long long>(sentinel_row);
  } else {
    uint64_t base = mix64(static_cast<uint64_t>(key)) & mask;
    while (true) {
      const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(&index[(base + sub) & mask]));
      const int64_t k = static_cast<int64_t>((static_cast<unsigned long long>(raw.y) << 32) | raw.x);
      const unsigned long long r = (static_cast<unsigned long long>(raw.w) << 32) | raw.z;
      const unsigned hit = (__ballot_sync(group_mask, k == key) & group_mask) >> group_shift;
      const unsigned empty = (__ballot_sync(group_mask, k == kEmptyKey) & group_mask) >> group_shift;
      if (hit != 0u && (empty == 0u || __ffs(hit) < __ffs(empty))) {
        row = __shfl_sync(group_mask, r, group_shift + __ffs(hit) - 1);
        break;
      }
      if (empty != 0u) break;
      base = (base + 8) & mask;
    }
  }
  if (sub == 0) {
    addr[g] = row;
    idx[g] = static_cast<uint32_t>(g);
  }
}

// Host-row loads.  mode 0: default-cached ld.global (full 128-B line requests to the host link),
// 1: ld.global.nc.L1::no_allocate, 2: ld.global.cg.
template <typename VecT>
__device__ __forceinline__ VecT ld_host(const VecT* p, int mode) {
  if (mode == 1) return ld_stream(p);
  if (mode == 2) return __ldcg(p);
  return *p;
}

template <typename VecT, int kRows>
__global__ void __launch_bounds__(kBlock) pull_misses_kernel(const PullArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t nwarps = (static_cast<size_t>(gridDim.x) * kBlock) >> 5;
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const uint32_t m = *a.miss_count;
  bool sync_mode = a.insert_mode != 0;
  if (a.insert_mode < 0) {
    // [UPSTREAM] hit_rate < hit_rate_threshold -> synchronous insertion; decided here, on the device
    const double hit_rate = 1.0 - static_cast<double>(m) / static_cast<double>(a.n_keys);
    sync_mode = hit_rate < static_cast<double>(a.hit_rate_threshold);
  }
  const bool write_out = (a.out != nullptr || a.batch != 0) && sync_mode;
  const VecT defv = splat<VecT>(a.default_value);
  uint32_t n_inserted = 0;
  for (size_t i0 = warp * kRows; i0 < m; i0 += nwarps * kRows) {
    int64_t key[kRows];
    const VecT* src[kRows];
    VecT x0[kRows], x1[kRows];
    // index probes of all rows first, then every PCIe read of the group: kRows x 512 B in flight per warp.
    // The PCIe reads are the long pole (~2-3 us); the bucket claims below overlap them.
    size_t idx[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const size_t i = i0 + r;
      idx[r] = i;
      key[r] = kEmptyKey;
      const float* row = nullptr;
      if (i < m) {
        if (a.sorted_addr != nullptr) {
          idx[r] = a.sorted_idx[i];
          key[r] = a.miss_keys[idx[r]];
          row = reinterpret_cast<const float*>(a.sorted_addr[i]);
        } else {
          key[r] = a.miss_keys[i];
          row = key[r] == kEmptyKey ? a.sentinel_row : index_find(a.index, a.index_mask, key[r], lane);
        }
      }
      src[r] = reinterpret_cast<const VecT*>(row);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      x0[r] = defv;
      x1[r] = defv;
      if (src[r] != nullptr) {
        if (lane < V) x0[r] = ld_host(src[r] + lane, a.load_mode);
        if (lane + 32u < V) x1[r] = ld_host(src[r] + lane + 32u, a.load_mode);
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (i0 + r >= m) break;
      const size_t i = idx[r];  // position in the miss list (stage row, miss_pos entry)
      VecT* dst_out = nullptr;
      if (write_out) {
        const uint32_t p = a.miss_pos[i];
        dst_out = a.batch ? reinterpret_cast<VecT*>(a.batch_out[p >> kShardPosBits]) +
                                static_cast<size_t>(p & ((1u << kShardPosBits) - 1u)) * V
                          : reinterpret_cast<VecT*>(a.out) + static_cast<size_t>(p) * V;
      }
      VecT* dst_stage = a.stage ? reinterpret_cast<VecT*>(a.stage) + i * V : nullptr;
      VecT* dst_slab = nullptr;
      if (a.insert && src[r] != nullptr && key[r] != kEmptyKey) {
        const uint32_t slot = claim_slot(a.buckets, a.num_buckets, key[r], a.epoch, lane);
        if (slot != kMissSlot) dst_slab = reinterpret_cast<VecT*>(a.values) + static_cast<size_t>(slot) * V;
      }
      for (uint32_t v = lane; v < V; v += 32u) {
        VecT x;
        if (v == lane)
          x = x0[r];
        else if (v == lane + 32u)
          x = x1[r];
        else
          x = src[r] != nullptr ? ld_host(src[r] + v, a.load_mode) : defv;
        if (dst_out) st_stream(dst_out + v, x);
        if (dst_stage) dst_stage[v] = x;
        if (dst_slab) dst_slab[v] = x;
        if constexpr (sizeof(VecT) == 16) {
          if (dst_out && a.out_bf16) st_bf16x4(a.out_bf16 + (static_cast<size_t>(a.miss_pos[i]) * V + v) * 4u, x);
        }
      }
      n_inserted += dst_slab != nullptr ? 1u : 0u;
      if (lane == 0 && src[r] == nullptr) {
        if (a.absent != nullptr) atomicAdd(a.absent, 1u);
        if (a.mark_absent != nullptr) a.mark_absent[i] = kEmptyKey;
      }
    }
  }
  if (lane == 0 && n_inserted != 0u && a.inserted != nullptr) atomicAdd(a.inserted, n_inserted);
}

// ------------------------------------------------------------------------------------------------
// Binned miss path: the probe kernels appended every miss to the list of its host-table partition;
// the kernels below walk those lists bin by bin.  A CTA first builds the exclusive prefix of the (clamped) bin
// counts in shared memory, then every warp takes flat entry indices warp, warp + nwarps, ... and maps them back to
// (bin, j) by binary search — so the whole grid works on one or two neighbouring bins at any moment, which is what
// keeps the PCIe reads in flight inside a <= 256-MiB window of host memory.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMaxBins = 4096;  // HostTable caps its partition count at this

// prefix[b] = entries in bins [0, b), b in [0, nb]; nb = num_bins + 1 (the spill list is the last bin)
__device__ __forceinline__ uint32_t build_bin_prefix(const MissBins& bins, uint32_t* prefix, uint32_t* warp_sums) {
  const uint32_t nb = bins.num_bins + 1;
  const uint32_t per = (nb + kBlock - 1) / kBlock;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t b0 = threadIdx.x * per;
  uint32_t local = 0;
  for (uint32_t b = b0; b < min(nb, b0 + per); ++b) {
    const uint32_t c = __ldcg(bins.count + b);
    local += b < bins.num_bins ? min(c, bins.bin_cap) : min(c, bins.spill_cap);
  }
  uint32_t incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(kFull, incl, off);
    if (lane >= static_cast<uint32_t>(off)) incl += o;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += warp_sums[w];
  uint32_t run = base + incl - local;
  for (uint32_t b = b0; b < min(nb, b0 + per); ++b) {
    prefix[b] = run;
    const uint32_t c = __ldcg(bins.count + b);
    run += b < bins.num_bins ? min(c, bins.bin_cap) : min(c, bins.spill_cap);
  }
  if (threadIdx.x == kBlock - 1) prefix[nb] = base + incl;  // total (threads past nb contribute 0)
  __syncthreads();
  return prefix[nb];
}

// flat index -> slot of the entry in bins.keys / bins.pos
__device__ __forceinline__ size_t bin_entry(const MissBins& bins, const uint32_t* prefix, uint32_t i) {
  uint32_t lo = 0, hi = bins.num_bins + 1;  // find the last b with prefix[b] <= i
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (prefix[mid] <= i) lo = mid; else hi = mid;
  }
  return static_cast<size_t>(lo) * bins.bin_cap + (i - prefix[lo]);
}

struct PullBinnedArgs {
  MissBins bins;
  const IndexSlot* index;
  uint64_t index_mask;
  const float* sentinel_row;
  uint32_t dim;
  float default_value;
  float* out;
  __nv_bfloat16* out_bf16;
  uint32_t* absent;
  int batch;
  float* batch_out[kMaxBatchOuts];
  // fused insert (kInsert): the warp that pulled a row also puts it into the cache — HBM-only work hidden behind the
  // PCIe reads of the other warps.  Needs the cache to itself (no probe may run beside it).
  Bucket* buckets;
  float* values;
  uint32_t num_buckets;
  uint32_t epoch;
  uint32_t* inserted;
};

// Group form of index_find: the warp resolves FOUR keys at once, 8 lanes (one 128-B index line) per key.  Lanes
// 8g..8g+7 work on key[g]; returns, in every lane, the row address of the key of ITS group (nullptr: not in the table).
__device__ __forceinline__ const float* index_find4(const IndexSlot* __restrict__ index, uint64_t mask, int64_t key,
                                                    bool active, uint32_t lane) {
  const uint32_t sub = lane & 7u, shift = lane & ~7u;
  uint64_t base = mix64(static_cast<uint64_t>(key)) & mask;
  const float* found = nullptr;
  bool done = !active;
  while (__any_sync(kFull, !done)) {
    int64_t k = kEmptyKey;
    unsigned long long row = 0;
    if (!done) {
      const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(&index[(base + sub) & mask]));
      k = static_cast<int64_t>((static_cast<unsigned long long>(raw.y) << 32) | raw.x);
      row = (static_cast<unsigned long long>(raw.w) << 32) | raw.z;
    }
    const unsigned hit = (__ballot_sync(kFull, !done && k == key) >> shift) & 0xffu;
    const unsigned empty = (__ballot_sync(kFull, !done && k == kEmptyKey) >> shift) & 0xffu;
    const bool is_hit = hit != 0u && (empty == 0u || __ffs(hit) < __ffs(empty));
    const unsigned long long r = __shfl_sync(kFull, row, shift + (is_hit ? __ffs(hit) - 1 : 0));
    if (!done) {
      if (is_hit) {
        found = reinterpret_cast<const float*>(r);
        done = true;
      } else if (empty != 0u) {
        done = true;
      } else {
        base = (base + 8) & mask;
      }
    }
  }
  return found;
}

// kRows = 1: one row per warp at a time — the host link needs few bytes in flight, and the closer together they lie
// in host memory the faster it runs (engine.hpp pull_grid_ctas).  kRows = 4 (rows of <= 32 vectors; NVLink tier):
// four entries per warp iteration — their index lines are read together (8 lanes each), then all four rows are
// in flight before the first is stored: NVLink / HBM want megabytes in flight, not locality.
template <typename VecT, bool kInsert, int kRows>
__global__ void __launch_bounds__(kBlock) pull_binned_kernel(const PullBinnedArgs a) {
  __shared__ uint32_t prefix[kMaxBins + 2];
  __shared__ uint32_t warp_sums[kBlock / 32];
  const uint32_t total = build_bin_prefix(a.bins, prefix, warp_sums);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * kBlock) >> 5;
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const VecT defv = splat<VecT>(a.default_value);
  uint32_t n_inserted = 0;  // per warp; one atomic at the end (see insert_merge_kernel)
  if constexpr (kRows == 4) {
    const uint32_t g = lane >> 3;  // the entry of the quad whose index line this lane reads
    for (uint32_t i0 = warp * 4u; i0 < total; i0 += nwarps * 4u) {
      // lane 8e reads entry e's key and position; the other lanes get them by shuffle
      int64_t my_key = kEmptyKey;
      uint32_t my_pos = 0;
      size_t my_r = 0;
      const bool have = i0 + g < total;
      if (have) {
        my_r = bin_entry(a.bins, prefix, i0 + g);
        my_key = a.bins.keys[my_r];
        my_pos = a.bins.pos[my_r];
      }
      const float* my_row = index_find4(a.index, a.index_mask, my_key, have && my_key != kEmptyKey, lane);
      if (have && my_key == kEmptyKey) my_row = a.sentinel_row;
      int64_t key[4];
      uint32_t pos[4];
      const VecT* src[4];
      VecT x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        key[e] = __shfl_sync(kFull, my_key, 8 * e);
        pos[e] = __shfl_sync(kFull, my_pos, 8 * e);
        src[e] = reinterpret_cast<const VecT*>(
            __shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), 8 * e));
      }
      // The four slot claims together, BEFORE the rows are requested: releasing a bucket lock is a MEMBAR that waits
      // for every memory operation the thread has in flight — with the row loads outstanding it would wait for NVLink.
      uint32_t slot4[4] = {kMissSlot, kMissSlot, kMissSlot, kMissSlot};
      if constexpr (kInsert) {
        unsigned deferred = 0;
        const uint32_t my_slot = claim_slot4(a.buckets, a.num_buckets, my_key, have && my_row != nullptr && my_key != kEmptyKey,
                                             a.epoch, lane, &deferred);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          slot4[e] = __shfl_sync(kFull, my_slot, 8 * e);
          if ((deferred >> e) & 1u) slot4[e] = claim_slot(a.buckets, a.num_buckets, key[e], a.epoch, lane);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[e] = defv;
        if (src[e] != nullptr && lane < V) x[e] = ld_stream(src[e] + lane);  // rows are read-only while lookups run
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (i0 + e >= total) break;
        VecT* dst = a.batch ? reinterpret_cast<VecT*>(a.batch_out[pos[e] >> kShardPosBits]) +
                                  static_cast<size_t>(pos[e] & ((1u << kShardPosBits) - 1u)) * V
                            : reinterpret_cast<VecT*>(a.out) + static_cast<size_t>(pos[e]) * V;
        VecT* slab = nullptr;
        if constexpr (kInsert) {
          if (slot4[e] != kMissSlot) slab = reinterpret_cast<VecT*>(a.values) + static_cast<size_t>(slot4[e]) * V;
        }
        if (lane < V) {
          st_stream(dst + lane, x[e]);
          if (kInsert && slab != nullptr) slab[lane] = x[e];
          if constexpr (sizeof(VecT) == 16) {
            if (a.out_bf16) st_bf16x4(a.out_bf16 + (static_cast<size_t>(pos[e]) * V + lane) * 4u, x[e]);
          }
        }
        n_inserted += (kInsert && slab != nullptr) ? 1u : 0u;
        if (lane == 0) {
          if (src[e] == nullptr) {
            atomicAdd(a.absent, 1u);
            a.bins.keys[bin_entry(a.bins, prefix, i0 + e)] = kEmptyKey;  // the insert pass skips it
          }
        }
      }
    }
  } else {
  for (uint32_t i = warp; i < total; i += nwarps) {
    const size_t r = bin_entry(a.bins, prefix, i);
    const int64_t key = a.bins.keys[r];
    const uint32_t p = a.bins.pos[r];
    const float* row = key == kEmptyKey ? a.sentinel_row : index_find(a.index, a.index_mask, key, lane);
    const VecT* src = reinterpret_cast<const VecT*>(row);
    VecT* dst = a.batch ? reinterpret_cast<VecT*>(a.batch_out[p >> kShardPosBits]) +
                              static_cast<size_t>(p & ((1u << kShardPosBits) - 1u)) * V
                        : reinterpret_cast<VecT*>(a.out) + static_cast<size_t>(p) * V;
    // all reads of the row first (up to 4 per lane in flight); the slot claim runs while they are in flight
    VecT* slab = nullptr;
    for (uint32_t v0 = 0; v0 < V; v0 += 128u) {
      VecT x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t v = v0 + u * 32u + lane;
        x[u] = defv;
        if (src != nullptr && v < V) x[u] = src[v];
      }
      if constexpr (kInsert) {
        if (v0 == 0 && src != nullptr && key != kEmptyKey) {
          const uint32_t slot = claim_slot(a.buckets, a.num_buckets, key, a.epoch, lane);
          if (slot != kMissSlot) slab = reinterpret_cast<VecT*>(a.values) + static_cast<size_t>(slot) * V;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t v = v0 + u * 32u + lane;
        if (v < V) {
          st_stream(dst + v, x[u]);
          if (kInsert && slab != nullptr) slab[v] = x[u];
          if constexpr (sizeof(VecT) == 16) {
            if (a.out_bf16) st_bf16x4(a.out_bf16 + (static_cast<size_t>(p) * V + v) * 4u, x[u]);
          }
        }
      }
    }
    n_inserted += (kInsert && slab != nullptr) ? 1u : 0u;
    if (lane == 0) {
      if (src == nullptr) {
        atomicAdd(a.absent, 1u);
        a.bins.keys[r] = kEmptyKey;  // the insert pass skips it
      }
    }
  }
  }
  if (kInsert && lane == 0 && n_inserted != 0u && a.inserted != nullptr) atomicAdd(a.inserted, n_inserted);
}

// Tables that live ONLY in the NVLink tier (model-parallel rows, no host copy and no local cache: BASELINE configs[3]):
// out[i] = row(keys[i]) read from the owner's shard through the index, default vector for keys in no shard.  A warp
// resolves four keys per iteration (one 128-B index line per key, 8 lanes each) and has their four rows in flight
// before it stores the first.  Bound by NVLink ingress ((world-1)/world of the rows) — no bins, no flags, no peer SMs.
template <typename VecT>
__global__ void __launch_bounds__(kBlock) tier_gather_kernel(const int64_t* __restrict__ keys, uint32_t n,
                                                             const IndexSlot* __restrict__ index, uint64_t mask,
                                                             uint32_t dim, float default_value, float* out,
                                                             uint32_t* absent) {
  // A warp owns a tile of 32 consecutive keys (one coalesced 256-B load — also when `keys` is the caller's pinned host
  // buffer, read in place over PCIe) and serves it quad by quad.
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * kBlock) >> 5;
  const uint32_t V = dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const VecT defv = splat<VecT>(default_value);
  const uint32_t g = lane >> 3;
  uint32_t missing = 0;
  for (uint32_t t0 = warp * 32u; t0 < n; t0 += nwarps * 32u) {
    const int64_t tile_key = t0 + lane < n ? keys[t0 + lane] : kEmptyKey;
#pragma unroll 2
    for (uint32_t q = 0; q < 8u; ++q) {
      const uint32_t i0 = t0 + 4u * q;
      if (i0 >= n) break;
      const bool have = i0 + g < n;
      const int64_t my_key = __shfl_sync(kFull, tile_key, 4u * q + g);
      const float* my_row = index_find4(index, mask, my_key, have && my_key != kEmptyKey, lane);
      const VecT* src[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        src[e] = reinterpret_cast<const VecT*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_row), 8 * e));
      for (uint32_t v0 = 0; v0 < V; v0 += 32u) {
        const uint32_t v = v0 + lane;
        VecT x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          x[e] = defv;
          if (src[e] != nullptr && v < V) x[e] = ld_stream(src[e] + v);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (i0 + e < n && v < V) st_stream(reinterpret_cast<VecT*>(out) + static_cast<size_t>(i0 + e) * V + v, x[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) missing += (i0 + e < n && src[e] == nullptr) ? 1u : 0u;
    }
  }
  if (lane == 0 && missing != 0u && absent != nullptr) atomicAdd(absent, missing);
}

// Synthetic shard (tables too large for a host copy): of the keys [key_lo, key_hi) append those this rank owns, with
// their synth_value() rows generated in place.  A CTA scans 1024 keys, compacts the owned ones in shared memory,
// reserves their shard positions with one atomic and lets its warps write the rows.
constexpr int kFillKeysPerCta = 4 * kBlock;
__global__ void __launch_bounds__(kBlock) tier_fill_procedural_kernel(unsigned long long key_lo, unsigned long long key_hi,
                                                                      unsigned long long seed, uint32_t rank, uint32_t world,
                                                                      uint32_t dim, int64_t* shard_keys, float* shard_rows,
                                                                      unsigned long long cap, unsigned long long* count) {
  __shared__ int64_t owned[kFillKeysPerCta];
  __shared__ uint32_t n_owned;
  __shared__ unsigned long long base;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long chunks = (key_hi - key_lo + kFillKeysPerCta - 1) / kFillKeysPerCta;
  for (unsigned long long chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    if (threadIdx.x == 0) n_owned = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long key = key_lo + chunk * kFillKeysPerCta + j * kBlock + threadIdx.x;
      if (key < key_hi && owner_of(static_cast<int64_t>(key), world) == rank)
        owned[atomicAdd(&n_owned, 1u)] = static_cast<int64_t>(key);
    }
    __syncthreads();
    if (threadIdx.x == 0) base = atomicAdd(count, static_cast<unsigned long long>(n_owned));
    __syncthreads();
    for (uint32_t i = warp; i < n_owned; i += kBlock / 32) {
      const unsigned long long j = base + i;
      if (j >= cap) continue;
      const int64_t key = owned[i];
      if (lane == 0) shard_keys[j] = key;
      for (uint32_t v = lane; v < dim; v += 32u) shard_rows[j * dim + v] = synth_value(key, v, seed);
    }
    __syncthreads();
  }
}

// index[key] = shard_rows + i * dim for every entry of a shard (insert or overwrite; shard memory may be a peer's).
__global__ void __launch_bounds__(kBlock) index_insert_shard_kernel(IndexSlot* slots, uint64_t mask,
                                                                    const int64_t* __restrict__ shard_keys,
                                                                    const float* shard_rows, unsigned long long n,
                                                                    uint32_t dim) {
  const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= n) return;
  const int64_t key = shard_keys[i];
  if (key == kEmptyKey) return;
  uint64_t slot = mix64(static_cast<uint64_t>(key)) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&slots[slot].key),
                                              static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
    if (prev == static_cast<unsigned long long>(kEmptyKey) || prev == static_cast<unsigned long long>(key)) {
      slots[slot].row = shard_rows + i * dim;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

struct InsertBinnedArgs {
  MissBins bins;
  Bucket* buckets;
  float* values;
  uint32_t num_buckets;
  uint32_t dim;
  const float* out;
  uint32_t epoch;
  uint32_t* inserted;
  int batch;
  float* batch_out[kMaxBatchOuts];
};

template <typename VecT>
__global__ void __launch_bounds__(kBlock) insert_binned_kernel(const InsertBinnedArgs a) {
  __shared__ uint32_t prefix[kMaxBins + 2];
  __shared__ uint32_t warp_sums[kBlock / 32];
  const uint32_t total = build_bin_prefix(a.bins, prefix, warp_sums);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * kBlock) >> 5;
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  uint32_t n_inserted = 0;
  for (uint32_t i = warp; i < total; i += nwarps) {
    const size_t r = bin_entry(a.bins, prefix, i);
    const int64_t key = a.bins.keys[r];
    if (key == kEmptyKey) continue;  // not in the host table (or the key that is never cached)
    const uint32_t p = a.bins.pos[r];
    const VecT* src = a.batch ? reinterpret_cast<const VecT*>(a.batch_out[p >> kShardPosBits]) +
                                    static_cast<size_t>(p & ((1u << kShardPosBits) - 1u)) * V
                              : reinterpret_cast<const VecT*>(a.out) + static_cast<size_t>(p) * V;
    const uint32_t slot = claim_slot(a.buckets, a.num_buckets, key, a.epoch, lane);
    if (slot != kMissSlot) {
      VecT* dst = reinterpret_cast<VecT*>(a.values) + static_cast<size_t>(slot) * V;
      for (uint32_t v = lane; v < V; v += 32u) dst[v] = ld_stream(src + v);
    }
    n_inserted += slot != kMissSlot ? 1u : 0u;
  }
  if (lane == 0 && n_inserted != 0u && a.inserted != nullptr) atomicAdd(a.inserted, n_inserted);
}

// ------------------------------------------------------------------------------------------------
// a8 pooled gather + reduce.  A group of `kGroup` lanes owns one bag; fp32 adds in ascending j.
// ------------------------------------------------------------------------------------------------
template <typename VecT, int kGroup>
__global__ void __launch_bounds__(kBlock) pooled_gather_kernel(
    const float* __restrict__ values, const float* __restrict__ stage,
    const uint32_t* __restrict__ src, size_t num_bags, uint32_t hotness, uint32_t dim, int mean,
    float* __restrict__ out) {
  constexpr uint32_t kBagsPerWarp = 32 / kGroup;
  const uint32_t lane = threadIdx.x & 31u;
  const size_t warp = (static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5;
  const size_t bag = warp * kBagsPerWarp + lane / kGroup;
  const uint32_t v0 = lane % kGroup;
  if (bag >= num_bags) return;
  const uint32_t V = dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const VecT* vals = reinterpret_cast<const VecT*>(values);
  const VecT* stg = reinterpret_cast<const VecT*>(stage);
  const uint32_t* s = src + bag * hotness;
  for (uint32_t v = v0; v < V; v += kGroup) {
    VecT acc = splat<VecT>(0.f);
    uint32_t j = 0;
    for (; j + 4 <= hotness; j += 4) {
      VecT r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t sj = s[j + u];
        const VecT* row = (sj & kSrcMissBit)
                              ? stg + static_cast<size_t>(sj & ~kSrcMissBit) * V
                              : vals + static_cast<size_t>(sj) * V;
        r[u] = ld_stream(row + v);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) vadd(acc, r[u]);
    }
    for (; j < hotness; ++j) {
      const uint32_t sj = s[j];
      const VecT* row = (sj & kSrcMissBit) ? stg + static_cast<size_t>(sj & ~kSrcMissBit) * V
                                           : vals + static_cast<size_t>(sj) * V;
      vadd(acc, ld_stream(row + v));
    }
    if (mean) vdiv(acc, static_cast<float>(hotness));
    st_stream(reinterpret_cast<VecT*>(out) + bag * V + v, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// table maintenance
// ------------------------------------------------------------------------------------------------
__global__ void table_clear_kernel(Bucket* buckets, uint32_t num_buckets) {
  // 8 threads per bucket, 16 B each
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t b = i >> 3;
  if (b >= num_buckets) return;
  const uint32_t part = i & 7u;
  uint4* p = reinterpret_cast<uint4*>(&buckets[b]) + part;
  if (part < 4) {
    const unsigned long long e = static_cast<unsigned long long>(kEmptyKey);
    *p = make_uint4(static_cast<uint32_t>(e), static_cast<uint32_t>(e >> 32),
                    static_cast<uint32_t>(e), static_cast<uint32_t>(e >> 32));
  } else {
    *p = make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void count_resident_kernel(const Bucket* buckets, uint32_t num_buckets,
                                      unsigned long long* count) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool used = false;
  if ((i >> 3) < num_buckets) used = buckets[i >> 3].keys[i & 7u] != kEmptyKey;
  const unsigned m = __ballot_sync(kFull, used);
  if ((threadIdx.x & 31u) == 0 && m != 0u) atomicAdd(count, static_cast<unsigned long long>(__popc(m)));
}

__global__ void dump_keys_kernel(const Bucket* buckets, uint32_t num_buckets, int64_t* out,
                                 unsigned long long cap, unsigned long long* count) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((i >> 3) >= num_buckets) return;
  const int64_t k = buckets[i >> 3].keys[i & 7u];
  if (k == kEmptyKey) return;
  const unsigned long long r = atomicAdd(count, 1ull);
  if (r < cap) out[r] = k;
}

// ------------------------------------------------------------------------------------------------
// K1 dedup: open-addressing scratch table, CAS claim, warp-aggregated id allocation.
// counter[0] = #unique, counter[1] = sentinel-key flag, counter[2] = sentinel-key id.
// ------------------------------------------------------------------------------------------------
__global__ void fill_empty_kernel(int64_t* ws_keys, size_t cap) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < cap) ws_keys[i] = kEmptyKey;
}

__global__ void __launch_bounds__(kBlock) unique_claim_kernel(const int64_t* __restrict__ keys,
                                                              size_t n, int64_t* ws_keys,
                                                              uint32_t* ws_ids, size_t mask,
                                                              int64_t* unique, uint32_t* counter) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool won = false;
  bool sentinel = false;
  size_t slot = 0;
  int64_t key = 0;
  if (i < n) {
    key = keys[i];
    if (key == kEmptyKey) {
      sentinel = true;
      won = atomicCAS(&counter[1], 0u, 1u) == 0u;
    } else {
      slot = mix64(static_cast<uint64_t>(key)) & mask;
      while (true) {
        const unsigned long long prev =
            atomicCAS(reinterpret_cast<unsigned long long*>(&ws_keys[slot]),
                      static_cast<unsigned long long>(kEmptyKey), static_cast<unsigned long long>(key));
        if (prev == static_cast<unsigned long long>(kEmptyKey)) {
          won = true;
          break;
        }
        if (prev == static_cast<unsigned long long>(key)) break;
        slot = (slot + 1) & mask;
      }
    }
  }
  const unsigned m = __ballot_sync(kFull, won);
  if (m == 0u) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&counter[0], static_cast<uint32_t>(__popc(m)));
  base = __shfl_sync(kFull, base, 0);
  if (won) {
    const uint32_t id = base + __popc(m & ((1u << lane) - 1u));
    unique[id] = key;
    if (sentinel)
      counter[2] = id;
    else
      ws_ids[slot] = id;
  }
}

__global__ void __launch_bounds__(kBlock) unique_resolve_kernel(const int64_t* __restrict__ keys,
                                                                size_t n,
                                                                const int64_t* __restrict__ ws_keys,
                                                                const uint32_t* __restrict__ ws_ids,
                                                                size_t mask, uint32_t* inverse,
                                                                const uint32_t* counter) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  if (i >= n) return;
  const int64_t key = keys[i];
  if (key == kEmptyKey) {
    inverse[i] = counter[2];
    return;
  }
  size_t slot = mix64(static_cast<uint64_t>(key)) & mask;
  while (ws_keys[slot] != key) slot = (slot + 1) & mask;
  inverse[i] = ws_ids[slot];
}

// ------------------------------------------------------------------------------------------------
// multi-GPU key routing: histogram by owner, then scatter into per-owner contiguous ranges.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxShards = 64;

__global__ void __launch_bounds__(kBlock) route_hist_kernel(const int64_t* __restrict__ keys,
                                                            size_t n, uint32_t shards,
                                                            uint32_t* counts) {
  __shared__ uint32_t h[kMaxShards];
  if (threadIdx.x < kMaxShards) h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = static_cast<size_t>(blockIdx.x) * kRouteChunk;
  for (uint32_t j = threadIdx.x; j < kRouteChunk; j += kBlock) {
    const size_t i = base + j;
    if (i < n) atomicAdd(&h[owner_of(keys[i], shards)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < shards && h[threadIdx.x] != 0u) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(kBlock) route_scatter_kernel(const int64_t* __restrict__ keys,
                                                               size_t n, uint32_t shards,
                                                               const uint32_t* __restrict__ counts,
                                                               uint32_t* cursor, int64_t* routed,
                                                               uint32_t* perm) {
  __shared__ uint32_t h[kMaxShards];
  __shared__ uint32_t start[kMaxShards];
  if (threadIdx.x < kMaxShards) h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = static_cast<size_t>(blockIdx.x) * kRouteChunk;
  for (uint32_t j = threadIdx.x; j < kRouteChunk; j += kBlock) {
    const size_t i = base + j;
    if (i < n) atomicAdd(&h[owner_of(keys[i], shards)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < shards) {
    uint32_t prefix = 0;
    for (uint32_t g = 0; g < threadIdx.x; ++g) prefix += counts[g];
    const uint32_t mine = h[threadIdx.x];
    start[threadIdx.x] = prefix + (mine ? atomicAdd(&cursor[threadIdx.x], mine) : 0u);
  }
  __syncthreads();
  if (threadIdx.x < kMaxShards) h[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < kRouteChunk; j += kBlock) {
    const size_t i = base + j;
    if (i < n) {
      const int64_t k = keys[i];
      const uint32_t o = owner_of(k, shards);
      const uint32_t p = start[o] + atomicAdd(&h[o], 1u);
      routed[p] = k;
      perm[p] = static_cast<uint32_t>(i);
    }
  }
}

template <typename VecT>
__global__ void __launch_bounds__(kBlock) scatter_rows_kernel(const float* __restrict__ rows,
                                                              const uint32_t* __restrict__ perm,
                                                              size_t n, uint32_t V, float* out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  const size_t row = i / V;
  if (row >= n) return;
  const uint32_t v = static_cast<uint32_t>(i - row * V);
  const VecT x = ld_stream(reinterpret_cast<const VecT*>(rows) + i);
  st_stream(reinterpret_cast<VecT*>(out) + static_cast<size_t>(perm[row]) * V + v, x);
}

// Plain random row gather out[i] = table[idx[i]] with the same tile/unroll shape as the probe+gather
// kernel but no hashing and no bucket reads: the measured ceiling of "HBM random-gather" on this part.
template <int kV, int kUnroll>
__global__ void __launch_bounds__(kBlock) gather_rows_kernel(const float4* __restrict__ table,
                                                             const uint32_t* __restrict__ idx, size_t n,
                                                             float4* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const size_t tile_base = ((static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x) >> 5) * 32;
  if (tile_base >= n) return;
  const uint32_t nk = static_cast<uint32_t>(min(static_cast<size_t>(32), n - tile_base));
  const uint32_t slot = lane < nk ? idx[tile_base + lane] : 0u;
  constexpr uint32_t V = kV;
  float4* __restrict__ outv = out + tile_base * V;
  const uint32_t total = nk * V;
  for (uint32_t i0 = 0; i0 < total; i0 += 32u * kUnroll) {
    float4 buf[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      const uint32_t kk = min(i / V, 31u);
      const uint32_t s = __shfl_sync(kFull, slot, kk);
      if (i < total) buf[u] = ld_stream(table + static_cast<size_t>(s) * V + (i - kk * V));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      if (i < total) st_stream(outv + i, buf[u]);
    }
  }
}

__global__ void __launch_bounds__(kBlock) synth_rows_kernel(const int64_t* __restrict__ keys, size_t n,
                                                            uint32_t dim, uint64_t seed, float* rows) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  const size_t row = i / dim;
  if (row >= n) return;
  const uint32_t j = static_cast<uint32_t>(i - row * dim);
  rows[i] = synth_value(keys[row], j, seed);
}

// widest vector (bytes) usable for rows of `dim` floats given the pointer alignments involved
inline int vec_bytes(uint32_t dim, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr,
                     const void* p3 = nullptr) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) |
                      reinterpret_cast<uintptr_t>(p2) | reinterpret_cast<uintptr_t>(p3) |
                      (static_cast<uintptr_t>(dim) * 4u);
  if ((a & 15u) == 0) return 16;
  if ((a & 7u) == 0) return 8;
  return 4;
}


template <typename VecT>
cudaError_t launch_probe_ldg_scatter(const ProbeArgs& a, cudaStream_t stream) {
  const unsigned grid = grid_for(a.n);
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  if (V == 32)
    probe_gather_ldg_kernel<VecT, 32, 8, true><<<grid, kBlock, 0, stream>>>(a);
  else
    probe_gather_ldg_kernel<VecT, 0, 4, true><<<grid, kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

template <typename VecT>
cudaError_t launch_probe_ldg_vec(const ProbeArgs& a, cudaStream_t stream) {
  const unsigned grid = grid_for(a.n);
  const uint32_t V = a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  switch (V) {
    case 32:
      probe_gather_ldg_kernel<VecT, 32, 8><<<grid, kBlock, 0, stream>>>(a);
      break;
    case 16:
      probe_gather_ldg_kernel<VecT, 16, 8><<<grid, kBlock, 0, stream>>>(a);
      break;
    case 8:
      probe_gather_ldg_kernel<VecT, 8, 8><<<grid, kBlock, 0, stream>>>(a);
      break;
    case 4:
      probe_gather_ldg_kernel<VecT, 4, 4><<<grid, kBlock, 0, stream>>>(a);
      break;
    default:
      probe_gather_ldg_kernel<VecT, 0, 4><<<grid, kBlock, 0, stream>>>(a);
      break;
  }
  return cudaGetLastError();
}

template <int kWarps, int kStages>
cudaError_t launch_probe_tma_cfg(const ProbeArgs& a, uint32_t tiles_per_warp, cudaStream_t stream) {
  const uint32_t row_bytes = a.dim * 4u;
  const size_t smem = static_cast<size_t>(kWarps) * kStages * kTmaTileKeys * row_bytes +
                      kWarps * kStages * sizeof(uint64_t);
  if (smem > 224 * 1024) return cudaErrorNotSupported;
  static size_t attr_smem = 0;  // largest size this instantiation was configured for
  if (smem > attr_smem) {
    const cudaError_t e =
        cudaFuncSetAttribute(probe_gather_tma_kernel<kWarps, kStages>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  const size_t num_tiles = (a.n + kTmaTileKeys - 1) / kTmaTileKeys;
  const size_t warps = (num_tiles + tiles_per_warp - 1) / tiles_per_warp;
  const unsigned grid = static_cast<unsigned>((warps + kWarps - 1) / kWarps);
  probe_gather_tma_kernel<kWarps, kStages><<<grid, kWarps * 32, smem, stream>>>(a, tiles_per_warp);
  return cudaGetLastError();
}

// Ring shape: 4 warps x 3 stages x 8 tiles per warp (the best of the shapes measured in round 1); rows too large
// for that ring fall back to 2 x 2, then to the LDG variant.
cudaError_t launch_probe_tma(const ProbeArgs& a, cudaStream_t stream) {
  const uint32_t row_bytes = a.dim * 4u;
  if (4u * 3u * 32u * row_bytes <= 200u * 1024u) return launch_probe_tma_cfg<4, 3>(a, 8, stream);
  if (2u * 2u * 32u * row_bytes <= 200u * 1024u) return launch_probe_tma_cfg<2, 2>(a, 8, stream);
  return cudaErrorNotSupported;
}

}  // namespace

cudaError_t launch_probe_gather(const DeviceTable& t, const int64_t* d_keys, size_t n, float* d_out,
                                uint32_t epoch, bool touch, uint32_t* d_miss_count,
                                uint32_t* d_miss_pos, int64_t* d_miss_keys, int64_t* hd_miss_keys,
                                int variant, cudaStream_t stream, const uint32_t* d_pos, uint32_t pos_base,
                                void* d_out_bf16, const MissBins* bins, bool skip_miss_rows) {
  if (n == 0) return cudaSuccess;
  ProbeArgs a{};
  a.pos_base = pos_base;
  a.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16);
  if (bins != nullptr) a.bins = *bins;
  a.skip_miss_rows = skip_miss_rows ? 1 : 0;
  if (bins != nullptr && variant == kProbeTma) variant = kProbeV8;  // the TMA variant only knows the flat miss list
  if (d_out_bf16 != nullptr) {
    // the bf16 mirror is written by the 256-bit kernel only
    if (d_pos != nullptr || t.dim % 8 != 0 || (reinterpret_cast<uintptr_t>(d_out_bf16) & 15u) != 0 ||
        ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(t.values)) & 31u) != 0)
      return cudaErrorNotSupported;
    variant = kProbeV8;
  }
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.default_value = t.default_value;
  a.keys = d_keys;
  a.n = n;
  a.out = d_out;
  a.epoch = epoch;
  a.touch = touch ? 1 : 0;
  a.miss_count = d_miss_count;
  a.miss_pos = d_miss_pos;
  a.miss_keys = d_miss_keys;
  a.miss_keys_host = hd_miss_keys;
  a.src = nullptr;
  a.pos = d_pos;
  const int vb = vec_bytes(t.dim, d_out, t.values);
  if (d_pos != nullptr) {
    if (vb == 16) return launch_probe_ldg_scatter<float4>(a, stream);
    if (vb == 8) return launch_probe_ldg_scatter<float2>(a, stream);
    return launch_probe_ldg_scatter<float>(a, stream);
  }
  if (variant == kProbeV8 && t.dim % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(t.values)) & 31u) == 0) {
    const unsigned grid = grid_for(n);
    if (a.out_bf16 != nullptr) {
      if (t.dim == 128)
        probe_gather_v8_kernel<16, 4, true><<<grid, kBlock, 0, stream>>>(a);
      else
        probe_gather_v8_kernel<0, 2, true><<<grid, kBlock, 0, stream>>>(a);
      return cudaGetLastError();
    }
    if (t.dim == 128)
      probe_gather_v8_kernel<16, 4><<<grid, kBlock, 0, stream>>>(a);
    else
      probe_gather_v8_kernel<0, 2><<<grid, kBlock, 0, stream>>>(a);
    return cudaGetLastError();
  }
  if (variant == kProbeTma && vb == 16) {
    const cudaError_t e = launch_probe_tma(a, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  if (vb == 16) return launch_probe_ldg_vec<float4>(a, stream);
  if (vb == 8) return launch_probe_ldg_vec<float2>(a, stream);
  return launch_probe_ldg_vec<float>(a, stream);
}

cudaError_t launch_probe_index(const DeviceTable& t, const int64_t* d_keys, size_t n, uint32_t epoch,
                               bool touch, uint32_t* d_src, uint32_t* d_miss_count,
                               uint32_t* d_miss_pos, int64_t* d_miss_keys, int64_t* hd_miss_keys,
                               cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  ProbeArgs a{};
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.default_value = t.default_value;
  a.keys = d_keys;
  a.n = n;
  a.out = nullptr;
  a.epoch = epoch;
  a.touch = touch ? 1 : 0;
  a.miss_count = d_miss_count;
  a.miss_pos = d_miss_pos;
  a.miss_keys = d_miss_keys;
  a.miss_keys_host = hd_miss_keys;
  a.src = d_src;
  probe_index_kernel<<<grid_for(n), kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_insert_merge(const DeviceTable& t, const int64_t* d_miss_keys,
                                const uint32_t* d_miss_pos, const float* d_stage, size_t m,
                                float* d_out, bool insert, uint32_t epoch, uint32_t* d_inserted,
                                cudaStream_t stream, void* d_out_bf16) {
  if (m == 0) return cudaSuccess;
  InsertArgs a{};
  a.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16);
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.miss_keys = d_miss_keys;
  a.miss_pos = d_miss_pos;
  a.stage = d_stage;
  a.m = m;
  a.out = d_out;
  a.insert = insert ? 1 : 0;
  a.epoch = epoch;
  a.inserted = d_inserted;
  const size_t warps = m;
  const unsigned grid = static_cast<unsigned>(
      min(static_cast<size_t>(148 * 32), (warps * 32 + kBlock - 1) / kBlock));
  const int vb = vec_bytes(t.dim, d_out, d_stage, t.values);
  if (d_out_bf16 != nullptr && (vb != 16 || (reinterpret_cast<uintptr_t>(d_out_bf16) & 7u) != 0)) return cudaErrorNotSupported;
  if (vb == 16)
    insert_merge_kernel<float4><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 8)
    insert_merge_kernel<float2><<<grid, kBlock, 0, stream>>>(a);
  else
    insert_merge_kernel<float><<<grid, kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_update_values(const DeviceTable& t, const int64_t* d_keys, const float* d_stage, size_t n,
                                 uint32_t* d_updated, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>(148 * 32, (n * 32 + kBlock - 1) / kBlock));
  const int vb = vec_bytes(t.dim, d_stage, t.values);
  if (vb == 16)
    update_values_kernel<float4><<<grid, kBlock, 0, stream>>>(t.buckets, t.values, t.num_buckets, t.dim, d_keys, d_stage, n, d_updated);
  else if (vb == 8)
    update_values_kernel<float2><<<grid, kBlock, 0, stream>>>(t.buckets, t.values, t.num_buckets, t.dim, d_keys, d_stage, n, d_updated);
  else
    update_values_kernel<float><<<grid, kBlock, 0, stream>>>(t.buckets, t.values, t.num_buckets, t.dim, d_keys, d_stage, n, d_updated);
  return cudaGetLastError();
}

cudaError_t launch_pull_misses(const DeviceTable& t, const int64_t* d_miss_keys, const uint32_t* d_miss_pos,
                               const uint32_t* d_miss_count, size_t n_keys, float* d_out, float* d_stage,
                               bool insert, int insert_mode, float hit_rate_threshold, uint32_t epoch,
                               uint32_t* d_inserted, uint32_t* d_absent, const unsigned long long* d_sorted_addr,
                               const uint32_t* d_sorted_idx, size_t m_hint, cudaStream_t stream, int max_ctas_per_sm,
                               void* d_out_bf16, int64_t* d_mark_absent, float* const* batch_outs, int batch_count) {
  if (n_keys == 0) return cudaSuccess;
  if (t.index == nullptr) return cudaErrorInvalidValue;
  if (batch_count < 0 || batch_count > kMaxBatchOuts || (batch_count > 0 && (batch_outs == nullptr || d_out_bf16 != nullptr)))
    return cudaErrorInvalidValue;
  PullArgs a{};
  a.batch = batch_count;
  for (int r = 0; r < batch_count; ++r) a.batch_out[r] = batch_outs[r];
  a.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16);
  a.mark_absent = d_mark_absent;
  a.sorted_addr = d_sorted_addr;
  a.sorted_idx = d_sorted_idx;
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.default_value = t.default_value;
  a.index = t.index;
  a.index_mask = t.index_mask;
  a.sentinel_row = t.sentinel_row;
  a.miss_keys = d_miss_keys;
  a.miss_pos = d_miss_pos;
  a.miss_count = d_miss_count;
  a.n_keys = static_cast<uint32_t>(n_keys);
  a.out = d_out;
  a.stage = d_stage;
  a.insert = insert ? 1 : 0;
  a.insert_mode = insert_mode;
  a.hit_rate_threshold = hit_rate_threshold;
  a.epoch = epoch;
  a.inserted = d_inserted;
  a.absent = d_absent;
  // The miss count is only known on the device: a fixed grid of 8 CTAs per SM loops over the list.
  // PCIe needs ~130 KB in flight (51 GB/s x ~2.5 us); 9472 warps x 512 B is far more than enough.
  const int ctas_per_sm = 8, load_mode = 0, rows_per_warp = 1;
  a.load_mode = load_mode;
  const size_t warps_needed = m_hint > 0 ? m_hint : n_keys;
  const int ctas = max_ctas_per_sm > 0 ? std::min(max_ctas_per_sm, ctas_per_sm) : ctas_per_sm;
  const unsigned grid = static_cast<unsigned>(
      min(static_cast<size_t>(148 * ctas), (warps_needed * 32 + kBlock - 1) / kBlock));
  // host rows are only guaranteed 4-B aligned relative to dim; slabs are 4096-B aligned, rows dim*4 apart
  uintptr_t batch_bits = 0;
  for (int r = 0; r < batch_count; ++r) batch_bits |= reinterpret_cast<uintptr_t>(batch_outs[r]);
  const int vb = vec_bytes(t.dim, d_out, d_stage, t.values, reinterpret_cast<const void*>(batch_bits & 15u));
  if (d_out_bf16 != nullptr && (vb != 16 || (reinterpret_cast<uintptr_t>(d_out_bf16) & 7u) != 0)) return cudaErrorNotSupported;
  if (vb == 16 && rows_per_warp == 2)
    pull_misses_kernel<float4, 2><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 16)
    pull_misses_kernel<float4, 1><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 8)
    pull_misses_kernel<float2, 1><<<grid, kBlock, 0, stream>>>(a);
  else
    pull_misses_kernel<float, 1><<<grid, kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

size_t sort_misses_temp_bytes(size_t max_items) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const unsigned long long*>(nullptr),
                                  static_cast<unsigned long long*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<int>(max_items), 12, 48);
  return bytes;
}

cudaError_t launch_resolve_and_sort_misses(const DeviceTable& t, const int64_t* d_miss_keys, size_t m,
                                           unsigned long long* d_addr_tmp, uint32_t* d_idx_tmp,
                                           unsigned long long* d_addr_sorted, uint32_t* d_idx_sorted, void* d_temp,
                                           size_t temp_bytes, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  if (t.index == nullptr) return cudaErrorInvalidValue;
  resolve_rows_kernel<<<grid_for(m * 8), kBlock, 0, stream>>>(t.index, t.index_mask, t.sentinel_row, d_miss_keys,
                                                              static_cast<uint32_t>(m), d_addr_tmp, d_idx_tmp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // 4-KiB page number and above: bits [12, 48) of the host virtual address
  return cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_addr_tmp, d_addr_sorted, d_idx_tmp, d_idx_sorted,
                                         static_cast<int>(m), 12, 48, stream);
}

cudaError_t launch_pull_binned(const DeviceTable& t, const MissBins& bins, float* d_out, void* d_out_bf16,
                               float* const* batch_outs, int batch_count, uint32_t* d_absent, int grid_ctas,
                               cudaStream_t stream, int insert, uint32_t epoch, uint32_t* d_inserted, int rows_in_flight) {
  if (t.index == nullptr || bins.count == nullptr || bins.num_bins == 0 || bins.num_bins > kMaxBins || d_absent == nullptr)
    return cudaErrorInvalidValue;
  if (batch_count < 0 || batch_count > kMaxBatchOuts || (batch_count > 0 && (batch_outs == nullptr || d_out_bf16 != nullptr)))
    return cudaErrorInvalidValue;
  PullBinnedArgs a{};
  a.bins = bins;
  a.index = t.index;
  a.index_mask = t.index_mask;
  a.sentinel_row = t.sentinel_row;
  a.dim = t.dim;
  a.default_value = t.default_value;
  a.out = d_out;
  a.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16);
  a.absent = d_absent;
  a.batch = batch_count;
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.epoch = epoch;
  a.inserted = d_inserted;
  uintptr_t bits = reinterpret_cast<uintptr_t>(d_out) | (insert ? reinterpret_cast<uintptr_t>(t.values) : 0);
  for (int r = 0; r < batch_count; ++r) {
    a.batch_out[r] = batch_outs[r];
    bits |= reinterpret_cast<uintptr_t>(batch_outs[r]);
  }
  // The miss count is only known on the device: a persistent grid walks the lists.  PCIe needs ~130 KB in flight
  // (51 GB/s x ~2.5 us); one CTA per SM already keeps 1184 rows of 512 B in flight, two reach the full link rate
  // with 16-B loads (profiles/pcie_probe2_r02.txt) and leave the SMs to the probes that run beside this kernel.
  const unsigned grid = static_cast<unsigned>(std::max(1, std::min(grid_ctas, 148 * 8)));
  const int vb = vec_bytes(t.dim, reinterpret_cast<const void*>(bits & 15u));
  if (d_out_bf16 != nullptr && (vb != 16 || (reinterpret_cast<uintptr_t>(d_out_bf16) & 7u) != 0)) return cudaErrorNotSupported;
  // host rows: slabs are 4096-B aligned and rows dim*4 apart, so the row alignment is that of dim*4 (tier shards: 512-B
  // aligned row arrays)
  const bool quad = rows_in_flight >= 4 && t.dim / (vb / 4) <= 32;
  if (vb == 16 && insert && quad)
    pull_binned_kernel<float4, true, 4><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 16 && quad)
    pull_binned_kernel<float4, false, 4><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 16 && insert)
    pull_binned_kernel<float4, true, 1><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 16)
    pull_binned_kernel<float4, false, 1><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 8 && insert)
    pull_binned_kernel<float2, true, 1><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 8)
    pull_binned_kernel<float2, false, 1><<<grid, kBlock, 0, stream>>>(a);
  else if (insert)
    pull_binned_kernel<float, true, 1><<<grid, kBlock, 0, stream>>>(a);
  else
    pull_binned_kernel<float, false, 1><<<grid, kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_insert_binned(const DeviceTable& t, const MissBins& bins, const float* d_out,
                                 float* const* batch_outs, int batch_count, uint32_t epoch, uint32_t* d_inserted,
                                 cudaStream_t stream) {
  if (bins.count == nullptr || bins.num_bins == 0 || bins.num_bins > kMaxBins) return cudaErrorInvalidValue;
  if (batch_count < 0 || batch_count > kMaxBatchOuts || (batch_count > 0 && batch_outs == nullptr)) return cudaErrorInvalidValue;
  InsertBinnedArgs a{};
  a.bins = bins;
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.out = d_out;
  a.epoch = epoch;
  a.inserted = d_inserted;
  a.batch = batch_count;
  uintptr_t bits = reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(t.values);
  for (int r = 0; r < batch_count; ++r) {
    a.batch_out[r] = batch_outs[r];
    bits |= reinterpret_cast<uintptr_t>(batch_outs[r]);
  }
  const unsigned grid = 148u * 8u;
  const int vb = vec_bytes(t.dim, reinterpret_cast<const void*>(bits & 15u));
  if (vb == 16)
    insert_binned_kernel<float4><<<grid, kBlock, 0, stream>>>(a);
  else if (vb == 8)
    insert_binned_kernel<float2><<<grid, kBlock, 0, stream>>>(a);
  else
    insert_binned_kernel<float><<<grid, kBlock, 0, stream>>>(a);
  return cudaGetLastError();
}

namespace {
template <typename K>
void preload_one(K kernel, cudaError_t* e) {
  cudaFuncAttributes a;
  const cudaError_t r = cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(kernel));
  if (*e == cudaSuccess) *e = r;
}
}  // namespace

cudaError_t preload_miss_path_kernels() {
  cudaError_t e = cudaSuccess;
  preload_one(insert_merge_kernel<float4>, &e);
  preload_one(insert_merge_kernel<float2>, &e);
  preload_one(insert_merge_kernel<float>, &e);
  preload_one(pull_misses_kernel<float4, 2>, &e);
  preload_one(pull_misses_kernel<float4, 1>, &e);
  preload_one(pull_misses_kernel<float2, 1>, &e);
  preload_one(pull_misses_kernel<float, 1>, &e);
  preload_one(pull_binned_kernel<float4, false, 1>, &e);
  preload_one(pull_binned_kernel<float2, false, 1>, &e);
  preload_one(pull_binned_kernel<float, false, 1>, &e);
  preload_one(pull_binned_kernel<float4, true, 1>, &e);
  preload_one(pull_binned_kernel<float2, true, 1>, &e);
  preload_one(pull_binned_kernel<float, true, 1>, &e);
  preload_one(pull_binned_kernel<float4, false, 4>, &e);
  preload_one(pull_binned_kernel<float4, true, 4>, &e);
  preload_one(insert_binned_kernel<float4>, &e);
  preload_one(insert_binned_kernel<float2>, &e);
  preload_one(insert_binned_kernel<float>, &e);
  preload_one(resolve_rows_kernel, &e);
  return e;
}

cudaError_t launch_index_clear(IndexSlot* slots, uint64_t capacity, cudaStream_t stream) {
  if (capacity == 0) return cudaSuccess;
  index_clear_kernel<<<grid_for(capacity), kBlock, 0, stream>>>(slots, capacity);
  return cudaGetLastError();
}

cudaError_t launch_tier_fill(const int64_t* d_keys, const uint64_t* d_row_addrs, size_t n, uint32_t rank, uint32_t world,
                             size_t dim, int64_t* shard_keys, float* shard_rows, unsigned long long cap,
                             unsigned long long* d_count, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (world == 0 || rank >= world || dim == 0 || dim > 0xFFFFFFFFull) return cudaErrorInvalidValue;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n * 32 + kBlock - 1) / kBlock, 148u * 16u));
  tier_fill_kernel<<<grid, kBlock, 0, stream>>>(d_keys, d_row_addrs, n, rank, world, static_cast<uint32_t>(dim), shard_keys,
                                                shard_rows, cap, d_count);
  return cudaGetLastError();
}

cudaError_t launch_tier_fill_procedural(unsigned long long num_rows, unsigned long long seed, uint32_t rank, uint32_t world,
                                        size_t dim, int64_t* shard_keys, float* shard_rows, unsigned long long cap,
                                        unsigned long long* d_count, cudaStream_t stream) {
  if (num_rows == 0) return cudaSuccess;
  if (world == 0 || rank >= world || dim == 0 || dim > 0xFFFFFFFFull) return cudaErrorInvalidValue;
  const unsigned long long chunks = (num_rows + kFillKeysPerCta - 1) / kFillKeysPerCta;
  const unsigned grid = static_cast<unsigned>(std::min<unsigned long long>(chunks, 148ull * 16ull));
  tier_fill_procedural_kernel<<<grid, kBlock, 0, stream>>>(0ull, num_rows, seed, rank, world, static_cast<uint32_t>(dim),
                                                           shard_keys, shard_rows, cap, d_count);
  return cudaGetLastError();
}

cudaError_t launch_index_insert_shard(IndexSlot* slots, uint64_t mask, const int64_t* shard_keys, const float* shard_rows,
                                      unsigned long long n, size_t dim, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned long long blocks = (n + kBlock - 1) / kBlock;
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
  index_insert_shard_kernel<<<static_cast<unsigned>(blocks), kBlock, 0, stream>>>(slots, mask, shard_keys, shard_rows, n,
                                                                                 static_cast<uint32_t>(dim));
  return cudaGetLastError();
}

cudaError_t launch_tier_gather(const DeviceTable& t, const int64_t* d_keys, size_t n, float* d_out, uint32_t* d_absent,
                               cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (t.index == nullptr || n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + kBlock - 1) / kBlock, 148u * 16u));  // a warp per 32-key tile
  const int vb = vec_bytes(t.dim, d_out);
  if (vb == 16)
    tier_gather_kernel<float4><<<grid, kBlock, 0, stream>>>(d_keys, static_cast<uint32_t>(n), t.index, t.index_mask, t.dim,
                                                            t.default_value, d_out, d_absent);
  else if (vb == 8)
    tier_gather_kernel<float2><<<grid, kBlock, 0, stream>>>(d_keys, static_cast<uint32_t>(n), t.index, t.index_mask, t.dim,
                                                            t.default_value, d_out, d_absent);
  else
    tier_gather_kernel<float><<<grid, kBlock, 0, stream>>>(d_keys, static_cast<uint32_t>(n), t.index, t.index_mask, t.dim,
                                                           t.default_value, d_out, d_absent);
  return cudaGetLastError();
}

cudaError_t launch_index_repoint(IndexSlot* slots, uint64_t mask, const int64_t* shard_keys, const float* shard_rows,
                                 unsigned long long n, size_t dim, unsigned long long* d_repointed, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  index_repoint_kernel<<<grid_for(n), kBlock, 0, stream>>>(slots, mask, shard_keys, shard_rows, n, static_cast<uint32_t>(dim),
                                                           d_repointed);
  return cudaGetLastError();
}

cudaError_t launch_index_build(IndexSlot* slots, uint64_t mask, const int64_t* d_keys,
                               const uint64_t* d_row_addrs, size_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  index_build_kernel<<<grid_for(n), kBlock, 0, stream>>>(slots, mask, d_keys, d_row_addrs, n);
  return cudaGetLastError();
}

namespace {
template <typename VecT>
cudaError_t launch_pooled_vec(const DeviceTable& t, const uint32_t* d_src, const float* d_stage,
                              size_t num_bags, size_t hotness, bool mean, float* d_pooled,
                              cudaStream_t stream) {
  const uint32_t V = t.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const uint32_t h = static_cast<uint32_t>(hotness);
#define HPSX_POOLED(G)                                                                          \
  do {                                                                                          \
    const size_t warps = (num_bags + (32 / G) - 1) / (32 / G);                                  \
    pooled_gather_kernel<VecT, G><<<grid_for(warps * 32), kBlock, 0, stream>>>(                 \
        t.values, d_stage, d_src, num_bags, h, t.dim, mean ? 1 : 0, d_pooled);                  \
  } while (0)
  if (V >= 32)
    HPSX_POOLED(32);
  else if (V >= 16)
    HPSX_POOLED(16);
  else if (V >= 8)
    HPSX_POOLED(8);
  else if (V >= 4)
    HPSX_POOLED(4);
  else if (V >= 2)
    HPSX_POOLED(2);
  else
    HPSX_POOLED(1);
#undef HPSX_POOLED
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_pooled_gather(const DeviceTable& t, const uint32_t* d_src, const float* d_stage,
                                 size_t num_bags, size_t hotness, bool mean, float* d_pooled,
                                 cudaStream_t stream) {
  if (num_bags == 0) return cudaSuccess;
  const int vb = vec_bytes(t.dim, d_pooled, d_stage, t.values);
  if (vb == 16)
    return launch_pooled_vec<float4>(t, d_src, d_stage, num_bags, hotness, mean, d_pooled, stream);
  if (vb == 8)
    return launch_pooled_vec<float2>(t, d_src, d_stage, num_bags, hotness, mean, d_pooled, stream);
  return launch_pooled_vec<float>(t, d_src, d_stage, num_bags, hotness, mean, d_pooled, stream);
}

cudaError_t launch_table_clear(const DeviceTable& t, cudaStream_t stream) {
  if (t.num_buckets == 0) return cudaSuccess;
  table_clear_kernel<<<grid_for(static_cast<size_t>(t.num_buckets) * 8), kBlock, 0, stream>>>(
      t.buckets, t.num_buckets);
  return cudaGetLastError();
}

cudaError_t launch_count_resident(const DeviceTable& t, unsigned long long* d_count,
                                  cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  count_resident_kernel<<<grid_for(static_cast<size_t>(t.num_buckets) * 8), kBlock, 0, stream>>>(
      t.buckets, t.num_buckets, d_count);
  return cudaGetLastError();
}

cudaError_t launch_dump_keys(const DeviceTable& t, int64_t* d_keys, unsigned long long cap,
                             unsigned long long* d_count, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  dump_keys_kernel<<<grid_for(static_cast<size_t>(t.num_buckets) * 8), kBlock, 0, stream>>>(
      t.buckets, t.num_buckets, d_keys, cap, d_count);
  return cudaGetLastError();
}

cudaError_t launch_unique(const int64_t* d_keys, size_t n, int64_t* ws_keys, uint32_t* ws_ids,
                          size_t cap, int64_t* d_unique, uint32_t* d_inverse, uint32_t* d_counter,
                          cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(d_counter, 0, 4 * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  fill_empty_kernel<<<grid_for(cap), kBlock, 0, stream>>>(ws_keys, cap);
  unique_claim_kernel<<<grid_for(n), kBlock, 0, stream>>>(d_keys, n, ws_keys, ws_ids, cap - 1,
                                                          d_unique, d_counter);
  unique_resolve_kernel<<<grid_for(n), kBlock, 0, stream>>>(d_keys, n, ws_keys, ws_ids, cap - 1,
                                                            d_inverse, d_counter);
  return cudaGetLastError();
}

cudaError_t launch_route_keys(const int64_t* d_keys, size_t n, uint32_t num_shards,
                              int64_t* d_routed_keys, uint32_t* d_perm, uint32_t* d_counts,
                              uint32_t* d_cursor, cudaStream_t stream) {
  if (num_shards == 0 || num_shards > kMaxShards) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(d_counts, 0, num_shards * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d_cursor, 0, num_shards * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((n + kRouteChunk - 1) / kRouteChunk);
  route_hist_kernel<<<grid, kBlock, 0, stream>>>(d_keys, n, num_shards, d_counts);
  route_scatter_kernel<<<grid, kBlock, 0, stream>>>(d_keys, n, num_shards, d_counts, d_cursor,
                                                    d_routed_keys, d_perm);
  return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const float* d_rows, const uint32_t* d_perm, size_t n, size_t dim,
                                float* d_out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int vb = vec_bytes(static_cast<uint32_t>(dim), d_rows, d_out);
  if (vb == 16) {
    const uint32_t V = static_cast<uint32_t>(dim / 4);
    scatter_rows_kernel<float4><<<grid_for(n * V), kBlock, 0, stream>>>(d_rows, d_perm, n, V, d_out);
  } else if (vb == 8) {
    const uint32_t V = static_cast<uint32_t>(dim / 2);
    scatter_rows_kernel<float2><<<grid_for(n * V), kBlock, 0, stream>>>(d_rows, d_perm, n, V, d_out);
  } else {
    const uint32_t V = static_cast<uint32_t>(dim);
    scatter_rows_kernel<float><<<grid_for(n * V), kBlock, 0, stream>>>(d_rows, d_perm, n, V, d_out);
  }
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* d_table, const uint32_t* d_idx, size_t n, size_t dim,
                               float* d_out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (dim != 128 || ((reinterpret_cast<uintptr_t>(d_table) | reinterpret_cast<uintptr_t>(d_out)) & 15u) != 0)
    return cudaErrorNotSupported;  // measurement primitive: the benchmark's row shape only
  gather_rows_kernel<32, 8><<<grid_for(n), kBlock, 0, stream>>>(reinterpret_cast<const float4*>(d_table), d_idx, n,
                                                                 reinterpret_cast<float4*>(d_out));
  return cudaGetLastError();
}

cudaError_t launch_synth_rows(const int64_t* d_keys, size_t n, size_t dim, uint64_t seed,
                              float* d_rows, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  synth_rows_kernel<<<grid_for(n * dim), kBlock, 0, stream>>>(d_keys, n, static_cast<uint32_t>(dim),
                                                              seed, d_rows);
  return cudaGetLastError();
}


}  // namespace hpsx
