// Kernels of the fused model-parallel exchange (host side: shard_group.cpp).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "device_helpers.cuh"
#include "kernels.h"

namespace hpsx {

// ================================================================================================
// Fused model-parallel exchange: one process per GPU, every rank maps the
// other ranks' exchange arenas through CUDA IPC.  A request is served with peer stores only:
//   dispatch  keys + request positions go straight into the owner's inbox (NVLink stores)
//   signal    counts + a sequence flag per (src, dst) pair; the owner spins on its own HBM
//   gather    the owner's probe+gather kernel writes every row INTO THE REQUESTER'S OUTPUT BUFFER
//   signal    a second flag tells the requester that all rows addressed to it have left
// No NCCL call and no host round trip sits between the steps.
// ================================================================================================
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Bucket this rank's keys by owner and store (key, request position) into the owner's inbox slot of this
// rank.  `cursor[o]` (local, zeroed by the caller) ends as the number of keys sent to owner o.
__global__ void __launch_bounds__(kBlock) shard_dispatch_kernel(const int64_t* __restrict__ keys, size_t n,
                                                                uint32_t world, const ShardPeers peers,
                                                                uint32_t* cursor) {
  __shared__ uint32_t h[kMaxPeers];
  __shared__ uint32_t start[kMaxPeers];
  if (threadIdx.x < kMaxPeers) h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = static_cast<size_t>(blockIdx.x) * kRouteChunk;
  constexpr int kPer = kRouteChunk / kBlock;
  int64_t k[kPer];
  uint32_t o[kPer], local[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const size_t i = base + static_cast<size_t>(j) * kBlock + threadIdx.x;
    o[j] = 0xffffffffu;
    if (i < n) {
      k[j] = keys[i];
      o[j] = owner_of(k[j], world);
      local[j] = atomicAdd(&h[o[j]], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < world) {
    const uint32_t mine = h[threadIdx.x];
    start[threadIdx.x] = mine ? atomicAdd(&cursor[threadIdx.x], mine) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (o[j] == 0xffffffffu) continue;
    const size_t i = base + static_cast<size_t>(j) * kBlock + threadIdx.x;
    const uint32_t p = start[o[j]] + local[j];
    peers.inbox_keys[o[j]][p] = k[j];
    peers.inbox_pos[o[j]][p] = static_cast<uint32_t>(i);
  }
}

// One warp.  Lane p < world: (publish) store this rank's count for peer p, then the sequence flag, into peer
// p's control block; (wait) spin on the flag cell peer p owns in OUR control block.  `status` (local):
// bit 0 = a peer reported an error / capacity overflow, bit 1 = timeout.
//   phase 0 (dispatch done): counts published; after the wait, totals are checked against `capacity`.
//   phase 1 (rows returned): flag carries this rank's error bit so that requesters learn about it.
__global__ void shard_signal_wait_kernel(const ShardPeers peers, uint32_t world, uint32_t seq, int phase,
                                         const uint32_t* cursor, const uint32_t* my_cnt, const uint32_t* my_flags,
                                         uint32_t capacity, uint32_t* status, unsigned long long timeout_ns,
                                         const uint32_t* skip_if_nonzero, uint32_t* done, uint32_t* seen) {
  // speculative return wave: enqueued right behind the gather so that a request without misses needs no host
  // round trip in between; when the gather did record misses the host resolves them first and signals later
  if (skip_if_nonzero != nullptr && *reinterpret_cast<const volatile uint32_t*>(skip_if_nonzero) != 0u) return;
  // after a timeout in the dispatch wave nobody is listening any more: publish, do not wait again
  if (phase == 1 && (*reinterpret_cast<volatile uint32_t*>(status) & 2u) != 0u) timeout_ns = 0;
  const uint32_t p = threadIdx.x;
  const bool active = p < world;
  uint32_t err = phase == 1 ? (*reinterpret_cast<volatile uint32_t*>(status) & 1u) : 0u;
  __threadfence_system();
  if (active) {
    if (phase == 0) {
      *reinterpret_cast<volatile uint32_t*>(peers.inbox_cnt[p]) = cursor[p];
      __threadfence_system();
      st_release_sys(peers.flag_dispatch[p], seq << 1);
    } else {
      st_release_sys(peers.flag_return[p], (seq << 1) | err);
    }
  }
  uint32_t got = 0, timed_out = 0;
  if (active) {
    const unsigned long long t0 = global_timer_ns();
    while (true) {
      got = ld_acquire_sys(my_flags + p);
      if ((got >> 1) == seq) break;
      if (global_timer_ns() - t0 > timeout_ns) {
        timed_out = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  if (active && timed_out && seen != nullptr) seen[p] = got;
  uint32_t bad = active ? (got & 1u) : 0u;
  uint32_t cnt = (active && phase == 0 && !timed_out) ? *reinterpret_cast<const volatile uint32_t*>(my_cnt + p) : 0u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    bad |= __shfl_xor_sync(kFull, bad, off);
    timed_out |= __shfl_xor_sync(kFull, timed_out, off);
    cnt += __shfl_xor_sync(kFull, cnt, off);
  }
  if (p == 0) {
    uint32_t st = 0;
    if (bad) st |= 1u;
    if (timed_out) st |= 3u;
    if (phase == 0 && cnt > capacity) st |= 5u;  // bit 2: this rank received more keys than it has room for
    if (st) atomicOr(status, st);
    if (done != nullptr) *done = 1u;
  }
  __threadfence_system();
}

struct InboxArgs {
  Bucket* buckets;
  const float* values;
  uint32_t num_buckets;
  uint32_t dim;
  float default_value;
  uint32_t epoch;
  int touch;
  uint32_t world;
  uint32_t rank;
  uint32_t slot_cap;            // keys per (src, dst) inbox slot
  const int64_t* inbox_keys;    // [world][slot_cap], local
  const uint32_t* inbox_pos;    // [world][slot_cap], local
  const uint32_t* inbox_cnt;    // [world], local
  const uint32_t* status;       // non-zero: skip (the error travels with the return flag)
  uint32_t* miss_count;
  uint32_t* miss_pos;           // (src << kShardPosBits) | position in src's request
  int64_t* miss_keys;
  int64_t* miss_keys_host;
};

// The owner's gather: persistent grid, warps stride over 32-key tiles of the inbox; tiles never straddle two
// sources.  Hit rows leave as 512-B peer stores into the requester's output (`peers.out[src]`), misses get
// the default vector there and are appended to the miss list with their (src, position) destination.
template <typename VecT, int kV, int kUnroll, bool kStreamStores = false>
__global__ void __launch_bounds__(kBlock, 4) probe_gather_inbox_kernel(const InboxArgs a, const ShardPeers peers) {
  // Tile order interleaves the sources (tile t -> source (t + rank) % world) so that at any moment the SMs are
  // storing to every peer and to local HBM at once: NVLink egress (measured 709 GB/s, tools/nvlink_probe.cu) and
  // the local gather overlap instead of running one after the other.  Sources with more tiles than the
  // shortest one keep their surplus for a sequential tail.
  __shared__ uint32_t tail_end[kMaxPeers];  // inclusive prefix of per-source surplus tiles
  __shared__ uint32_t sh_min, sh_total;
  if (*a.status != 0u) return;
  if (threadIdx.x == 0) {
    uint32_t mn = 0xffffffffu, acc = 0;
    for (uint32_t s = 0; s < a.world; ++s) mn = min(mn, (a.inbox_cnt[s] + 31u) / 32u);
    for (uint32_t s = 0; s < a.world; ++s) {
      acc += (a.inbox_cnt[s] + 31u) / 32u - mn;
      tail_end[s] = acc;
    }
    sh_min = mn;
    sh_total = mn * a.world + acc;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t num_tiles = sh_total;
  const uint32_t striped = sh_min * a.world;
  const uint32_t warp_global = (blockIdx.x * kBlock + threadIdx.x) >> 5;
  const uint32_t total_warps = (gridDim.x * kBlock) >> 5;
  const uint32_t V = kV > 0 ? static_cast<uint32_t>(kV) : a.dim / static_cast<uint32_t>(sizeof(VecT) / sizeof(float));
  const VecT* __restrict__ vals = reinterpret_cast<const VecT*>(a.values);
  const VecT defv = splat<VecT>(a.default_value);
  for (uint32_t tile = warp_global; tile < num_tiles; tile += total_warps) {
    uint32_t src, idx;
    if (tile < striped) {
      idx = tile / a.world;
      src = (tile - idx * a.world + a.rank) % a.world;
    } else {
      const uint32_t rem = tile - striped;
      src = 0;
      while (rem >= tail_end[src]) ++src;
      idx = sh_min + rem - (src ? tail_end[src - 1] : 0u);
    }
    const uint32_t tile_base = idx * 32u;
    const uint32_t nk = min(32u, a.inbox_cnt[src] - tile_base);
    const size_t in_off = static_cast<size_t>(src) * a.slot_cap + tile_base;
    const bool valid = lane < nk;
    const int64_t key = valid ? a.inbox_keys[in_off + lane] : kEmptyKey;
    const uint32_t dst = valid ? a.inbox_pos[in_off + lane] : 0u;
    uint32_t slot = kMissSlot;
    if (valid) slot = probe_bucket(a.buckets, a.num_buckets, key, a.epoch, a.touch != 0);
    const bool is_miss = valid && slot == kMissSlot;
    unsigned miss_mask;
    const uint32_t miss_base = warp_claim_misses(is_miss, lane, a.miss_count, &miss_mask);
    VecT* __restrict__ outv = reinterpret_cast<VecT*>(peers.out[src]);
    const uint32_t total = nk * V;
    for (uint32_t i0 = 0; i0 < total; i0 += 32u * kUnroll) {
      VecT buf[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = i0 + u * 32u + lane;
        const uint32_t kk = min(i / V, 31u);
        const uint32_t s = __shfl_sync(kFull, slot, kk);
        buf[u] = defv;
        if (i < total && s != kMissSlot) buf[u] = ld_stream(vals + static_cast<size_t>(s) * V + (i - kk * V));
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = i0 + u * 32u + lane;
        const uint32_t kk = min(i / V, 31u);
        const uint32_t d = __shfl_sync(kFull, dst, kk);
        if (i < total) {
          if (kStreamStores)
            st_stream(outv + static_cast<size_t>(d) * V + (i - kk * V), buf[u]);
          else
            outv[static_cast<size_t>(d) * V + (i - kk * V)] = buf[u];
        }
      }
    }
    if (is_miss) {
      const uint32_t r = miss_base + __popc(miss_mask & ((1u << lane) - 1u));
      a.miss_pos[r] = (src << kShardPosBits) | dst;
      a.miss_keys[r] = key;
      if (a.miss_keys_host != nullptr) a.miss_keys_host[r] = key;
    }
  }
}

// Rows of resolved misses (stage[i]) -> row (pos & mask) of the output of requester (pos >> kShardPosBits).
template <typename VecT>
__global__ void __launch_bounds__(kBlock) shard_scatter_stage_kernel(const float* __restrict__ stage,
                                                                     const uint32_t* __restrict__ miss_pos, uint32_t m,
                                                                     uint32_t V, const ShardPeers peers) {
  const size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
  const size_t row = i / V;
  if (row >= m) return;
  const uint32_t v = static_cast<uint32_t>(i - row * V);
  const uint32_t p = miss_pos[row];
  const VecT x = ld_stream(reinterpret_cast<const VecT*>(stage) + i);
  reinterpret_cast<VecT*>(peers.out[p >> kShardPosBits])[static_cast<size_t>(p & ((1u << kShardPosBits) - 1u)) * V + v] = x;
}

}  // namespace

cudaError_t preload_shard_kernels() {
  cudaError_t e = cudaSuccess;
  auto one = [&e](const void* k) {
    cudaFuncAttributes a;
    const cudaError_t r = cudaFuncGetAttributes(&a, k);
    if (e == cudaSuccess) e = r;
  };
  one(reinterpret_cast<const void*>(shard_dispatch_kernel));
  one(reinterpret_cast<const void*>(shard_signal_wait_kernel));
  one(reinterpret_cast<const void*>(probe_gather_inbox_kernel<float4, 32, 8, true>));
  one(reinterpret_cast<const void*>(probe_gather_inbox_kernel<float4, 32, 8>));
  one(reinterpret_cast<const void*>(probe_gather_inbox_kernel<float4, 0, 4>));
  one(reinterpret_cast<const void*>(probe_gather_inbox_kernel<float, 0, 4>));
  one(reinterpret_cast<const void*>(shard_scatter_stage_kernel<float4>));
  one(reinterpret_cast<const void*>(shard_scatter_stage_kernel<float>));
  return e;
}

cudaError_t launch_shard_dispatch(const int64_t* d_keys, size_t n, uint32_t world, const ShardPeers& peers,
                                  uint32_t* d_cursor, cudaStream_t stream) {
  if (world == 0 || world > kMaxPeers) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;  // d_cursor was zeroed by the caller
  const unsigned grid = static_cast<unsigned>((n + kRouteChunk - 1) / kRouteChunk);
  shard_dispatch_kernel<<<grid, kBlock, 0, stream>>>(d_keys, n, world, peers, d_cursor);
  return cudaGetLastError();
}

cudaError_t launch_shard_signal_wait(const ShardPeers& peers, uint32_t world, uint32_t seq, int phase,
                                     const uint32_t* d_cursor, const uint32_t* d_my_cnt, const uint32_t* d_my_flags,
                                     uint32_t capacity, uint32_t* d_status, unsigned long long timeout_ns,
                                     cudaStream_t stream, const uint32_t* d_skip_if_nonzero, uint32_t* d_done,
                                     uint32_t* d_seen) {
  if (world == 0 || world > kMaxPeers) return cudaErrorInvalidValue;
  shard_signal_wait_kernel<<<1, 32, 0, stream>>>(peers, world, seq, phase, d_cursor, d_my_cnt, d_my_flags, capacity,
                                                 d_status, timeout_ns, d_skip_if_nonzero, d_done, d_seen);
  return cudaGetLastError();
}

cudaError_t launch_probe_gather_inbox(const DeviceTable& t, const ShardPeers& peers, uint32_t world, uint32_t rank,
                                      uint32_t slot_cap,
                                      const int64_t* d_inbox_keys, const uint32_t* d_inbox_pos,
                                      const uint32_t* d_inbox_cnt, const uint32_t* d_status, uint32_t epoch, bool touch,
                                      uint32_t* d_miss_count, uint32_t* d_miss_pos, int64_t* d_miss_keys,
                                      int64_t* hd_miss_keys, size_t expected_keys, cudaStream_t stream) {
  if (world == 0 || world > kMaxPeers) return cudaErrorInvalidValue;
  InboxArgs a{};
  a.buckets = t.buckets;
  a.values = t.values;
  a.num_buckets = t.num_buckets;
  a.dim = t.dim;
  a.default_value = t.default_value;
  a.epoch = epoch;
  a.touch = touch ? 1 : 0;
  a.world = world;
  a.rank = rank;
  a.slot_cap = slot_cap;
  a.inbox_keys = d_inbox_keys;
  a.inbox_pos = d_inbox_pos;
  a.inbox_cnt = d_inbox_cnt;
  a.status = d_status;
  a.miss_count = d_miss_count;
  a.miss_pos = d_miss_pos;
  a.miss_keys = d_miss_keys;
  a.miss_keys_host = hd_miss_keys;
  // the received count is only known on the device: the grid is sized for an even split of the request
  const size_t tiles = (std::max<size_t>(expected_keys, 32) + 31) / 32 + world;
  // measured (profiles/): one tile per warp + 3 % slack beats a persistent grid of 4-8 CTAs/SM by ~10 %; under
  // heavier skew the surplus tiles are picked up by the stride loop
  const int ctas_per_sm = 0;
  const size_t want = ((tiles + tiles / 32) * 32 + kBlock - 1) / kBlock;
  const unsigned grid = static_cast<unsigned>(ctas_per_sm > 0 ? std::min<size_t>(148 * ctas_per_sm, want) : want);
  bool aligned = ((reinterpret_cast<uintptr_t>(t.values) | (static_cast<uintptr_t>(t.dim) * 4u)) & 15u) == 0;
  for (uint32_t p = 0; p < world; ++p) aligned = aligned && (reinterpret_cast<uintptr_t>(peers.out[p]) & 15u) == 0;
  const bool stream_stores = true;  // st.global.cs: output rows are written once and not re-read by this kernel
  if (aligned && t.dim == 128 && stream_stores)
    probe_gather_inbox_kernel<float4, 32, 8, true><<<grid, kBlock, 0, stream>>>(a, peers);
  else if (aligned && t.dim == 128)
    probe_gather_inbox_kernel<float4, 32, 8><<<grid, kBlock, 0, stream>>>(a, peers);
  else if (aligned)
    probe_gather_inbox_kernel<float4, 0, 4><<<grid, kBlock, 0, stream>>>(a, peers);
  else
    probe_gather_inbox_kernel<float, 0, 4><<<grid, kBlock, 0, stream>>>(a, peers);
  return cudaGetLastError();
}

cudaError_t launch_shard_scatter_stage(const float* d_stage, const uint32_t* d_miss_pos, size_t m, size_t dim,
                                       const ShardPeers& peers, uint32_t world, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  bool aligned = ((reinterpret_cast<uintptr_t>(d_stage) | (dim * 4u)) & 15u) == 0;
  for (uint32_t p = 0; p < world; ++p) aligned = aligned && (reinterpret_cast<uintptr_t>(peers.out[p]) & 15u) == 0;
  if (aligned) {
    const uint32_t V = static_cast<uint32_t>(dim / 4);
    shard_scatter_stage_kernel<float4><<<grid_for(m * V), kBlock, 0, stream>>>(d_stage, d_miss_pos, static_cast<uint32_t>(m), V, peers);
  } else {
    const uint32_t V = static_cast<uint32_t>(dim);
    shard_scatter_stage_kernel<float><<<grid_for(m * V), kBlock, 0, stream>>>(d_stage, d_miss_pos, static_cast<uint32_t>(m), V, peers);
  }
  return cudaGetLastError();
}

}  // namespace hpsx
