// filter.cuh -- the delegate-side pre-filters of BASELINE configs[3] fused in front of the solve.
//
// Before a delegate daemon asks the scheduler for a grant it (1) consults the compilation cache's bloom filter
// (yadcc/daemon/local/distributed_cache_reader.cc:70-77: a possible hit is served from the cache, no grant is asked) and
// (2) looks the task's digest up among the tasks already running (running_task_keeper.cc:67-75, caller
// distributed_task_dispatcher.cc:257: an identical translation unit being compiled somewhere is joined).  Only what
// passes both is offered to the scheduler.  With the whole queue in HBM the three stages are one pipeline: bloom
// probes (bloom.cuh), index probes (running_index.cuh), an order-preserving compaction of the survivors (here), the
// solve.
//
//   k_keep_count    verdict per request (0 offered, 1 cache hit, 2 joined) + survivors per tile of 1024
//   (k_scan_u32     exclusive scan of the tile counts, total behind them)
//   k_keep_scatter  survivors -> the solver's queue, FIFO order kept
#pragma once
#include "common.cuh"

namespace yd {

__global__ void __launch_bounds__(1024) k_keep_count(const uint8_t* __restrict__ bloom_hit /* may be null */,
                                                     const uint4* __restrict__ rt_hit /* yd_running_hit, may be null */,
                                                     uint32_t n, uint8_t* __restrict__ verdict,
                                                     uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t warp_cnt[32];
  const uint32_t q = blockIdx.x * 1024 + threadIdx.x;
  uint32_t v = 3;  // beyond the queue's end
  if (q < n) {
    v = (bloom_hit && bloom_hit[q]) ? 1u : (rt_hit && rt_hit[q].w) ? 2u : 0u;  // the cache is consulted first
    verdict[q] = (uint8_t)v;
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, v == 0);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t c = __reduce_add_sync(0xffffffffu, warp_cnt[threadIdx.x]);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) tile_cnt[gridDim.x] = 0;  // the scan's end cell
}

__global__ void __launch_bounds__(1024) k_keep_scatter(const yd_task_req* __restrict__ in, const uint8_t* __restrict__ verdict,
                                                       const uint32_t* __restrict__ tile_off, uint32_t n,
                                                       yd_task_req* __restrict__ out) {
  __shared__ uint32_t warp_cnt[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = blockIdx.x * 1024 + tid;
  const bool keep = q < n && verdict[q] == 0;
  const uint32_t bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  if (!keep) return;
  uint32_t before = tile_off[blockIdx.x];
  for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[w];
  before += __popc(bal & ((1u << lane) - 1));
  const uint2* src = reinterpret_cast<const uint2*>(in + q);
  uint2* dst = reinterpret_cast<uint2*>(out + before);
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}

}  // namespace yd
