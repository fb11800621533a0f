// radix.cuh -- stable LSD radix sort of (slot code, payload) pairs.
//
// Input order is (registry position, running_tasks) -- the order the slot table is
// generated in -- so a STABLE sort by code alone yields the reference's full pick
// order (tier, utilisation, registry position): ties on the code are broken by
// position exactly like `first minimum wins` (task_dispatcher.cc:444).
//
// One pass = two kernels: per-tile digit histogram (shared-memory atomics), then a stable
// scatter in which every block derives its own global bases from the raw tile histograms
// (so there is no separate scan kernel on the critical path).  Stability inside a tile:
// warp w owns a contiguous chunk of the tile and walks it 32 elements at a time;
// __match_any_sync ranks equal digits inside a group, per-warp digit counters in
// shared memory carry the rank across groups, and a prefix over the warps' counts
// orders the warps.  Algorithmic bytes per pass: read 2 x (key + payload), write 1 x.
#pragma once
#include "common.cuh"

namespace yd {

constexpr int kRsBits = 7;
constexpr int kRsBins = 1 << kRsBits;
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItemsPerWarp = 256;                    // 8 groups of 32
constexpr int kRsTile = kRsWarps * kRsItemsPerWarp;     // 2048 elements per block

template <typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, int shift) {
  return (uint32_t)(k >> shift) & (kRsBins - 1);
}

template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads) k_rs_hist(const KeyT* __restrict__ keys,
                                                        const unsigned long long* __restrict__ n_ptr, int shift,
                                                        uint32_t nb, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[kRsBins];
  const uint32_t n = (uint32_t)*n_ptr;  // exact element count (the host sized the grid from a bound)
  for (int i = threadIdx.x; i < kRsBins; i += kRsThreads) h[i] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kRsTile;
  for (uint32_t i = threadIdx.x; i < kRsTile; i += kRsThreads) {
    uint32_t g = base + i;
    if (g < n) atomicAdd(&h[rs_digit(keys[g], shift)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRsBins; i += kRsThreads) hist[i * nb + blockIdx.x] = h[i];
}

// Generic one-block exclusive scan (in place).  Each thread owns 8 consecutive
// values (two 16-byte loads), so 8192 values need one block-wide round.  The length
// is n_static, or (*n_dyn) * per_dyn + 1 when n_dyn != nullptr (sizes that only the
// device knows, e.g. number of classes x tiles).
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ data, uint32_t n_static,
                                                   const uint32_t* __restrict__ n_dyn, uint32_t per_dyn,
                                                   uint32_t* __restrict__ total_out) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = n_dyn ? (*n_dyn) * per_dyn + 1 : n_static;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024 * kScanItems) {
    const uint32_t i0 = base + tid * kScanItems;
    uint32_t v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += v[k];
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t run = carry + (warp ? warp_sums[warp - 1] : 0) + x - sum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (i0 + k < n) data[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0 && total_out) *total_out = carry_s;
}

template <typename KeyT>
__global__ void __launch_bounds__(kRsThreads) k_rs_scatter(const KeyT* __restrict__ keys_in,
                                                           const uint32_t* __restrict__ vals_in,
                                                           const unsigned long long* __restrict__ n_ptr,
                                                           int shift, uint32_t nb,
                                                           const uint32_t* __restrict__ hist /* raw [bin][tile] */,
                                                           KeyT* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t wcnt[kRsWarps][kRsBins];  // per-warp digit counts, then running offsets
  __shared__ uint32_t dig_base[kRsBins];        // global base of (digit, this tile)
  __shared__ uint32_t wsum[kRsThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = (uint32_t)*n_ptr;
  for (int i = tid; i < kRsWarps * kRsBins; i += kRsThreads) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t wbase = blockIdx.x * kRsTile + warp * kRsItemsPerWarp;
  // pass 1: count digits of my chunk
  for (int g = 0; g < kRsItemsPerWarp; g += 32) {
    uint32_t idx = wbase + g + lane;
    if (idx < n) atomicAdd(&wcnt[warp][rs_digit(keys_in[idx], shift)], 1u);
  }
  __syncthreads();
  // This tile's global base per digit, straight from the raw tile histograms (no separate
  // scan kernel): base(d) = sum of all tiles' counts of smaller digits
  //                        + counts of digit d in earlier tiles.
  {
    uint32_t total = 0, before = 0;
    if (tid < kRsBins) {
      const uint32_t* row = hist + tid * nb;
      for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t c = row[b];
        before += b < blockIdx.x ? c : 0u;
        total += c;
      }
    }
    // exclusive scan of `total` over the digits (kRsBins <= kRsThreads)
    uint32_t x = total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w = 0; w < warp; ++w) woff += wsum[w];
    if (tid < kRsBins) dig_base[tid] = woff + x - total + before;
    __syncthreads();
  }
  // per digit: exclusive prefix over warps, offset by this tile's global base
  for (int d = tid; d < kRsBins; d += kRsThreads) {
    uint32_t run = dig_base[d];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      uint32_t c = wcnt[w][d];
      wcnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass 2: stable ranks inside my chunk, group by group
  for (int g = 0; g < kRsItemsPerWarp; g += 32) {
    uint32_t idx = wbase + g + lane;
    bool valid = idx < n;
    KeyT k = valid ? keys_in[idx] : (KeyT)0;
    uint32_t d = valid ? rs_digit(k, shift) : (uint32_t)kRsBins;  // invalid lanes form their own group
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t rank = __popc(peers & ((1u << lane) - 1));
    uint32_t dst = 0;
    if (valid) dst = wcnt[warp][d] + rank;
    __syncwarp();
    if (valid && rank == 0) wcnt[warp][d] += __popc(peers);  // leader advances the running offset
    __syncwarp();
    if (valid) {
      keys_out[dst] = k;
      vals_out[dst] = vals_in ? vals_in[idx] : idx;  // first pass: payload = original slot index
    }
  }
}

}  // namespace yd
