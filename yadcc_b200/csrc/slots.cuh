// slots.cuh -- the per-solve slot-code table: the cost column of every servant.
//
// For fixed heartbeat facts a servant's pick key depends only on its
// running_tasks value r (UnsafeTryPickServantFor, task_dispatcher.cc:417-451;
// dedicated predicate :405-409; GetCapacityAvailable :283-313):
//
//   key(s, r) = ( tier(s,r), double(r)/cap(s,r), registry position )
//
// and the servant is free on a prefix [run[s], free_end(s)) of r values.  One
// u32 per (servant, r) encodes (tier, r/cap) order-exactly:
//
//   code = tier << 30 | floor(r * 2^27 / cap) << 3   (cap <= 8192; low 3 bits are
//                                                      stamped by the solver)
// or, for larger capacities, a 64-bit word holding the reference's own double
// r/cap (see the kWide branch below).
//
// Row s of the table holds the codes for r = run[s] .. run[s]+len-1, followed (row-scan
// solver only) by a kFull sentinel, so that solver advances a servant by bumping one index.
// Algorithmic bytes: 4 B written per slot; facts read once (20 B per servant).
#pragma once
#include "common.cuh"

namespace yd {

// Single CTA: row lengths + exclusive scan -> row offsets.  S is a few thousand,
// so one 1024-thread block with a running carry is launch-latency bound anyway.
// `sentinel` = 1 appends the kFull end marker to every row (row-scan solver); the slot-stream
// solver sorts the table and wants slots only.
__global__ void __launch_bounds__(1024) k_slot_rows(uint32_t S, const DynParams* __restrict__ dp,
                                                    ServantArrays sv, uint32_t* __restrict__ row_off,
                                                    uint32_t* __restrict__ row_len,
                                                    Counters* __restrict__ counters, uint32_t sentinel,
                                                    uint32_t static_rows) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // static_rows: the rows of ALL running_tasks values [0, free_end) -- the table then depends on the heartbeat
  // facts only and is kept across solves (slots a servant has already filled are filtered out per solve)
  const uint32_t n_requests = static_rows ? 0xFFFFFFFFu : dp->slot_clamp;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < S; base += 1024) {
    uint32_t s = base + tid;
    uint32_t len = 0;
    if (s < S) {
      uint32_t end = free_end(sv.max_tasks[s], sv.nproc[s], sv.load[s], sv.flags[s]);
      uint32_t r0 = static_rows ? 0u : sv.run[s];
      len = end > r0 ? end - r0 : 0;
      if (len > n_requests) len = n_requests;  // a servant cannot win more than n times
      row_len[s] = len;
    }
    uint32_t v = len + sentinel;
    if (s >= S) v = 0;
    // inclusive warp scan
    uint32_t x = v;
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
      warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    uint32_t carry = carry_s;
    uint32_t excl = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
    if (s < S) row_off[s] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) {
    row_off[S] = carry_s;
    counters->slots = carry_s;
    counters->pad[0] = counters->pad[1] = counters->pad[2] = counters->pad[3] = 0;  // solver diagnostics
  }
}

// One warp per servant row; lanes stride over r.  The u64 division is the only
// expensive instruction and it is fully parallel here, off the solver's
// dependency chain.
template <bool kWide>
__global__ void __launch_bounds__(256) k_slot_fill(uint32_t S, ServantArrays sv,
                                                   const uint32_t* __restrict__ row_off,
                                                   const uint32_t* __restrict__ row_len,
                                                   uint32_t* __restrict__ codes,
                                                   unsigned long long* __restrict__ codes_wide,
                                                   uint32_t* __restrict__ slot_owner, uint32_t sentinel,
                                                   uint32_t static_rows) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const uint32_t len = row_len[s], off = row_off[s];
  const uint32_t M = sv.max_tasks[s], P = sv.nproc[s], L = sv.load[s], fl = sv.flags[s];
  const uint32_t r0 = static_rows ? 0u : sv.run[s];
  for (uint32_t i = lane; i < len + sentinel; i += 32) {
    if (slot_owner) slot_owner[off + i] = s;  // slot-stream solver: slot -> registry position
    if (i == len) {
      if (kWide) codes_wide[off + i] = ~0ull; else codes[off + i] = kFull;
      break;
    }
    uint64_t r = (uint64_t)r0 + i;
    uint64_t cap = (uint64_t)capacity_at(M, P, L, r);  // > r by construction
    bool tier0 = (fl & kFlagDedicated) && (r * 2 < (uint64_t)P);
    if (kWide) {
      // Any capacity: keep the reference's own key, the IEEE double r/cap
      // (cc:440-441).  Non-negative doubles order like their bit patterns and
      // bits 63 (sign) and 62 (u < 2) are free for self/tier.  (For cap > 2^26 distinct fractions
      // can round to the same double; the reference then falls back to "first
      // index wins", and so do we.)
      double u = (double)r / (double)cap;
      codes_wide[off + i] = (tier0 ? 0ull : (1ull << 62)) | (unsigned long long)__double_as_longlong(u);
    } else {
      // packed key: bit 30 tier | bits 29..3 floor(r * 2^27 / cap) | bits 2..0 left for
      // the solver thread's servant index
      uint32_t frac = (uint32_t)((r << kFracBits) / cap);
      codes[off + i] = (tier0 ? 0u : kTierBit) | (frac << 3);
    }
  }
}

}  // namespace yd
