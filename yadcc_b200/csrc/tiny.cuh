// tiny.cuh -- a handful of decisions in ONE launch (dispatch latency).
//
// A batch of one to eight requests does not pay for the slot-stream pipeline: one block scores the
// servants of the request's digest component exactly as UnsafePickServantFor does
// (task_dispatcher.cc:362-451) -- eligibility :316-344, capacity :283-313, the requestor's own
// servant set aside :372-379, dedicated tier first :399-410, arg-min of double(running)/capacity
// with the first minimum winning :440-446 -- one request after the other, and writes the grants,
// the task ids and the leases itself.  The requests travel as kernel ARGUMENTS and the grants are
// written straight into pinned host memory, so the whole call is one launch and one stream
// synchronisation: no copy engine, no graph.
#pragma once
#include "classes.cuh"

namespace yd {

constexpr uint32_t kTinyMax = 8;

struct TinyArgs {
  yd_task_req reqs[kTinyMax];
  uint32_t n;
  long long now_ns;
  TopoView t;
  ServantArrays sv;
  TaskRing ring;             // .next = the first id to hand out
  yd_grant* out;             // pinned host memory (device-visible)
  unsigned long long* granted_out;  // pinned host memory
  Counters* counters;
};

__device__ __forceinline__ bool tiny_less(unsigned long long ah, uint32_t al, unsigned long long bh, uint32_t bl) {
  return ah < bh || (ah == bh && al < bl);
}

__global__ void __launch_bounds__(1024) k_solve_tiny(TinyArgs a) {
  __shared__ unsigned long long s_hi[32];
  __shared__ uint32_t s_lo[32], s_any[32];
  __shared__ uint32_t s_self;
  __shared__ uint32_t s_res[kTinyMax];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t i = 0; i < a.n; ++i) {
    const yd_task_req rq = a.reqs[i];
    const uint32_t env = rq.env_id, mv = rq.min_version, ip = rq.requestor_ip;
    uint32_t comp = kNone;
    if (env < a.t.n_envs) comp = a.t.env_comp[env];
    if (comp == kNone) {  // nobody holds that digest: EnvironmentNotFound (cc:105-108)
      if (tid == 0) s_res[i] = kResEnvNotFound;
      __syncthreads();
      continue;
    }
    // "self": the first FREE eligible servant on the requestor's IP, in registry order (cc:372-375)
    if (tid == 0) {
      uint32_t self = kNone;
      if (ip < a.t.n_ips) {
        for (uint32_t u = a.t.ip_off[ip], e = a.t.ip_off[ip + 1]; u < e && self == kNone; ++u) {
          const uint32_t p = a.t.ip_sv[u];
          if (a.t.sv_comp[p] != comp) continue;
          if (a.sv.max_tasks[p] == 0 || (uint32_t)a.sv.version[p] < mv || !servant_has_env(a.t, p, env)) continue;
          if (__ldcg(a.sv.run + p) < free_end(a.sv.max_tasks[p], a.sv.nproc[p], a.sv.load[p], a.sv.flags[p])) self = p;
        }
      }
      s_self = self;
    }
    __syncthreads();
    const uint32_t self = s_self;
    // every thread scores its share of the component's servants
    unsigned long long bh = ~0ull;
    uint32_t bl = kNone, any_elig = 0;
    for (uint32_t k = a.t.comp_sv_off[comp] + tid, e = a.t.comp_sv_off[comp + 1]; k < e; k += 1024) {
      const uint32_t p = a.t.comp_sv[k];
      const uint32_t M = a.sv.max_tasks[p];
      if (M == 0 || (uint32_t)a.sv.version[p] < mv || !servant_has_env(a.t, p, env)) continue;
      any_elig = 1;
      const uint32_t P = a.sv.nproc[p], L = a.sv.load[p], fl = a.sv.flags[p];
      const uint32_t r = __ldcg(a.sv.run + p);  // (earlier requests of this batch have bumped it: read past L1)
      if (r >= free_end(M, P, L, fl) || p == self) continue;  // full, or the requestor's own machine
      const double u = (double)r / (double)capacity_at(M, P, L, r);  // cc:440-441 (capacity > r here)
      const bool tier0 = (fl & kFlagDedicated) && ((unsigned long long)r * 2 < P);  // cc:405-409
      const unsigned long long h = (tier0 ? 0ull : (1ull << 62)) | (unsigned long long)__double_as_longlong(u);
      if (tiny_less(h, p, bh, bl)) { bh = h; bl = p; }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      const unsigned long long oh = __shfl_xor_sync(0xffffffffu, bh, d);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, bl, d);
      if (tiny_less(oh, ol, bh, bl)) { bh = oh; bl = ol; }
      any_elig |= __shfl_xor_sync(0xffffffffu, any_elig, d);
    }
    if (lane == 0) { s_hi[warp] = bh; s_lo[warp] = bl; s_any[warp] = any_elig; }
    __syncthreads();
    if (warp == 0) {
      bh = s_hi[lane]; bl = s_lo[lane]; any_elig = s_any[lane];
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        const unsigned long long oh = __shfl_xor_sync(0xffffffffu, bh, d);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, bl, d);
        if (tiny_less(oh, ol, bh, bl)) { bh = oh; bl = ol; }
        any_elig |= __shfl_xor_sync(0xffffffffu, any_elig, d);
      }
      if (lane == 0) {
        uint32_t pick = bl;
        if (pick == kNone) pick = self;  // nothing else is free: the own machine after all (cc:394-396)
        if (pick != kNone) {
          atomicAdd(&a.sv.run[pick], 1u);   // cc:123-124
          atomicAdd(&a.sv.ever[pick], 1ull);
          s_res[i] = pick;
        } else {
          s_res[i] = any_elig ? kResTimeout : kResEnvNotFound;  // cc:116-118 / :105-108
        }
      }
    }
    __threadfence_block();
    __syncthreads();
  }
  // ---- grants, task ids (FIFO ordinals), leases ---------------------------------------------------------
  if (tid < a.n) {
    uint32_t before = 0, total = 0;
    for (uint32_t k = 0; k < a.n; ++k) {
      const bool g = s_res[k] < kResTimeout;
      if (k < tid && g) ++before;
      if (g) ++total;
    }
    const uint32_t r = s_res[tid];
    yd_grant g;
    if (r < kResTimeout) {
      const uint64_t id = a.ring.next + before;
      g.task_id = a.ring.ext(id);
      g.servant_index = r;
      g.status = YD_STATUS_GRANTED;
      const uint64_t slot = id & a.ring.mask;
      a.ring.exp[slot] = a.now_ns + a.reqs[tid].expires_in_ns;
      a.ring.srv[slot] = r;
      a.ring.flags[slot] = kTaskAlive | ((a.reqs[tid].flags & YD_REQ_FLAG_PREFETCH) ? kTaskPrefetch : 0u);
    } else {
      g.task_id = 0;
      g.servant_index = YD_NO_SERVANT;
      g.status = r == kResTimeout ? YD_STATUS_TIMEOUT : YD_STATUS_ENVIRONMENT_NOT_FOUND;
    }
    a.out[tid] = g;
    if (tid == 0) {
      a.counters->granted = total;
      a.counters->alive += total;
      __threadfence_system();
      *a.granted_out = total;
    }
  }
}

}  // namespace yd
