// fused.cuh -- the front of the slot-stream pipeline as ONE persistent kernel.
//
// At the headline size (100 k requests x 2 k servants) every kernel of the pipeline in ydsched.cu:LaunchStream does a
// few microseconds of work and costs a few more to launch and drain: ten dependent kernels are ten launch latencies
// (profiles/r2_launches_cfg2-mod.csv: 13 kernels, sum 88 us, ~3.6 us for a kernel that does nothing).  Here the same
// device functions (classes.cuh, parallel.cuh, tasks.cuh) run as phases of one co-resident grid -- one block of 1024
// threads per SM, tiles handed out round-robin -- separated by grid barriers:
//
//   P1  class table insert of every request                       (cls_insert_one; unpacks a 16-byte upload)
//   E1  barrier; the LAST block to arrive numbers the classes and picks the solver modes (cls_finalize_block)
//   P3  per-tile FIFO rank counts, per-tile list membership ballots + counts, per-class eligible counts
//   E2  barrier; the last block to arrive scans both (class, tile) count matrices
//   P5  per-class sorted slot lists                                 (list_fill_tile)
//   B3  barrier
//   P6  verdicts of the data-parallel components, FIFO records of the merge components (rank_assign_one) and --
//       `solo`, when every component with requests is data-parallel -- task ids (look-back scan over the tiles),
//       grants, leases, ++running_tasks (final_tile): the whole solve in one launch.
//
// Not solo: res[] goes to HBM and the merge / sequential solvers and k_final_fused follow as separate launches.
// A solo kernel that finds a component it cannot decide raises flag 4 and decides nothing; the host replays the batch
// with the general sequence (and remembers which one the workload needs).
//
// The barrier is the cooperative-groups pattern (bar.sync; one thread: fence, atomic arrive, spin, fence; bar.sync).  The
// grid never exceeds the number of SMs, so all blocks are resident; a block that has to wait for other kernels to drain
// first only delays the barrier.
#pragma once
#include "parallel.cuh"
#include "tasks.cuh"

namespace yd {

struct FusedArgs {
  const yd_task_req* reqs;  // the 24-byte queue in HBM
  yd_task_req* reqs_w;      // packed upload, not solo: the 24-byte records are written here for the kernels that follow
  const uint4* reqs16;      // packed upload (yd_task_req16), or null
  const DynParams* dp;
  TopoView t;
  ClassTable ct;
  ServantArrays sv;
  SlotDecode dec;
  const unsigned long long* m_ptr;  // slots in the kept order
  uint32_t* comp_mode;
  uint32_t n_comps;
  uint32_t n_rtiles, n_ltiles;  // row strides of the two count matrices (sized for the batch's / slot table's size class)
  uint32_t* rcls;
  uint32_t* rrank;
  uint32_t* rself;
  uint32_t* rank_cnt;
  uint32_t* list_cnt;
  uint32_t* list_bal;
  uint2* list;
  uint32_t list_cap;
  uint2* rq;
  uint32_t* res;
  RqLayout L;
  uint32_t* bar;  // [2], zeroed per solve: arrivals, release epoch
  // solo
  uint32_t solo, packed_out;
  unsigned long long* look;
  const uint32_t* comp_sv;
  TaskRing ring;
  void* out;
  Counters* counters;
};

// Arrive at barrier episode `epoch` (1, 2, ...).  Returns true in exactly one block: the last one to arrive, which has
// already acquired everybody's writes and must call fused_release after its leader work; the others call fused_wait.
__device__ __forceinline__ bool fused_arrive(uint32_t* bar, uint32_t epoch) {
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t old = atomicAdd(&bar[0], 1u);
    const uint32_t last = (old + 1 == epoch * gridDim.x) ? 1u : 0u;
    if (last) __threadfence();
    s_last = last;
  }
  __syncthreads();
  return s_last != 0;
}
__device__ __forceinline__ void fused_release(uint32_t* bar, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicExch(&bar[1], epoch);
  }
  __syncthreads();
}
__device__ __forceinline__ void fused_wait(uint32_t* bar, uint32_t epoch) {
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile uint32_t*>(&bar[1]) < epoch) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

// In-place exclusive scan of data[0 .. cells) by one block of 1024 threads (8 values per thread and round).
__device__ __forceinline__ void fused_scan_flat(uint32_t* __restrict__ data, uint32_t cells) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < cells; base += 1024 * 8) {
    const uint32_t i0 = base + tid * 8;
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (i0 + k < cells) ? data[i0 + k] : 0;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += v[k];
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t run = carry + (warp ? warp_sums[warp - 1] : 0) + x - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (i0 + k < cells) data[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024, 1) k_fused_front(FusedArgs a) {
  __shared__ unsigned long long s_seen[64];
  const uint32_t tid = threadIdx.x, G = gridDim.x;
  const uint32_t n = a.dp->n;
  const uint32_t nb_live = (n + 1023) / 1024;            // request tiles that hold requests
  const uint32_t m = (uint32_t)*a.m_ptr;
  const uint32_t lt_live = min((m + kListTile - 1) / kListTile, a.n_ltiles);  // slot tiles that hold slots
  const ReqView rv{a.reqs, a.reqs16};

  // ---- P1: classes ---------------------------------------------------------------------------------------------
  if (tid < 64) s_seen[tid] = kClsEmpty;
  __syncthreads();
  for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
    const uint32_t q = tile * 1024 + tid;
    if (q < n) {
      uint32_t env, mv, ip;
      if (a.reqs16) {
        const uint4 w = __ldg(a.reqs16 + q);
        env = w.x; mv = w.y; ip = w.z;
        if (a.reqs_w) {  // the kernels after this one read 24-byte records
          uint2* dst = reinterpret_cast<uint2*>(a.reqs_w + q);
          const unsigned long long ns = (unsigned long long)(w.w & 0x7fffffffu) * 1000000ull;
          dst[0] = make_uint2(env, mv);
          dst[1] = make_uint2(ip, (w.w >> 31) ? YD_REQ_FLAG_PREFETCH : 0u);
          dst[2] = make_uint2((uint32_t)ns, (uint32_t)(ns >> 32));
        }
      } else {
        rv.head(q, env, mv);
        ip = rv.ip(q);
      }
      cls_insert_one(env, mv, ip, a.t, a.ct, s_seen);
    }
  }

  // ---- E1: class numbering and solver modes, by the last block to arrive ------------------------------------------
  if (fused_arrive(a.bar, 1)) {
    cls_finalize_block(a.t, a.ct, a.n_comps, a.comp_mode, a.solo);
    fused_release(a.bar, 1);
  } else {
    fused_wait(a.bar, 1);
  }
  // Overflow (1, 2) or a solo kernel facing a coupled component (4): nothing is decided, the host replays the batch.
  // Every block reads the same value: nobody writes the flag between E1's release and the next barrier's arrival
  // ... except list_fill_tile (P5), which is behind E2; the check is repeated after B3.
  if (*reinterpret_cast<volatile uint32_t*>(&a.ct.meta[1]) != 0) return;
  const uint32_t ncls = min(a.ct.meta[0], a.ct.cls_bound);
  const uint32_t nlists = min(a.ct.meta[3], a.ct.cls_bound);

  // ---- P3: rank counts, list ballots and counts, eligible servants per class ---------------------------------------
  {
    const uint32_t items = lt_live + a.n_rtiles + ncls;
    for (uint32_t it = blockIdx.x; it < items; it += G) {
      if (it < lt_live) {
        list_count_tile(it, m, a.dec, a.t, a.ct, a.sv, a.n_ltiles, a.list_cnt, a.list_bal);
      } else if (it < lt_live + a.n_rtiles) {
        const uint32_t tile = it - lt_live;
        if (tile < nb_live) {
          rank_count_tile(tile, rv, n, a.t, a.ct, a.comp_mode, a.n_rtiles, a.rcls, a.rrank, a.rself, a.rank_cnt);
        } else if (tid < a.ct.cls_bound) {  // beyond the queue's end: empty cells (the matrix is not pre-zeroed)
          a.rank_cnt[tid * a.n_rtiles + tile] = 0;
        }
      } else {
        cls_elig_class(it - lt_live - a.n_rtiles, a.t, a.ct, a.sv);
      }
    }
  }

  // ---- E2: both count matrices -> offsets (class-major, tile-minor, + the end cell) -------------------------------
  if (fused_arrive(a.bar, 2)) {
    fused_scan_flat(a.rank_cnt, ncls * a.n_rtiles + 1);
    fused_scan_flat(a.list_cnt, nlists * a.n_ltiles + 1);
    fused_release(a.bar, 2);
  } else {
    fused_wait(a.bar, 2);
  }

  // ---- P5: per-class sorted slot lists ------------------------------------------------------------------------------
  for (uint32_t tile = blockIdx.x; tile < lt_live; tile += G) {
    list_fill_tile(tile, m, a.dec, a.t, a.ct, a.n_ltiles, a.list_cnt, a.list_bal, a.list, a.list_cap);
  }

  // ---- B3 ---------------------------------------------------------------------------------------------------------
  if (fused_arrive(a.bar, 3)) fused_release(a.bar, 3);
  else fused_wait(a.bar, 3);
  if (*reinterpret_cast<volatile uint32_t*>(&a.ct.meta[1]) != 0) return;  // a list outgrew its buffer: nothing is decided

  // ---- P6: verdicts (+ solo: ids, grants, leases) -----------------------------------------------------------------
  const long long now_ns = a.dp->now_ns;
  TaskRing ring = a.ring;
  ring.next = a.dp->ring_next;
  for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
    const uint32_t q = tile * 1024 + tid;
    uint32_t r = kResEnvNotFound;
    if (q < n) {
      uint32_t v;
      if (rank_assign_one(q, a.n_rtiles, a.t, a.ct, a.rcls, a.rrank, a.rself, a.rank_cnt, a.list_cnt, a.n_ltiles, a.list,
                          a.comp_mode, a.rq, a.L, v)) {
        r = v;
        if (!a.solo) a.res[q] = v;
      }
    }
    if (a.solo) {
      if (a.packed_out) final_tile<true>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, a.out, a.counters, a.sv.run, a.sv.ever);
      else final_tile<false>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, a.out, a.counters, a.sv.run, a.sv.ever);
    }
  }
}

// Converts a packed upload into the 24-byte queue (the sequences that do not start with k_fused_front).
__global__ void __launch_bounds__(256) k_unpack_reqs(const uint4* __restrict__ reqs16, const DynParams* __restrict__ dp,
                                                     yd_task_req* __restrict__ reqs) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= dp->n) return;
  const uint4 w = __ldg(reqs16 + q);
  uint2* dst = reinterpret_cast<uint2*>(reqs + q);
  const unsigned long long ns = (unsigned long long)(w.w & 0x7fffffffu) * 1000000ull;
  dst[0] = make_uint2(w.x, w.y);
  dst[1] = make_uint2(w.z, (w.w >> 31) ? YD_REQ_FLAG_PREFETCH : 0u);
  dst[2] = make_uint2((uint32_t)ns, (uint32_t)(ns >> 32));
}

// 16-byte grants -> 8-byte grants (the sequences that do not end inside k_fused_front).  task_id - first_id is the
// grant's FIFO ordinal when ids are dense; with strided ids (sharded deployments) it is (task_id - first) / stride.
__global__ void __launch_bounds__(256) k_pack_grants(const uint4* __restrict__ grants, const DynParams* __restrict__ dp,
                                                     TaskRing ring, uint2* __restrict__ out) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= dp->n) return;
  const uint4 g = grants[q];  // {id lo, id hi, servant, status}
  uint32_t ordinal = 0;
  if (g.w == YD_STATUS_GRANTED) {
    const unsigned long long xid = ((unsigned long long)g.y << 32) | g.x;
    unsigned long long local = 0;
    ring.loc(xid, &local);
    ordinal = (uint32_t)(local - dp->ring_next);
  }
  out[q] = make_uint2(g.z, (g.w << 30) | ordinal);
}

}  // namespace yd
