// solve_merge.cuh -- the merge solver: coupled components without self-requests.
//
// When several classes compete for the same servants the request-side walk
// (solve_stream.cuh) is inherently one decision at a time: other classes keep consuming
// the slots at a class's front.  For components in which NO request comes from one of the
// component's own servant IPs every request has the same preference order over slots (the
// global sorted order), so the sequential fold -- serial dictatorship in FIFO order -- is
// the unique stable matching and can equally be produced from the other side:
//
//     walk the component's slots in sorted order; each slot takes the EARLIEST unserved
//     request among the classes its servant is eligible for.
//
// [Proof sketch: both procedures produce a matching without blocking pairs w.r.t. "requests
// prefer smaller slots, slots prefer earlier requests"; with a common order on one side the
// stable matching is unique.]  Requests of one class are served in FIFO order, so the state
// is one counter h_c per class and slot j's choice is
//     argmin over c in mask_j with h_c < n_c of  rq_c[h_c].
// A warp handles 32 consecutive slots per step.  Slot j's choice depends on how many earlier
// slots of the window chose each class; the warp iterates "recompute my choice given the
// current choices of the lanes before me" until nothing changes.  Lane 0 is right after
// round 1, lane j after at most j + 1 rounds, and the fixed point is unique, so the result is
// exactly the sequential one; in practice a handful of rounds suffice.
#pragma once
#include "classes.cuh"

namespace yd {

struct MergeArgs {
  TopoView t;
  ClassTable ct;
  ServantArrays sv;
  const uint32_t* comp_mode;
  const uint32_t* list_off;  // scanned (class-major, tile-minor) list counts
  uint32_t n_list_tiles;
  const uint2* list;         // pseudo-class lists: (servant local index, class mask)
  const uint32_t* rank_off;  // scanned (class-major, tile-minor) request counts
  uint32_t n_rank_tiles;
  const uint32_t* rq;        // per-class FIFO request lists
  uint32_t* res;
};

__global__ void __launch_bounds__(32, 1) k_solve_merge(MergeArgs a) {
  const uint32_t comp = blockIdx.x;
  if (a.ct.meta[1] || a.comp_mode[comp] != 2) return;
  __shared__ uint32_t tbl[32][33];  // tbl[k][i] = request index of class k's (h_k + i)-th request, or kNone
  const uint32_t lane = threadIdx.x;
  const uint32_t lt_mask = (1u << lane) - 1;
  const uint32_t ncls = a.ct.meta[0];
  const uint32_t K = a.ct.comp_ncls[comp];
  // lane k < K owns the component's k-th class
  uint32_t cid = kNone;
  for (uint32_t c = 0; c < ncls; ++c) {
    if (a.ct.cls_comp[c] == comp && a.ct.cls_lbit[c] == lane) cid = c;
  }
  uint32_t h = 0, n = 0, rq_base = 0;
  if (cid != kNone) {
    rq_base = a.rank_off[cid * a.n_rank_tiles];
    n = a.ct.cls_nelig[cid] ? a.rank_off[(cid + 1) * a.n_rank_tiles] - rq_base : 0u;  // nobody eligible: all ENF already
  }
  const uint32_t pc = ncls + a.ct.comp_midx[comp];
  const uint32_t lb = a.list_off[pc * a.n_list_tiles], le = a.list_off[(pc + 1) * a.n_list_tiles];
  const uint32_t sv_begin = a.t.comp_sv_off[comp];

  for (uint32_t base = lb; base < le; base += 32) {
    if (__reduce_add_sync(0xffffffffu, n - h) == 0) break;  // every request is served
    const uint32_t idx = base + lane;
    const uint2 e = idx < le ? a.list[idx] : make_uint2(0, 0);  // mask 0: no slot in this lane
    // the next 32 requests of every class
    for (uint32_t k = 0; k < K; ++k) {
      const uint32_t hk = __shfl_sync(0xffffffffu, h, k), nk = __shfl_sync(0xffffffffu, n, k);
      const uint32_t bk = __shfl_sync(0xffffffffu, rq_base, k);
      tbl[k][lane] = hk + lane < nk ? a.rq[bk + hk + lane] : kNone;
    }
    __syncwarp();
    uint32_t pick = 32, best = kNone;
    for (int round = 0; round < 33; ++round) {
      uint32_t npick = 32, nbest = kNone;
      for (uint32_t k = 0; k < K; ++k) {
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, pick == k) & lt_mask);  // earlier slots on class k
        if ((e.y >> k) & 1u) {
          const uint32_t tq = tbl[k][cnt];
          if (tq < nbest) { nbest = tq; npick = k; }
        }
      }
      const bool changed = npick != pick;
      pick = npick;
      best = nbest;
      if (!__any_sync(0xffffffffu, changed)) break;
    }
    // commit: served counters, grants
    for (uint32_t k = 0; k < K; ++k) {
      const uint32_t took = __popc(__ballot_sync(0xffffffffu, pick == k));
      if (lane == k) h += took;
    }
    if (pick < 32) {
      a.res[best] = sv_begin + e.x;
      const uint32_t pos = a.t.comp_sv[sv_begin + e.x];
      atomicAdd(&a.sv.run[pos], 1u);  // ++running_tasks, ++ever_assigned_tasks (cc:123-124)
      atomicAdd(&a.sv.ever[pos], 1ull);
    }
    __syncwarp();
  }
}

}  // namespace yd
