// solve_merge.cuh -- the merge solver: the slot-side formulation of the sequential fold,
// chunk-parallel with exact speculation.
//
// The reference decides requests one at a time (task_dispatcher.cc:93-140): request i takes the
// smallest untaken slot (tier, r/cap, position) among its class's eligible servants, skipping
// the first free servant on its own IP ("self", cc:372-379) unless nothing else is free
// (cc:394-396).  Requests rank slots by the common sorted order (own servant's slots last),
// slots "rank" requests by FIFO index: with a common order on one side the stable matching is
// unique and it is what the sequential fold produces.  It can equally be produced from the slot
// side:
//
//     walk the component's slots in sorted order; each slot (s, r) takes the EARLIEST unserved
//     request i among the classes servant s is eligible for, with self(i) != s.
//
// Requests of a class are served in FIFO order except that a slot of servant s passes over
// requests from s itself; those stay "pending" (a handful at most) and are the class's first
// candidates for the next slots.  State = one served-or-passed counter h_c per class + the
// pending set.  The only way this differs from the sequential fold is the last-resort rule: a
// request that ends up unserved although a slot of its own servant went to a LATER request (or
// stayed empty) would have taken that slot.  k_merge_check detects exactly that blocking pair
// after the fact (it needs every class to be running out of slots while the requestor's own
// servant still has one -- the end game of a saturated class); the component is then decided by
// the sequential solver (solve_stream.cuh) instead.  No blocking pair => stable => identical to
// the reference.
//
// Parallelism.  The slot list of a component is cut into chunks of `chunk` slots, one warp each.
// The state at a chunk boundary is not known in advance, so it is GUESSED -- after x slots the
// served requests are (nearly) the first x requests of the component in FIFO order, so h_c = the
// number of class-c requests among those -- and every chunk runs from its guess at once.  Round
// r >= 1 re-runs every chunk whose start state differs from what its predecessor produced in
// round r-1 (Jacobi iteration on the boundary states).  After round r the first r+1 chunks are
// final, so the iteration always ends with the sequential result; it ends after 2-3 rounds when
// a wrong guess "heals" inside a chunk (discrepancies between two trajectories annihilate as
// soon as one slot can serve both affected classes; measured: a few hundred slots).  A round in
// which no chunk re-ran certifies the chain: every start state equals its predecessor's end state.
//
// Inside a chunk the warp takes 32 slots per step: lane j's choice depends on how many earlier
// lanes chose each class; the lanes iterate "recompute my choice given the lanes before me" to
// the (unique) fixed point.  A step in which a lane would pick a request from its own servant
// commits the lanes before it and handles that slot -- and every slot while requests are
// pending -- one at a time.
#pragma once
#include "classes.cuh"

namespace yd {

// Pending requests are kept as runs: a slot of servant s passes over the CONSECUTIVE requests from s
// at the head of a class queue, i.e. a range [j0, j1) of that queue with one own servant.
constexpr uint32_t kMergePend = 8;         // pending runs a state can carry
constexpr uint32_t kMergeStateWords = 60;  // h[32] | np | pad[3] | pk[8] | pj0[8] | pj1[8]
constexpr uint32_t kMsNp = 32, kMsPk = 36, kMsJ0 = 44, kMsJ1 = 52;
constexpr uint32_t kMergeSkipMax = 1u << 20;  // longest run of own-servant requests walked over (then: sequential solver)
constexpr uint32_t kMergeMaxRounds = 16;   // changed[] is indexed by round & 15

struct MergePlan {
  uint32_t* chunk_base;  // [cls_bound + 1] exclusive scan of chunks per merge component
  uint32_t* viol;        // [n_comps] 1: the component needs the sequential solver after all
  uint32_t* changed;     // [16] chunks re-run in round r (r & 15)
  uint32_t* tau;         // [S] request index that took the servant's LAST slot (kNone: still free)
};

struct MergeArgs {
  TopoView t;
  ClassTable ct;
  MergePlan mp;
  ServantArrays sv;
  const DynParams* dp;
  const uint32_t* comp_mode;
  const uint32_t* list_off;  // scanned (class-major, tile-minor) list counts
  uint32_t n_list_tiles;
  const uint2* list;         // pseudo-class lists: (servant local index, class mask)
  const uint32_t* rank_off;  // scanned (class-major, tile-minor) request counts
  uint32_t n_rank_tiles;
  const uint2* rq;           // per-class FIFO request lists: (request, own servant or kNone)
  const uint32_t* rcls;      // [n] class of each request
  const uint32_t* rself;     // [n] own servant of each request
  uint32_t* slot_pick;       // [list capacity] request that took the slot, or kNone
  uint32_t* st_in;           // [max_chunks][kMergeStateWords] start state of the chunk's last run
  uint32_t* st_out;          // [max_chunks][kMergeStateWords] end state of the chunk's last run
  uint32_t chunk;            // slots per chunk (multiple of 32)
  uint32_t max_chunks;
  uint32_t round;
  uint32_t* res;
};

// Chunks per merge component and their scan (one small block).
__global__ void __launch_bounds__(256) k_merge_plan(MergeArgs a) {
  __shared__ uint32_t s_n[kMaxClasses];
  if (a.ct.meta[1]) return;
  const uint32_t ncls = min(a.ct.meta[0], a.ct.cls_bound);
  const uint32_t nmerge = min(a.ct.meta[2], a.ct.cls_bound - ncls);
  const uint32_t tid = threadIdx.x;
  for (uint32_t m = tid; m < nmerge; m += 256) {
    const uint32_t pc = ncls + m;
    const uint32_t len = a.list_off[(pc + 1) * a.n_list_tiles] - a.list_off[pc * a.n_list_tiles];
    s_n[m] = (len + a.chunk - 1) / a.chunk;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t m = 0; m < nmerge; ++m) { a.mp.chunk_base[m] = run; run += s_n[m]; }
    a.mp.chunk_base[nmerge] = run;
  }
}

// (pseudo-class index, chunk inside the component) of global chunk t; false if t is past the end.
__device__ __forceinline__ bool merge_locate(const MergeArgs& a, uint32_t t, uint32_t& ncls, uint32_t& midx,
                                             uint32_t& b) {
  ncls = min(a.ct.meta[0], a.ct.cls_bound);
  const uint32_t nmerge = min(a.ct.meta[2], a.ct.cls_bound - ncls);
  if (t >= a.mp.chunk_base[nmerge] || t >= a.max_chunks) return false;
  uint32_t lo = 0, hi = nmerge;  // largest m with chunk_base[m] <= t
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.mp.chunk_base[mid] <= t) lo = mid; else hi = mid;
  }
  midx = lo;
  b = t - a.mp.chunk_base[lo];
  return true;
}

// Boundary guess: h_k = number of class-k requests among the first `x` requests of the component
// (classes without eligible servants do not count: they are EnvironmentNotFound and take nothing).
// Lane k < K owns class k: `cid` its class id, `base` its row start in rank_off, `n` its requests.
__device__ __forceinline__ uint32_t merge_guess(const MergeArgs& a, uint32_t comp, uint32_t K, uint32_t lane,
                                                uint32_t cid, uint32_t base, uint32_t n, uint32_t x,
                                                uint32_t* s_cnt) {
  const uint32_t nrt = a.n_rank_tiles;
  const uint32_t total = __reduce_add_sync(0xffffffffu, lane < K ? n : 0u);
  if (x >= total) return n;
  // F(tile) = component requests in tiles [0, tile); find T with F(T) <= x < F(T + 1), 32 probes per step
  uint32_t lo = 0, hi = nrt;
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + 31) / 32;
    const uint32_t cand = min(lo + lane * step, hi);
    uint32_t f = 0;
    for (uint32_t k = 0; k < K; ++k) {
      const uint32_t ck = __shfl_sync(0xffffffffu, cid, k), bk = __shfl_sync(0xffffffffu, base, k);
      const uint32_t nk = __shfl_sync(0xffffffffu, n, k);
      if (nk) f += a.rank_off[ck * nrt + cand] - bk;
    }
    const uint32_t okb = __ballot_sync(0xffffffffu, cand < hi && f <= x);  // lane 0 (cand == lo) always qualifies
    const uint32_t j = 31 - __clz(okb);
    const uint32_t nlo = __shfl_sync(0xffffffffu, cand, j);
    const uint32_t nhi = j < 31 ? min(hi, __shfl_sync(0xffffffffu, cand, (j + 1) & 31)) : hi;
    lo = nlo;
    hi = max(nhi, lo + 1);
  }
  const uint32_t T = lo;
  uint32_t before = 0;  // F(T)
  {
    const uint32_t mine = (lane < K && n) ? a.rank_off[cid * nrt + T] - base : 0u;
    before = __reduce_add_sync(0xffffffffu, mine);
  }
  uint32_t h = (lane < K && n) ? a.rank_off[cid * nrt + T] - base : 0u;
  uint32_t need = x - before;  // requests of the component to take from tile T
  s_cnt[lane] = 0;
  __syncwarp();
  const uint32_t nreq = a.dp->n;
  uint32_t cls_v[32];
#pragma unroll
  for (int it = 0; it < 32; ++it) {
    const uint32_t q = T * 1024u + it * 32u + lane;
    cls_v[it] = q < nreq ? __ldg(a.rcls + q) : kNone;
  }
#pragma unroll
  for (int it = 0; it < 32; ++it) {
    if (need == 0) break;
    const uint32_t c = cls_v[it];
    uint32_t lk = kNone;
    if (c != kNone && a.ct.cls_comp[c] == comp && a.ct.cls_nelig[c] != 0) lk = a.ct.cls_lbit[c];
    const uint32_t memb = __ballot_sync(0xffffffffu, lk != kNone);
    const bool valid = lk != kNone && (uint32_t)__popc(memb & ((1u << lane) - 1)) < need;
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? lk : 64u + lane);
    if (valid && (peers & ((1u << lane) - 1)) == 0) s_cnt[lk & 31u] += __popc(peers);
    need -= min(need, (uint32_t)__popc(memb));
    __syncwarp();
  }
  __syncwarp();
  if (lane < K && n) h = min(n, h + s_cnt[lane]);
  return h;
}

// One round: every chunk whose start state is new runs its slots.
__global__ void __launch_bounds__(32) k_merge_round(MergeArgs a) {
  __shared__ uint32_t tq[32][33];  // tq[k][i] = request index of class k's (h_k + i)-th request, or kNone
  __shared__ uint32_t ts[32][33];  // its own servant
  __shared__ uint32_t st[kMergeStateWords];
  __shared__ uint32_t s_cnt[32];
  __shared__ uint32_t s_ovf;
  const uint32_t lane = threadIdx.x;
  const uint32_t lt_mask = (1u << lane) - 1;
  if (a.ct.meta[1]) return;
  const uint32_t r = a.round;
  if (blockIdx.x == 0 && lane == 0) a.mp.changed[(r + 1) & 15u] = 0;  // (nobody reads or counts that cell during this round)
  if (r >= 2 && a.mp.changed[(r - 1) & 15u] == 0) return;  // the chain was certified by the previous round
  uint32_t ncls, midx, b;
  const uint32_t t = blockIdx.x;
  if (!merge_locate(a, t, ncls, midx, b)) return;
  const uint32_t comp = a.ct.merge_comp[midx];
  if (a.mp.viol[comp]) return;
  if (b == 0 && r > 0) return;  // the first chunk starts from the true state: final after round 0
  const uint32_t K = a.ct.comp_ncls[comp];
  const uint32_t nrt = a.n_rank_tiles;
  uint32_t cid = 0, rq_base = 0, n = 0, h = 0;
  if (lane < K) {
    cid = a.ct.comp_cls[midx * 32 + lane];
    rq_base = a.rank_off[cid * nrt];
    n = a.ct.cls_nelig[cid] ? a.rank_off[(cid + 1) * nrt] - rq_base : 0u;  // nobody eligible: all ENF already
  }
  const uint32_t pc = ncls + midx;
  const uint32_t L0 = a.list_off[pc * a.n_list_tiles], L1 = a.list_off[(pc + 1) * a.n_list_tiles];
  const uint32_t lb = L0 + b * a.chunk, le = min(L1, lb + a.chunk);
  uint32_t* my_in = a.st_in + size_t(t) * kMergeStateWords;
  uint32_t* my_out = a.st_out + size_t(t) * kMergeStateWords;

  // ---- start state -------------------------------------------------------------------------
  if (lane < kMergeStateWords - 32) st[32 + lane] = 0;
  if (lane == 0) s_ovf = 0;
  __syncwarp();
  if (b == 0) {
    h = 0;
  } else if (r == 0) {
    h = merge_guess(a, comp, K, lane, cid, rq_base, n, b * a.chunk, s_cnt);
  } else {
    const uint32_t* prev = a.st_out + size_t(t - 1) * kMergeStateWords;
    const uint32_t p0 = __ldcg(prev + lane);
    const uint32_t p1 = lane < kMergeStateWords - 32 ? __ldcg(prev + 32 + lane) : 0u;
    const uint32_t m0 = my_in[lane];
    const uint32_t m1 = lane < kMergeStateWords - 32 ? my_in[32 + lane] : 0u;
    if (__all_sync(0xffffffffu, p0 == m0 && p1 == m1)) return;  // consistent with my predecessor: nothing to do
    h = lane < K ? min(p0, n) : 0u;
    if (lane < kMergeStateWords - 32) st[32 + lane] = p1;
    __syncwarp();
    if (lane == 0) st[kMsNp] = min(st[kMsNp], kMergePend);
    __syncwarp();
  }
  my_in[lane] = lane < K ? h : 0u;
  if (lane < kMergeStateWords - 32) my_in[32 + lane] = st[32 + lane];
  uint32_t np = st[kMsNp];

  // ---- the chunk's slots -----------------------------------------------------------------------
  uint32_t base = lb;
  bool dead = false;
  while (base < le) {
    if (np == 0 && __reduce_add_sync(0xffffffffu, lane < K ? n - h : 0u) == 0) break;  // every request is served
    if (np == 0) {
      const uint32_t idx = base + lane;
      const uint2 e = idx < le ? a.list[idx] : make_uint2(0, 0);  // mask 0: no slot in this lane
      for (uint32_t k = 0; k < K; ++k) {  // the next 32 requests of every class
        const uint32_t hk = __shfl_sync(0xffffffffu, h, k), nk = __shfl_sync(0xffffffffu, n, k);
        const uint32_t bk = __shfl_sync(0xffffffffu, rq_base, k);
        const uint2 v = hk + lane < nk ? a.rq[bk + hk + lane] : make_uint2(kNone, kNone);
        tq[k][lane] = v.x;
        ts[k][lane] = v.y;
      }
      __syncwarp();
      uint32_t pick = 32, best = kNone, bself = kNone;
      for (int round = 0; round < 33; ++round) {
        uint32_t npick = 32, nbest = kNone, nself = kNone;
        for (uint32_t k = 0; k < K; ++k) {
          const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, pick == k) & lt_mask);  // earlier slots on class k
          if ((e.y >> k) & 1u) {
            const uint32_t q = tq[k][cnt];
            if (q < nbest) { nbest = q; npick = k; nself = ts[k][cnt]; }
          }
        }
        const bool moved = npick != pick;
        pick = npick; best = nbest; bself = nself;
        if (!__any_sync(0xffffffffu, moved)) break;
      }
      // a lane that would serve a request from its own servant: commit the lanes before it only
      const uint32_t cb = __ballot_sync(0xffffffffu, pick < 32 && bself == e.x);
      const uint32_t fb = cb ? (uint32_t)__ffs(cb) - 1u : 32u;
      for (uint32_t k = 0; k < K; ++k) {
        const uint32_t took = __popc(__ballot_sync(0xffffffffu, lane < fb && pick == k));
        if (lane == k) h += took;
      }
      if (lane < fb && idx < le) a.slot_pick[idx] = pick < 32 ? best : kNone;
      __syncwarp();
      base += fb;
      if (fb == 32) continue;
      if (base >= le) break;
    }
    // ---- one slot, exactly: pending requests first, own-servant requests passed over -----------------
    {
      const uint2 e = a.list[base];
      const uint32_t s = e.x;
      uint32_t cq = kNone, cj = 0, pidx = kNone;
      bool ovf = false;
      if (lane < K && ((e.y >> lane) & 1u)) {
        for (uint32_t p = 0; p < np; ++p) {  // runs of a class are in queue order: the first match is the earliest
          const uint32_t pk = st[kMsPk + p];
          if ((pk >> 24) == lane && (pk & 0xFFFFFFu) != s) { cq = a.rq[rq_base + st[kMsJ0 + p]].x; pidx = p; break; }
        }
        if (pidx == kNone) {
          uint32_t j = h;
          while (j < n) {
            const uint2 v = a.rq[rq_base + j];
            if (v.y != s) { cq = v.x; break; }
            ++j;
            if (j - h > kMergeSkipMax) { ovf = true; break; }
          }
          cj = j;
        }
      }
      if (__any_sync(0xffffffffu, ovf)) { dead = true; break; }
      const uint32_t m = __reduce_min_sync(0xffffffffu, cq);
      if (m != kNone) {
        const uint32_t wl = (uint32_t)__ffs(__ballot_sync(0xffffffffu, cq == m)) - 1u;
        if (lane == wl) {
          uint32_t cur = st[kMsNp];
          if (pidx != kNone) {
            if (++st[kMsJ0 + pidx] == st[kMsJ1 + pidx]) {  // the run is used up
              for (uint32_t p = pidx; p + 1 < cur; ++p) {
                st[kMsPk + p] = st[kMsPk + p + 1]; st[kMsJ0 + p] = st[kMsJ0 + p + 1]; st[kMsJ1 + p] = st[kMsJ1 + p + 1];
              }
              --cur;
              st[kMsPk + cur] = 0; st[kMsJ0 + cur] = 0; st[kMsJ1 + cur] = 0;
            }
          } else {
            if (cj > h) {  // [h, cj) are from servant s itself: they wait for another servant's slot
              const uint32_t key = (lane << 24) | (s & 0xFFFFFFu);
              uint32_t last = kNone;  // my class's latest run
              for (uint32_t p = 0; p < cur; ++p) if ((st[kMsPk + p] >> 24) == lane) last = p;
              if (last != kNone && st[kMsPk + last] == key && st[kMsJ1 + last] == h) {
                st[kMsJ1 + last] = cj;  // contiguous with it: one run
              } else if (cur < kMergePend) {
                st[kMsPk + cur] = key; st[kMsJ0 + cur] = h; st[kMsJ1 + cur] = cj;
                ++cur;
              } else {
                s_ovf = 1;
              }
            }
            h = cj + 1;
          }
          st[kMsNp] = cur;
        }
        __syncwarp();
        if (s_ovf) { dead = true; break; }
        np = st[kMsNp];
      }
      if (lane == 0) a.slot_pick[base] = m;
      base += 1;
    }
  }
  if (dead) {  // more passed-over requests than a state carries: the sequential solver decides this component
    if (lane == 0) a.mp.viol[comp] = 1;
    return;
  }
  for (uint32_t idx = base + lane; idx < le; idx += 32) a.slot_pick[idx] = kNone;
  my_out[lane] = lane < K ? h : 0u;
  __syncwarp();
  if (lane < kMergeStateWords - 32) my_out[32 + lane] = st[32 + lane];
  if (lane == 0 && r > 0) atomicAdd(&a.mp.changed[r & 15u], 1u);
}

// Slots -> requests: verdicts and the take time of every servant's last slot.
// `a.round` = number of rounds that were launched: unless the last one found nothing to re-run the
// chain is not certified; meta[1] = 3 makes every later kernel stand down and the host reruns the
// batch with more rounds (or the sequential solver).
__global__ void __launch_bounds__(256) k_merge_scatter(MergeArgs a) {
  if (a.ct.meta[1]) return;
  if (a.mp.changed[(a.round - 1) & 15u] != 0) {
    // (every block sees the same counter: all of them leave; the flag is only read by later kernels)
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ct.meta[1] = 3;
    return;
  }
  uint32_t ncls, midx, b;
  if (!merge_locate(a, blockIdx.x, ncls, midx, b)) return;
  const uint32_t comp = a.ct.merge_comp[midx];
  if (a.mp.viol[comp]) return;
  const uint32_t pc = ncls + midx;
  const uint32_t L0 = a.list_off[pc * a.n_list_tiles], L1 = a.list_off[(pc + 1) * a.n_list_tiles];
  const uint32_t lb = L0 + b * a.chunk, le = min(L1, lb + a.chunk);
  const uint32_t sv_begin = a.t.comp_sv_off[comp];
  for (uint32_t idx = lb + threadIdx.x; idx < le; idx += blockDim.x) {
    const uint32_t q = a.slot_pick[idx];
    const uint32_t li = sv_begin + a.list[idx].x;
    atomicMax(&a.mp.tau[a.t.comp_sv[li]], q);  // slots of one servant are taken in order: the max is its last slot's
    if (q != kNone) a.res[q] = li;
  }
}

// The last-resort rule (cc:394-396): an unserved request whose own servant still had a slot at
// its turn would have taken it.  One such request and the component goes to the sequential solver.
__global__ void __launch_bounds__(256) k_merge_check(MergeArgs a) {
  if (a.ct.meta[1]) return;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.dp->n) return;
  const uint32_t c = a.rcls[q];
  if (c == kNone) return;
  const uint32_t comp = a.ct.cls_comp[c];
  if (a.comp_mode[comp] != 2 || !(a.ct.comp_flags[comp] & 1u)) return;
  const uint32_t self = a.rself[q];
  if (self == kNone || a.res[q] != kResTimeout) return;
  const uint32_t pos = a.t.comp_sv[a.t.comp_sv_off[comp] + self];
  if (a.mp.tau[pos] <= q) return;  // every slot of the own servant went to an earlier request
  if (a.sv.max_tasks[pos] != 0 && (uint32_t)a.sv.version[pos] >= a.ct.cls_mv[c] && servant_has_env(a.t, pos, a.ct.cls_env[c])) {
    a.mp.viol[comp] = 1;
  }
}

}  // namespace yd
