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
#include <cstddef>
#include "tasks.cuh"

namespace yd {

// The per-call scalars.  A solo solve is ONE plain kernel launch and gets them by value (kernel parameters); the graphed
// general sequence reads them from HBM, where the graph's first node copies them.  (They are NOT read from mapped host
// memory: a dependent load over PCIe at the top of the kernel cost ~70 us apiece on the measured boxes.)
struct FusedScalars {
  DynParams dyn;
  unsigned long long seq;  // launch counter, echoed in the report
  // Zero-copy I/O, when the caller's arrays are page-locked (yd_alloc_host): device-visible addresses of the caller's
  // request array (read once, by the first phase, which leaves a copy in HBM for the later ones) and grant array
  // (written by the last phase of a solo solve) -- no copy-engine transfer before or after the kernel.  Else null.
  const void* zc_in;
  void* zc_out;
};

// The solo kernel's result, in MAPPED pinned host memory (posted writes; the host reads it after the stream has drained):
// no copy node after the kernel either.
struct FusedHostIO {
  unsigned long long done_seq;   // = the launch's seq once the record below is complete (written last)
  unsigned long long granted;    // grants of the batch
  uint32_t meta[8];              // the class table's meta words (meta[1] != 0: nothing was decided, see ClassTable)
};

struct FusedArgs {
  FusedScalars sc;              // by value ...
  const FusedScalars* sc_dev;   // ... or, if not null, in HBM
  FusedHostIO* hio;         // device-side address of the mapped result record (solo)
  DynParams* dyn_out;       // device copy of the scalars for the kernels that follow (not solo)
  unsigned long long* clean_keys;  // solo: what the last block re-initialises for the next solve: the class-table keys ...
  uint4* clean_zero;               // ... and the zeroed scratch region
  uint32_t clean_zero_vec;         //     (16-byte words)
  const yd_task_req* reqs;  // the 24-byte queue in HBM
  yd_task_req* reqs_w;      // packed upload, not solo: the 24-byte records are written here for the kernels that follow
  const uint4* reqs16;      // packed upload (yd_task_req16), or null
  uint4* reqs16_w;          // where a zero-copy read of packed requests leaves its HBM copy (= reqs16)
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
  uint32_t* bar;  // [3], zeroed per solve: arrivals, release epoch, blocks that are done
  // solo
  uint32_t solo, packed_out;
  unsigned long long* look;
  const uint32_t* comp_sv;
  TaskRing ring;
  void* out;
  Counters* counters;
  uint32_t n_servants;
  uint32_t loff_cache_words;  // dynamic shared memory of the launch, in words
  uint32_t lite;              // solo: no leader scans -- every block derives the offsets it needs from the raw counts
  unsigned long long* prof;  // debug (YDSCHED_FUSED_PROF): block 0 stamps %globaltimer at every phase boundary, else null
};

__device__ __forceinline__ unsigned long long fused_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fused_stamp(const FusedArgs& a, int k) {
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[k] = fused_now();
}

// Pull a table into L2 while the first phase streams the requests: the bench flushes L2 between solves and a scheduler
// that has been idle finds it cold too; the later phases chase indices through these tables and would pay one DRAM
// round trip per dependent load.  One 128-byte line per thread and step, spread over the whole grid.
__device__ __forceinline__ void fused_prefetch(const void* p, size_t bytes) {
  const char* base = static_cast<const char*>(p);
  for (size_t off = (size_t(blockIdx.x) * 1024 + threadIdx.x) * 128; off < bytes; off += size_t(gridDim.x) * 1024 * 128) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }
}

__device__ __forceinline__ uint32_t fused_atom_add_acq_rel(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t fused_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fused_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Arrive at barrier episode `epoch` (1, 2, ...).  Returns true in exactly one block: the last one to arrive, which has
// already acquired everybody's writes and must call fused_release after its leader work; the others call fused_wait.
// (bar.sync orders the block's accesses before thread 0's release / after its acquire: the cooperative-groups grid
// barrier with release / acquire accesses instead of full fences.)
__device__ __forceinline__ bool fused_arrive(uint32_t* bar, uint32_t epoch) {
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (fused_atom_add_acq_rel(&bar[0], 1u) + 1 == epoch * gridDim.x) ? 1u : 0u;
  __syncthreads();
  return s_last != 0;
}
__device__ __forceinline__ void fused_release(uint32_t* bar, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) fused_st_release(&bar[1], epoch);
}
__device__ __forceinline__ void fused_wait(uint32_t* bar, uint32_t epoch) {
  if (threadIdx.x == 0) {
    while (fused_ld_acquire(&bar[1]) < epoch) {}
  }
  __syncthreads();
}
// A barrier without a leader section: arrive, then wait for everybody's arrival (one hop).
__device__ __forceinline__ void fused_barrier(uint32_t* bar, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    fused_atom_add_acq_rel(&bar[0], 1u);
    const uint32_t want = epoch * gridDim.x;
    while (fused_ld_acquire(&bar[0]) < want) {}
  }
  __syncthreads();
}
// "I am done": true in the block that finishes last -- it has acquired everything the grid wrote.
__device__ __forceinline__ bool fused_done_last(uint32_t* bar) {
  __shared__ uint32_t s_fin;
  __syncthreads();
  if (threadIdx.x == 0) s_fin = (fused_atom_add_acq_rel(&bar[2], 1u) + 1 == gridDim.x) ? 1u : 0u;
  __syncthreads();
  return s_fin != 0;
}

// The result record for the host (threads 0..8 of one block): the class table's meta words and the grant count, as
// posted writes into mapped host memory -- or, `report_dev`, into HBM for a copy node to fetch.  The host reads it after
// the stream has drained, so no ordering is needed among the writes.
__device__ __forceinline__ void fused_report(const FusedArgs& a, unsigned long long seq, unsigned long long granted) {
  FusedHostIO* h = a.hio;
  const uint32_t k = threadIdx.x;
  if (k < 8) h->meta[k] = a.ct.meta[k];
  else if (k == 8) { h->granted = granted; h->done_seq = seq; }
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

// Solo solves skip the compacted lists: the request with FIFO rank k in class c takes the k-th member of c's list, and
// that member is found straight from what the count phase left behind -- the scanned (class, tile) offsets (`loff`, in
// shared memory when they fit) name the slot tile, the tile's membership ballot (32 words) names the slot, the kept
// order's record names the servant.  Returns the verdict: a REGISTRY POSITION, kResTimeout or kResEnvNotFound.
__device__ __forceinline__ uint32_t fused_select(uint32_t q, const FusedArgs& a, const uint32_t* __restrict__ loff) {
  const uint32_t c = a.rcls[q];
  if (c == kNone) return kResEnvNotFound;           // unknown digest, or one nobody holds
  if (a.ct.cls_nelig[c] == 0) return kResEnvNotFound;  // cc:105-108
  const uint32_t rank = a.rank_cnt[c * a.n_rtiles + q / kRankTile] - a.rank_cnt[c * a.n_rtiles] + a.rrank[q];
  const uint32_t* row = loff + c * a.n_ltiles;
  const uint32_t lb = row[0], le = row[a.n_ltiles];
  if (rank >= le - lb) return kResTimeout;            // cc:116-118
  const uint32_t target = lb + rank;
  uint32_t lo = 0, hi = a.n_ltiles;                   // the last tile whose offset is <= target holds it
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (row[mid] <= target) lo = mid; else hi = mid;
  }
  uint32_t j = target - row[lo];                      // the j-th member inside tile lo
  const uint4* words = reinterpret_cast<const uint4*>(a.list_bal + (size_t(lo) * a.ct.cls_bound + c) * 32);
  uint32_t w = 0, word = 0;
  bool found = false;
#pragma unroll 1
  for (uint32_t v = 0; v < 8 && !found; ++v) {
    const uint4 x = words[v];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!found) {
        const uint32_t p = __popc(xs[k]);
        if (j < p) { found = true; w = v * 4 + k; word = xs[k]; }
        else j -= p;
      }
    }
  }
  const uint32_t bit = __fns(word, 0, (int)j + 1);
  return a.dec.rec[lo * kListTile + w * 32 + bit].x;
}

// The same with block-local tables built from the RAW counts (no scan by a leader): `rows` = per class the exclusive
// offsets of its slot tiles + its total (n_ltiles + 1 words per class), `rpre[c]` = class-c requests in the request tiles
// before this one.
__device__ __forceinline__ uint32_t fused_select_lite(uint32_t q, const FusedArgs& a, const uint32_t* __restrict__ rows,
                                                       const uint32_t* __restrict__ rpre) {
  const uint32_t c = a.rcls[q];
  if (c == kNone) return kResEnvNotFound;
  if (a.ct.cls_nelig[c] == 0) return kResEnvNotFound;  // cc:105-108
  const uint32_t target = rpre[c] + a.rrank[q];
  const uint32_t* row = rows + c * (a.n_ltiles + 1);
  if (target >= row[a.n_ltiles]) return kResTimeout;   // cc:116-118
  uint32_t lo = 0, hi = a.n_ltiles;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (row[mid] <= target) lo = mid; else hi = mid;
  }
  uint32_t j = target - row[lo];
  const uint4* words = reinterpret_cast<const uint4*>(a.list_bal + (size_t(lo) * a.ct.cls_bound + c) * 32);
  uint32_t w = 0, word = 0;
  bool found = false;
#pragma unroll 1
  for (uint32_t v = 0; v < 8 && !found; ++v) {
    const uint4 x = words[v];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!found) {
        const uint32_t p = __popc(xs[k]);
        if (j < p) { found = true; w = v * 4 + k; word = xs[k]; }
        else j -= p;
      }
    }
  }
  const uint32_t bit = __fns(word, 0, (int)j + 1);
  return a.dec.rec[lo * kListTile + w * 32 + bit].x;
}

__global__ void __launch_bounds__(1024, 1) k_fused_front(FusedArgs a) {
  extern __shared__ uint32_t s_loff[];  // solo: the scanned list offsets, when a.loff_cache_words holds them
  __shared__ unsigned long long s_seen[64];
  __shared__ DynParams s_dyn;
  __shared__ unsigned long long s_seq, s_zc_in, s_zc_out;
  const uint32_t tid = threadIdx.x, G = gridDim.x;
  if (tid == 0) {
    const FusedScalars sc = a.sc_dev ? *a.sc_dev : a.sc;
    s_dyn = sc.dyn;
    s_seq = sc.seq;
    s_zc_in = reinterpret_cast<unsigned long long>(sc.zc_in);
    s_zc_out = reinterpret_cast<unsigned long long>(sc.zc_out);
    if (blockIdx.x == 0 && a.dyn_out) *a.dyn_out = sc.dyn;
  }
  __syncthreads();
  const uint32_t n = s_dyn.n;
  const uint32_t nb_live = (n + 1023) / 1024;            // request tiles that hold requests
  const uint32_t m = (uint32_t)*a.m_ptr;
  const uint32_t lt_live = min((m + kListTile - 1) / kListTile, a.n_ltiles);  // slot tiles that hold slots
  const ReqView rv{a.reqs, a.reqs16};

  fused_stamp(a, 0);
  // ---- P1: classes ---------------------------------------------------------------------------------------------
  if (tid < 64) s_seen[tid] = kClsEmpty;
  __syncthreads();
  {
    const size_t S4 = size_t(a.n_servants) * 4;
    fused_prefetch(a.dec.rec, size_t(m) * 8);
    fused_prefetch(a.sv.run, S4);
    fused_prefetch(a.sv.version, S4);
    fused_prefetch(a.sv.max_tasks, S4);
    fused_prefetch(a.t.sv_comp, S4);
    fused_prefetch(a.t.sv_local, S4);
    fused_prefetch(a.t.comp_sv, S4);
    if (a.t.sv_emask) fused_prefetch(a.t.sv_emask, S4 * 2);
    else { fused_prefetch(a.t.sv_env_off, S4 + 4); }
  }
  const char* zc_in = reinterpret_cast<const char*>(s_zc_in);
  for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
    const uint32_t q = tile * 1024 + tid;
    if (zc_in) {
      // this tile of the caller's page-locked array -> HBM, 16 bytes per thread and step (coalesced reads over PCIe);
      // every later read of the requests, here and in the kernels that follow, hits the copy
      if (a.reqs16) {
        if (q < n) a.reqs16_w[q] = reinterpret_cast<const uint4*>(zc_in)[q];
      } else {
        const uint32_t bytes = min(n - tile * 1024, 1024u) * (uint32_t)sizeof(yd_task_req);  // a multiple of 8
        const char* src = zc_in + size_t(tile) * 1024 * sizeof(yd_task_req);
        char* dst = reinterpret_cast<char*>(const_cast<yd_task_req*>(a.reqs)) + size_t(tile) * 1024 * sizeof(yd_task_req);
        for (uint32_t o = tid * 16; o < bytes; o += 1024 * 16) {
          if (o + 16 <= bytes) *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(src + o);
          else *reinterpret_cast<uint2*>(dst + o) = *reinterpret_cast<const uint2*>(src + o);
        }
      }
      __syncthreads();
    }
    if (q < n) {
      uint32_t env, mv, ip;
      if (a.reqs16) {
        const uint4 w = a.reqs16[q];
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

  fused_stamp(a, 1);
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
  if (*reinterpret_cast<volatile uint32_t*>(&a.ct.meta[1]) != 0) {
    if (a.solo && blockIdx.x == 0) fused_report(a, s_seq, 0);
    return;
  }
  const uint32_t ncls = min(a.ct.meta[0], a.ct.cls_bound);
  const uint32_t nlists = min(a.ct.meta[3], a.ct.cls_bound);
  fused_stamp(a, 2);

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

  fused_stamp(a, 3);
  // ---- E2: both count matrices -> offsets (class-major, tile-minor, + the end cell) -------------------------------
  const bool lite = a.solo && a.lite && nlists * (a.n_ltiles + 1) <= a.loff_cache_words;  // (the same in every block)
  if (lite) {
    fused_barrier(a.bar, 2);  // the counts stay raw: each block derives what it needs below
  } else if (fused_arrive(a.bar, 2)) {
    fused_scan_flat(a.rank_cnt, ncls * a.n_rtiles + 1);
    fused_scan_flat(a.list_cnt, nlists * a.n_ltiles + 1);
    fused_release(a.bar, 2);
  } else {
    fused_wait(a.bar, 2);
  }

  fused_stamp(a, 4);
  if (!a.solo) {
  // ---- P5: per-class sorted slot lists (the coupled solvers read them) -------------------------------------------
  for (uint32_t tile = blockIdx.x; tile < lt_live; tile += G) {
    list_fill_tile(tile, m, a.dec, a.t, a.ct, a.n_ltiles, a.list_cnt, a.list_bal, a.list, a.list_cap);
  }

  fused_stamp(a, 5);
  // ---- B3 ---------------------------------------------------------------------------------------------------------
  if (fused_arrive(a.bar, 3)) fused_release(a.bar, 3);
  else fused_wait(a.bar, 3);
  if (*reinterpret_cast<volatile uint32_t*>(&a.ct.meta[1]) != 0) return;  // a list outgrew its buffer: nothing is decided
  }

  fused_stamp(a, 6);
  // ---- P6: verdicts (+ solo: ids, grants, leases) -----------------------------------------------------------------
  const long long now_ns = s_dyn.now_ns;
  TaskRing ring = a.ring;
  ring.next = s_dyn.ring_next;
  void* const out = s_zc_out ? reinterpret_cast<void*>(s_zc_out) : a.out;
  if (a.solo && lite) {
    // every component with requests is data-parallel: no lists, the members are selected from the ballots; the tables
    // the selection searches are built here, per block, from the raw (class, tile) counts
    __shared__ uint32_t s_rpre[kMaxClasses];
    const uint32_t lane = tid & 31, warp = tid >> 5, stride = a.n_ltiles + 1;
    for (uint32_t c = warp; c < nlists; c += 32) {  // row-local exclusive offsets of class c's slot tiles + its total
      uint32_t running = 0;
      for (uint32_t t0 = 0; t0 < a.n_ltiles; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t v = t < a.n_ltiles ? a.list_cnt[c * a.n_ltiles + t] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
          if (lane >= d) x += y;
        }
        if (t < a.n_ltiles) s_loff[c * stride + t] = running + x - v;
        running += __shfl_sync(0xffffffffu, x, 31);
      }
      if (lane == 0) s_loff[c * stride + a.n_ltiles] = running;
    }
    for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
      __syncthreads();  // (s_loff is complete; s_rpre of the previous tile has been consumed)
      for (uint32_t c = warp; c < ncls; c += 32) {  // class-c requests in the tiles before this one
        uint32_t sum = 0;
        for (uint32_t t = lane; t < tile; t += 32) sum += a.rank_cnt[c * a.n_rtiles + t];
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) s_rpre[c] = sum;
      }
      __syncthreads();
      const uint32_t q = tile * 1024 + tid;
      const uint32_t r = q < n ? fused_select_lite(q, a, s_loff, s_rpre) : kResEnvNotFound;
      if (a.packed_out) final_tile<true, true, true>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, out, a.counters, a.sv.run, a.sv.ever);
      else final_tile<false, true, true>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, out, a.counters, a.sv.run, a.sv.ever);
    }
  } else if (a.solo) {
    // (tables too big for shared memory, or YDSCHED_FUSED_NOLITE: offsets scanned by the leader of E2)
    const uint32_t cells = nlists * a.n_ltiles + 1;
    const uint32_t* loff = a.list_cnt;
    if (cells <= a.loff_cache_words) {
      for (uint32_t i = tid; i < cells; i += 1024) s_loff[i] = a.list_cnt[i];
      __syncthreads();
      loff = s_loff;
    }
    for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
      const uint32_t q = tile * 1024 + tid;
      const uint32_t r = q < n ? fused_select(q, a, loff) : kResEnvNotFound;
      if (a.packed_out) final_tile<true, true>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, out, a.counters, a.sv.run, a.sv.ever);
      else final_tile<false, true>(tile, nb_live - 1, r, n, now_ns, rv, a.look, a.comp_sv, ring, out, a.counters, a.sv.run, a.sv.ever);
    }
  } else {
    for (uint32_t tile = blockIdx.x; tile < nb_live; tile += G) {
      const uint32_t q = tile * 1024 + tid;
      uint32_t v;
      if (q < n && rank_assign_one(q, a.n_rtiles, a.t, a.ct, a.rcls, a.rrank, a.rself, a.rank_cnt, a.list_cnt, a.n_ltiles,
                                   a.list, a.comp_mode, a.rq, a.L, v)) {
        a.res[q] = v;
      }
    }
  }
  fused_stamp(a, 7);
  // ---- solo: the block that finishes last reports to the host and leaves the scratch as the next solve expects it ----
  if (a.solo && fused_done_last(a.bar)) {
    fused_report(a, s_seq, a.counters->granted);
    __syncthreads();  // (the report reads meta[], which lies in the region zeroed below)
    for (uint32_t i = tid; i < kClsTableSize; i += 1024) a.clean_keys[i] = kClsEmpty;
    for (uint32_t i = tid; i < a.clean_zero_vec; i += 1024) a.clean_zero[i] = make_uint4(0u, 0u, 0u, 0u);
    if (a.prof && tid == 0) a.prof[8] = fused_now();
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
