// solve_rowscan.cuh -- the row-scan assignment solver (solver 1).
//
// One CTA per *component* of the digest<->servant graph (servants that share no
// compiler digest can never compete for a request, so components are
// independent FIFO sub-queues).  Inside a component decisions are made strictly
// in request order, each one scoring EVERY servant of the component in
// parallel -- one (task x servant) matrix row per decision, exactly what the
// reference does per call (task_dispatcher.cc:93-140, :316-451), except that
// the row lives in registers:
//
//   * each solver thread owns K=8 consecutive servants of the component: their
//     current slot code (slots.cuh), the code they will have after one more
//     grant, the cursor into the slot table, and the daemon version;
//   * eligibility (ContainsEnvironmentSlow && max_tasks != 0 && version >=
//     min_version, cc:316-344) is one byte of precomputed digest-membership bits
//     per (digest, thread) plus K integer compares (cached per min_version);
//   * the pick (self-exclusion :372-379, dedicated tier :399-410, min
//     utilisation, first index wins :417-451) is a min over
//     (self|tier|code, local index): two REDUX.MIN per warp, one shared-memory
//     word per warp and one named barrier per decision for multi-warp components;
//   * only the winner's owner thread does anything afterwards, and nothing on the
//     grant path waits for memory.
//
// The CTA is warp-specialised: solver warps never touch the request queue in
// HBM.  PRODUCER warps stream it in tiles of 1024 requests with coalesced 8-byte
// loads, filter the tile to this component, answer classes already known to
// fail, resolve the requestor-ip -> self-servant lookup and compact the
// survivors (order-preserving) into a double-buffered shared-memory list of
// 16-byte descriptors; full/free hand-off uses bar.arrive / bar.sync pairs.
//
// Monotonicity used for the Timeout short-cut: inside one batch running_tasks
// only grows, so once a (digest, min_version) class found no free servant every
// later request of that class times out too (EnvironmentNotFound is static).
#pragma once
#include "common.cuh"

namespace yd {

constexpr int kK = 8;            // servants per solver thread
constexpr int kTile = 1024;      // requests per tile
constexpr int kFailSlots = 128;  // direct-mapped cache of failed (digest, min_version) classes
constexpr int kMaxProducers = 8;

// named barrier ids (0 is __syncthreads)
constexpr int kBarSolver = 1, kBarProducer = 2, kBarFull0 = 3, kBarFree0 = 5;

struct SolveArgs {
  const yd_task_req* reqs;
  const DynParams* dp;  // dp->n requests
  uint32_t* res;  // per request: index into comp_sv of the picked servant, or kRes*
  // topology (rebuilt on the host when the servant set / digests change)
  const uint32_t* env_comp;   // [n_envs] component of a digest id, kNone if nobody holds it
  const uint32_t* env_local;  // [n_envs] index of the digest inside its component
  uint32_t n_envs;
  const uint32_t* comp_sv_off;    // [C+1] offsets into comp_sv
  const uint32_t* comp_sv;        // registry positions, ascending inside a component
  const uint32_t* comp_mask_off;  // [C] byte offset of the component's membership table
  const uint32_t* comp_nwarps;    // [C] solver warps
  const uint8_t* envmask;         // per component [n_env_local][solver threads] bytes, bit j = servant tid*8+j
  const uint32_t* sv_comp;        // [S]
  const uint32_t* sv_local;       // [S] index inside its component
  const uint32_t* ip_off;         // CSR: interned requestor-ip id -> registry positions whose
  const uint32_t* ip_sv;          //      observed_location matches it (IsNetworkAddressEqual, cc:66-69)
  uint32_t n_ips;
  ServantArrays sv;
  const uint32_t* row_off;  // slot table
  const void* codes;        // uint32_t[] or unsigned long long[] (wide)
};

// Key layouts.  Both order (self, tier, utilisation) exactly like the reference's
// (last-resort self, dedicated tier, double(r)/cap) comparison chain.
//
//  packed (uint32_t, every capacity <= 8192):
//     bit 31 self | bit 30 tier | bits 29..3 floor(r * 2^27 / cap) | bits 2..0 j
//     j = index of the servant inside its owner thread, so a plain integer min over
//     a thread's 8 keys also yields WHICH servant won and "first index wins".
//     (distinct fractions with denominators <= 2^13 differ by >= 2^-26)
//  wide (unsigned long long, any capacity):
//     bit 63 self | bit 62 tier | bits 61..0 IEEE-754 bits of double(r)/cap
template <typename KeyT> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {
  static constexpr bool kPacked = true;
  static constexpr uint32_t kFullKey = 0xFFFFFFFFu;
  static constexpr uint32_t kSelf = 0x80000000u;
};
template <> struct KeyTraits<unsigned long long> {
  static constexpr bool kPacked = false;
  static constexpr unsigned long long kFullKey = ~0ull;
  static constexpr unsigned long long kSelf = 1ull << 63;
};

// Warp arg-min: returns the winning key class and the lowest local servant index
// (tid * 8 + j) that attains it; kNone if every key is FULL.
__device__ __forceinline__ void warp_argmin(uint32_t best, uint32_t tid, uint32_t& gkey, uint32_t& gli) {
  const uint32_t hi = best >> 3;  // FULL -> 0x1FFFFFFF
  gkey = __reduce_min_sync(0xffffffffu, hi);
  gli = __reduce_min_sync(0xffffffffu, (hi == gkey && gkey != 0x1FFFFFFFu) ? tid * 8 + (best & 7u) : kNone);
}
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
  uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
  uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xFFFFFFFFu);
  return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ void bar_sync(int id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// A failed class packed into one 8-byte word so that producers reading it while
// a solver thread writes it never see a torn entry:
//   bits 63..32 min_version | bit 31 status (1 = Timeout, 0 = EnvNotFound) | bits 30..0 local digest
__device__ __forceinline__ unsigned long long pack_fail(uint32_t el, uint32_t mv, uint32_t res) {
  return ((unsigned long long)mv << 32) | (res == kResTimeout ? 0x80000000u : 0u) | el;
}
__device__ __forceinline__ bool fail_hit(unsigned long long e, uint32_t el, uint32_t mv) {
  return (uint32_t)(e >> 32) == mv && ((uint32_t)e & 0x7FFFFFFFu) == el;
}
__device__ __forceinline__ uint32_t fail_res(unsigned long long e) {
  return ((uint32_t)e & 0x80000000u) ? kResTimeout : kResEnvNotFound;
}

struct TileShared {
  uint4 desc[2][kTile];  // {request index, local digest, min_version, self info}
  uint32_t cnt[2];
  uint32_t pcnt[2][32];
  unsigned long long fail[kFailSlots];
};

// Filter + order-preserving compaction of requests [base, tile_end) into
// sh.desc[buf], by the `gn` warps of one group (`gw` = my index in the group).
// Pass 0 decides which requests are "mine" (and answers classes already known to
// fail); pass 1 only replays those verdicts, because the fail cache can gain
// entries in between (solver warps run concurrently with producers) and the
// counts of the two passes must agree.
__device__ __forceinline__ void produce_tile(const SolveArgs& a, TileShared& sh, uint32_t comp, uint32_t buf,
                                             uint32_t base, uint32_t tile_end, uint32_t gw, uint32_t gn,
                                             uint32_t lane, int bar_id) {
  const uint32_t chunk = ((kTile + gn - 1) / gn + 31) & ~31u;  // <= 1024: at most 32 iterations
  const uint32_t c0 = min(tile_end, base + gw * chunk), c1 = min(tile_end, c0 + chunk);
  uint32_t minebits = 0;  // bit i: my request of iteration i goes into the list
  uint32_t my_count = 0;
  uint32_t it = 0;
#pragma unroll 2
  for (uint32_t q0 = c0; q0 < c1; q0 += 32, ++it) {
    uint32_t q = q0 + lane;
    bool mine = false;
    if (q < c1) {
      const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(a.reqs + q));  // env_id, min_version
      const uint32_t env = w0.x, mv = w0.y;
      if (env < a.n_envs && __ldg(a.env_comp + env) == comp) {
        const uint32_t el = __ldg(a.env_local + env);
        const unsigned long long fe = sh.fail[el & (kFailSlots - 1)];
        if (fail_hit(fe, el, mv)) {
          a.res[q] = fail_res(fe);  // answered right here, in parallel
        } else {
          mine = true;
        }
      }
    }
    minebits |= (mine ? 1u : 0u) << it;
    my_count += __popc(__ballot_sync(0xffffffffu, mine));
  }
  if (lane == 0) sh.pcnt[buf][gw] = my_count;
  if (gn > 1) bar_sync(bar_id, gn * 32); else __syncwarp();
  uint32_t woff = 0;
  for (uint32_t w = 0; w < gw; ++w) woff += sh.pcnt[buf][w];
  it = 0;
#pragma unroll 2
  for (uint32_t q0 = c0; q0 < c1; q0 += 32, ++it) {
    const uint32_t q = q0 + lane;
    const bool mine = (minebits >> it) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, mine);
    if (mine) {
      const uint2* rp = reinterpret_cast<const uint2*>(a.reqs + q);
      const uint2 w0 = __ldg(rp);
      const uint32_t ip = __ldg(rp + 1).x;  // requestor_ip
      uint32_t selfinfo = kNone;
      if (ip < a.n_ips) {
        uint32_t b = __ldg(a.ip_off + ip), e = __ldg(a.ip_off + ip + 1);
        if (e - b == 1) {
          uint32_t p = __ldg(a.ip_sv + b);
          if (__ldg(a.sv_comp + p) == comp) selfinfo = __ldg(a.sv_local + p);
        } else if (e - b > 1) {
          selfinfo = 0x80000000u | ip;  // several servants on that IP: resolved per decision
        }
      }
      sh.desc[buf][woff + __popc(bal & ((1u << lane) - 1))] =
          make_uint4(q, __ldg(a.env_local + w0.x), w0.y, selfinfo);
    }
    woff += __popc(bal);
  }
}

template <typename KeyT, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads, 1) k_solve_rowscan(SolveArgs a) {
  using KT = KeyTraits<KeyT>;
  constexpr KeyT FULLK = KT::kFullKey;
  const uint32_t comp = blockIdx.x;
  const uint32_t nwarps = a.comp_nwarps[comp];
  const uint32_t nprod = min((uint32_t)kMaxProducers, 32u - nwarps);  // 0 => solver warps produce inline
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= nwarps + nprod) return;  // surplus warps leave at once
  const uint32_t nthreads = nwarps * 32;               // solver threads
  const uint32_t nall = (nwarps + nprod) * 32;         // solver + producer threads
  const bool multi = nwarps > 1;
  const bool specialised = nprod > 0;

  __shared__ TileShared sh;
  __shared__ KeyT s_exch_key[2][32];  // packed: key class (key >> 3); wide: full key
  __shared__ uint32_t s_exch_li[2][32];
  __shared__ uint32_t s_selfmin, s_any;

  for (uint32_t i = tid; i < kFailSlots; i += nall) sh.fail[i] = ~0ull;  // el = 0x7FFFFFFF never occurs
  __syncthreads();  // everyone is still here (exited warps count as arrived)

  const uint32_t n_req = a.dp->n;
  const uint32_t n_tiles = (n_req + kTile - 1) / kTile;

  // ======================= producer warps ===================================
  if (warp >= nwarps) {
    const uint32_t pw = warp - nwarps;
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t buf = t & 1, base = t * kTile, tile_end = min(n_req, base + kTile);
      if (t >= 2) bar_sync(kBarFree0 + buf, nall);  // solvers are done with this buffer
      produce_tile(a, sh, comp, buf, base, tile_end, pw, nprod, lane, kBarProducer);
      if (pw == 0 && lane == 0) {
        uint32_t c = 0;
        for (uint32_t w = 0; w < nprod; ++w) c += sh.pcnt[buf][w];
        sh.cnt[buf] = c;
      }
      __threadfence_block();
      bar_arrive(kBarFull0 + buf, nall);
    }
    return;
  }

  // ======================= solver warps ======================================
  const KeyT* __restrict__ codes = reinterpret_cast<const KeyT*>(a.codes);
  const uint32_t sv_begin = a.comp_sv_off[comp];
  const uint32_t n_sv = a.comp_sv_off[comp + 1] - sv_begin;
  const uint8_t* __restrict__ mask_tab = a.envmask + a.comp_mask_off[comp];
  constexpr bool kPacked = KT::kPacked;

  // ---- load my K servants into registers --------------------------------
  KeyT code[kK], nxt[kK];
  uint32_t cur[kK];
  int32_t ver[kK];
  uint32_t mtnz = 0;  // bit j: max_tasks != 0 (needed only to tell Timeout from EnvironmentNotFound)
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    uint32_t li = tid * kK + j;
    code[j] = FULLK;
    nxt[j] = FULLK;
    cur[j] = 0;
    ver[j] = 0;
    if (li < n_sv) {
      uint32_t pos = a.comp_sv[sv_begin + li];
      uint32_t row = a.row_off[pos];
      cur[j] = row;
      code[j] = codes[row];
      // rows always end in a sentinel, so row+1 is readable iff code != FULL
      nxt[j] = code[j] != FULLK ? codes[row + 1] : FULLK;
      if (kPacked) {  // low 3 bits of a packed code are zero in the table: stamp j
        code[j] |= (KeyT)j;
        nxt[j] |= (KeyT)j;
      }
      ver[j] = a.sv.version[pos];
      mtnz |= (a.sv.max_tasks[pos] != 0 ? 1u : 0u) << j;
    }
  }

  uint32_t par = 0;
  // per-class mask words: 0 = candidate, kSelf = the requestor's own servant (last
  // resort), FULL = not eligible.  Recomputed only when the class changes.
  KeyT mk[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) mk[j] = FULLK;
  uint32_t cls_el = kNone, cls_mv = 0, cls_self = kNone, okbits = 0;

  for (uint32_t t = 0; t < n_tiles; ++t) {
    const uint32_t buf = t & 1;
    if (specialised) {
      bar_sync(kBarFull0 + buf, nall);
    } else {
      const uint32_t base = t * kTile, tile_end = min(n_req, base + kTile);
      produce_tile(a, sh, comp, buf, base, tile_end, warp, nwarps, lane, kBarSolver);
      if (tid == 0) {
        uint32_t c = 0;
        for (uint32_t w = 0; w < nwarps; ++w) c += sh.pcnt[buf][w];
        sh.cnt[buf] = c;
      }
      if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
    }
    const uint32_t cnt = sh.cnt[buf];
    const uint4* __restrict__ list = sh.desc[buf];

    // ---- sequential decisions over the compacted list -----------------------
    uint4 ds = cnt ? list[0] : make_uint4(0, 0, 0, kNone);
    uint32_t m = cnt ? mask_tab[ds.y * nthreads + tid] : 0;
    unsigned long long fe = cnt ? sh.fail[ds.y & (kFailSlots - 1)] : ~0ull;
    for (uint32_t d = 0; d < cnt; ++d) {
      const uint32_t q = ds.x, el = ds.y, mv = ds.z;
      uint32_t selfinfo = ds.w;
      const uint32_t mbits = m;
      const unsigned long long fcur = fe;
      // software prefetch of the next descriptor, its membership byte and cache line
      if (d + 1 < cnt) {
        ds = list[d + 1];
        m = mask_tab[ds.y * nthreads + tid];
        fe = sh.fail[ds.y & (kFailSlots - 1)];
      }
      if (fail_hit(fcur, el, mv)) {  // the class failed earlier (possibly in this tile)
        if (tid == 0) a.res[q] = fail_res(fcur);
        continue;
      }
      const bool multi_self = selfinfo != kNone && (selfinfo & 0x80000000u);
      if (el != cls_el || mv != cls_mv || selfinfo != cls_self || multi_self) {
        if (el != cls_el || mv != cls_mv) {
          uint32_t verbits = 0;
#pragma unroll
          for (int j = 0; j < kK; ++j) verbits |= ((uint32_t)ver[j] >= mv ? 1u : 0u) << j;  // int vs uint32, cc:333
          okbits = mbits & verbits;
        }
        cls_el = el; cls_mv = mv; cls_self = selfinfo;
        // several servants share the requestor's IP: "self" is the first of them that
        // is eligible and free right now (find_if over the free list, cc:372-375)
        if (multi_self) {
          uint32_t ip = selfinfo & 0x7FFFFFFFu;
          if (tid == 0) s_selfmin = kNone;
          if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
          for (uint32_t u = a.ip_off[ip]; u < a.ip_off[ip + 1]; ++u) {
            uint32_t p = a.ip_sv[u];
            if (a.sv_comp[p] != comp) continue;
            uint32_t l = a.sv_local[p];
            if ((l / kK) == tid) {
              int jj = l % kK;
              bool free_ok = false;
#pragma unroll
              for (int j = 0; j < kK; ++j) if (j == jj) free_ok = ((okbits >> j) & 1u) && code[j] != FULLK;
              if (free_ok) atomicMin(&s_selfmin, l);
            }
          }
          if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
          selfinfo = s_selfmin;
          if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
        }
        const int selfj = (selfinfo != kNone && (selfinfo / kK) == tid) ? (int)(selfinfo % kK) : -1;
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          mk[j] = ((okbits >> j) & 1u) ? (j == selfj ? KT::kSelf : (KeyT)0) : FULLK;
        }
      }

      // ---- score my K servants: min over (self | tier | code [| j]) ------------
      KeyT key[kK];
#pragma unroll
      for (int j = 0; j < kK; ++j) key[j] = code[j] | mk[j];  // FULL stays FULL
      KeyT b01 = key[0] < key[1] ? key[0] : key[1], b23 = key[2] < key[3] ? key[2] : key[3];
      KeyT b45 = key[4] < key[5] ? key[4] : key[5], b67 = key[6] < key[7] ? key[6] : key[7];
      KeyT b03 = b01 < b23 ? b01 : b23, b47 = b45 < b67 ? b45 : b67;
      const KeyT best = b03 < b47 ? b03 : b47;
      bool found;
      uint32_t gli;
      if constexpr (kPacked) {
        uint32_t gkey;
        warp_argmin((uint32_t)best, tid, gkey, gli);
        if (multi) {
          if (lane == 0) { s_exch_key[par][warp] = gkey; s_exch_li[par][warp] = gli; }
          bar_sync(kBarSolver, nthreads);
          uint32_t k2 = lane < nwarps ? (uint32_t)s_exch_key[par][lane] : 0x1FFFFFFFu;
          uint32_t l2 = lane < nwarps ? s_exch_li[par][lane] : kNone;
          gkey = __reduce_min_sync(0xffffffffu, k2);
          gli = __reduce_min_sync(0xffffffffu, (k2 == gkey) ? l2 : kNone);
          par ^= 1;
        }
        found = gli != kNone;
      } else {
        KeyT gmin = warp_min64(best);
        uint32_t eq = 0;  // lowest local index holding the minimum: first index wins (cc:444)
#pragma unroll
        for (int j = 0; j < kK; ++j) eq |= (key[j] == gmin ? 1u : 0u) << j;
        gli = __reduce_min_sync(0xffffffffu, (eq && gmin != FULLK) ? tid * kK + (__ffs(eq) - 1) : kNone);
        if (multi) {
          if (lane == 0) { s_exch_key[par][warp] = gmin; s_exch_li[par][warp] = gli; }
          bar_sync(kBarSolver, nthreads);
          KeyT k2 = lane < nwarps ? s_exch_key[par][lane] : FULLK;
          uint32_t l2 = lane < nwarps ? s_exch_li[par][lane] : kNone;
          gmin = warp_min64(k2);
          gli = __reduce_min_sync(0xffffffffu, (k2 == gmin) ? l2 : kNone);
          par ^= 1;
        }
        found = gli != kNone;
      }
      if (found) {
        // ---- grant: select-based update of the winner's registers; the only memory
        //      operations are one store and one prefetch nobody waits for
        const uint32_t wj = gli % kK;
        const bool owner = (gli / kK) == tid;
        uint32_t wcur = 0;
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const bool hit = owner && (uint32_t)j == wj;
          cur[j] += hit ? 1u : 0u;
          wcur = hit ? cur[j] : wcur;
          code[j] = hit ? nxt[j] : code[j];
        }
        if (owner) {
          KeyT nn = FULLK;
          bool more = false;
#pragma unroll
          for (int j = 0; j < kK; ++j) more |= ((uint32_t)j == wj) && nxt[j] != FULLK;
          if (more) {
            nn = codes[wcur + 1];
            if (kPacked) nn |= (KeyT)wj;
          }
#pragma unroll
          for (int j = 0; j < kK; ++j) nxt[j] = ((uint32_t)j == wj) ? nn : nxt[j];
          a.res[q] = sv_begin + gli;  // translated to a registry position by k_final_write
        }
      } else {
        // ---- no free servant: Timeout if the class has any eligible servant,
        //      EnvironmentNotFound otherwise (cc:104-118).  Remember the verdict.
        bool elig = (okbits & mtnz) != 0;
        uint32_t any = __any_sync(0xffffffffu, elig) ? 1u : 0u;
        if (multi) {
          if (tid == 0) s_any = 0;
          bar_sync(kBarSolver, nthreads);
          if (lane == 0 && any) atomicOr(&s_any, 1u);
          bar_sync(kBarSolver, nthreads);
          any = s_any;
        }
        uint32_t r = any ? kResTimeout : kResEnvNotFound;
        if (tid == 0) {
          a.res[q] = r;
          sh.fail[el & (kFailSlots - 1)] = pack_fail(el, mv, r);
        }
        if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
        // the prefetched entry for the next decision may predate this write
        if (d + 1 < cnt) fe = sh.fail[ds.y & (kFailSlots - 1)];
      }
    }
    if (specialised) {
      if (t + 2 < n_tiles) bar_arrive(kBarFree0 + buf, nall);
    } else {
      if (multi) bar_sync(kBarSolver, nthreads); else __syncwarp();
    }
  }

  // ---- write back running_tasks / ever_assigned_tasks (cc:123-124) ---------
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    uint32_t li = tid * kK + j;
    if (li < n_sv) {
      uint32_t pos = a.comp_sv[sv_begin + li];
      uint32_t taken = cur[j] - a.row_off[pos];
      if (taken) {
        a.sv.run[pos] += taken;
        a.sv.ever[pos] += taken;
      }
    }
  }
}

}  // namespace yd
