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
//   * each thread owns K=8 consecutive servants of the component: their current
//     slot code (slots.cuh), the code they will have after one more grant, the
//     cursor into the slot table, and the daemon version;
//   * eligibility (ContainsEnvironmentSlow && max_tasks != 0 && version >=
//     min_version, cc:316-344) is one byte of precomputed digest-membership bits
//     per (digest, thread) plus K integer compares;
//   * the pick (self-exclusion :372-379, dedicated tier :399-410, min
//     utilisation, first index wins :417-451) is a min over
//     (self|tier|code, local index): two REDUX.MIN per warp, one 8-byte shared
//     memory word per warp and one named barrier per decision for multi-warp
//     components;
//   * only the winner touches memory afterwards (one L2 load of its next code,
//     off the dependency chain unless it wins twice in a row).
//
// Requests are streamed from HBM in tiles of 1024 with coalesced 8-byte loads,
// filtered to this component and compacted (order-preserving) into shared
// memory; decisions then read 16-byte descriptors by broadcast LDS.
//
// Monotonicity used for the Timeout short-cut: inside one batch running_tasks
// only grows, so once a (digest, min_version) class found no free servant every
// later request of that class times out too (EnvironmentNotFound is static).
#pragma once
#include "common.cuh"

namespace yd {

constexpr int kK = 8;          // servants per thread
constexpr int kTile = 1024;    // requests per tile
constexpr int kFailSlots = 128;

struct SolveArgs {
  const yd_task_req* reqs;
  uint32_t n;
  uint32_t* res;
  // topology (rebuilt on the host when the servant set / digests change)
  const uint32_t* env_comp;   // [n_envs] component of a digest id, kNone if nobody holds it
  const uint32_t* env_local;  // [n_envs] index of the digest inside its component
  uint32_t n_envs;
  const uint32_t* comp_sv_off;    // [C+1] offsets into comp_sv
  const uint32_t* comp_sv;        // registry positions, ascending inside a component
  const uint32_t* comp_mask_off;  // [C] byte offset of the component's membership table
  const uint32_t* comp_nwarps;    // [C]
  const uint8_t* envmask;         // per component [n_env_local][threads] bytes, bit j = servant tid*8+j holds it
  const uint32_t* sv_comp;        // [S]
  const uint32_t* sv_local;       // [S] index inside its component
  const uint32_t* ip_off;         // CSR: interned requestor-ip id -> registry positions whose
  const uint32_t* ip_sv;          //      observed_location matches it (IsNetworkAddressEqual, cc:66-69)
  uint32_t n_ips;
  ServantArrays sv;
  const uint32_t* row_off;  // slot table
  const void* codes;        // uint32_t[] or unsigned long long[] (wide)
};

template <typename KeyT> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {
  static constexpr uint32_t kFullKey = 0xFFFFFFFFu;
  static constexpr uint32_t kSelf = 0x80000000u;
  static __device__ __forceinline__ uint32_t warp_min(uint32_t v) {
    return __reduce_min_sync(0xffffffffu, v);
  }
};
template <> struct KeyTraits<unsigned long long> {
  static constexpr unsigned long long kFullKey = ~0ull;
  static constexpr unsigned long long kSelf = 1ull << 63;
  static __device__ __forceinline__ unsigned long long warp_min(unsigned long long v) {
    uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xFFFFFFFFu);
    return ((unsigned long long)mh << 32) | ml;
  }
};

__device__ __forceinline__ void named_bar_sync(uint32_t nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <typename KeyT, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads, 1) k_solve_rowscan(SolveArgs a) {
  using KT = KeyTraits<KeyT>;
  constexpr KeyT FULLK = KT::kFullKey;
  const uint32_t comp = blockIdx.x;
  const uint32_t nwarps = a.comp_nwarps[comp];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp >= nwarps) return;  // surplus warps of small components leave at once
  const uint32_t nthreads = nwarps * 32;
  const bool multi = nwarps > 1;

  __shared__ uint4 s_desc[kTile];  // {request index, local digest, min_version, self info}
  __shared__ unsigned long long s_exch[2][32];
  __shared__ KeyT s_exch_key[2][32];
  __shared__ uint32_t s_wcnt[32];
  __shared__ uint32_t s_fail_env[kFailSlots], s_fail_mv[kFailSlots], s_fail_res[kFailSlots];
  __shared__ uint32_t s_selfmin, s_any;

  const KeyT* __restrict__ codes = reinterpret_cast<const KeyT*>(a.codes);
  const uint32_t sv_begin = a.comp_sv_off[comp];
  const uint32_t n_sv = a.comp_sv_off[comp + 1] - sv_begin;
  const uint8_t* __restrict__ mask_tab = a.envmask + a.comp_mask_off[comp];

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
      ver[j] = a.sv.version[pos];
      mtnz |= (a.sv.max_tasks[pos] != 0 ? 1u : 0u) << j;
    }
  }
  for (uint32_t i = tid; i < kFailSlots; i += nthreads) s_fail_env[i] = kNone;
  if (multi) named_bar_sync(nthreads); else __syncwarp();

  uint32_t par = 0;
  for (uint32_t base = 0; base < a.n; base += kTile) {
    const uint32_t tile_end = min(a.n, base + kTile);
    // ---- phase A: filter + order-preserving compaction of this tile --------
    // Warp w scans the contiguous chunk [c0, c1) twice: count, then write.
    const uint32_t chunk = ((kTile + nwarps - 1) / nwarps + 31) & ~31u;
    const uint32_t c0 = min(tile_end, base + warp * chunk), c1 = min(tile_end, c0 + chunk);
    uint32_t my_count = 0;
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t woff = 0;
      if (pass == 1) {
        for (uint32_t w = 0; w < warp; ++w) woff += s_wcnt[w];
      }
      for (uint32_t q0 = c0; q0 < c1; q0 += 32) {
        uint32_t q = q0 + lane;
        bool mine = false;
        uint32_t env = 0, mv = 0, ip = 0;
        if (q < c1) {
          const uint2* rp = reinterpret_cast<const uint2*>(a.reqs + q);
          uint2 w0 = __ldg(rp);      // env_id, min_version
          uint2 w1 = __ldg(rp + 1);  // requestor_ip, flags
          env = w0.x; mv = w0.y; ip = w1.x;
          mine = env < a.n_envs && __ldg(a.env_comp + env) == comp;
        }
        uint32_t el = 0;
        if (mine) {
          el = __ldg(a.env_local + env);
          // classes already known to fail are answered right here, in parallel
          uint32_t fs = el & (kFailSlots - 1);
          if (s_fail_env[fs] == el && s_fail_mv[fs] == mv) {
            if (pass == 1) a.res[q] = s_fail_res[fs];
            mine = false;
          }
        }
        uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (pass == 0) {
          my_count += __popc(bal);
        } else if (mine) {
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
          s_desc[woff + __popc(bal & ((1u << lane) - 1))] = make_uint4(q, el, mv, selfinfo);
        }
        if (pass == 1) woff += __popc(bal);
      }
      if (pass == 0) {
        if (lane == 0) s_wcnt[warp] = my_count;
        if (multi) named_bar_sync(nthreads); else __syncwarp();
      }
    }
    uint32_t cnt = 0;
    for (uint32_t w = 0; w < nwarps; ++w) cnt += s_wcnt[w];
    if (multi) named_bar_sync(nthreads); else __syncwarp();

    // ---- phase B: sequential decisions over the compacted list ------------
    uint4 ds = cnt ? s_desc[0] : make_uint4(0, 0, 0, kNone);
    uint32_t m = cnt ? mask_tab[ds.y * nthreads + tid] : 0;
    for (uint32_t d = 0; d < cnt; ++d) {
      const uint32_t q = ds.x, el = ds.y, mv = ds.z;
      uint32_t selfinfo = ds.w;
      const uint32_t mbits = m;
      // software prefetch of the next descriptor and its membership byte
      if (d + 1 < cnt) {
        ds = s_desc[d + 1];
        m = mask_tab[ds.y * nthreads + tid];
      }
      // a class that failed earlier in this tile
      {
        uint32_t fs = el & (kFailSlots - 1);
        if (s_fail_env[fs] == el && s_fail_mv[fs] == mv) {
          if (tid == 0) a.res[q] = s_fail_res[fs];
          continue;
        }
      }
      uint32_t okbits = 0;
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        bool ok = ((mbits >> j) & 1u) && ((uint32_t)ver[j] >= mv);  // int vs uint32 compare, cc:333
        okbits |= (ok ? 1u : 0u) << j;
      }
      // several servants share the requestor's IP: "self" is the first of them that
      // is eligible and free right now (find_if over the free list, cc:372-375)
      if (selfinfo != kNone && (selfinfo & 0x80000000u)) {
        uint32_t ip = selfinfo & 0x7FFFFFFFu;
        if (tid == 0) s_selfmin = kNone;
        if (multi) named_bar_sync(nthreads); else __syncwarp();
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
        if (multi) named_bar_sync(nthreads); else __syncwarp();
        selfinfo = s_selfmin;
        if (multi) named_bar_sync(nthreads); else __syncwarp();
      }
      const int selfj = (selfinfo != kNone && (selfinfo / kK) == tid) ? (int)(selfinfo % kK) : -1;

      // ---- score my K servants: min over (self | tier | code), lowest j wins ties
      KeyT best = FULLK;
      int bj = 0;
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        KeyT key = code[j] | (j == selfj ? KT::kSelf : (KeyT)0);  // FULL stays FULL
        key = ((okbits >> j) & 1u) ? key : FULLK;
        if (key < best) { best = key; bj = j; }
      }
      KeyT gmin = KT::warp_min(best);
      uint32_t gli = __reduce_min_sync(0xffffffffu, (best == gmin && gmin != FULLK) ? tid * kK + bj : kNone);
      if (multi) {
        if (lane == 0) { s_exch_key[par][warp] = gmin; s_exch[par][warp] = gli; }
        named_bar_sync(nthreads);
        KeyT k2 = lane < nwarps ? s_exch_key[par][lane] : FULLK;
        uint32_t l2 = lane < nwarps ? (uint32_t)s_exch[par][lane] : kNone;
        gmin = KT::warp_min(k2);
        gli = __reduce_min_sync(0xffffffffu, (k2 == gmin) ? l2 : kNone);
        par ^= 1;
      }
      if (gmin != FULLK) {
        // ---- grant: only the owner of the winning servant does any work ----
        if ((gli / kK) == tid) {
          const int wj = gli % kK;
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            if (j == wj) {
              code[j] = nxt[j];
              cur[j] += 1;
              nxt[j] = nxt[j] != FULLK ? codes[cur[j] + 1] : FULLK;
            }
          }
          a.res[q] = a.comp_sv[sv_begin + gli];
        }
      } else {
        // ---- no free servant: Timeout if the class has any eligible servant,
        //      EnvironmentNotFound otherwise (cc:104-118).  Remember the verdict.
        bool elig = (okbits & mtnz) != 0;
        uint32_t any = __any_sync(0xffffffffu, elig) ? 1u : 0u;
        if (multi) {
          if (tid == 0) s_any = 0;
          named_bar_sync(nthreads);
          if (lane == 0 && any) atomicOr(&s_any, 1u);
          named_bar_sync(nthreads);
          any = s_any;
        }
        uint32_t r = any ? kResTimeout : kResEnvNotFound;
        if (tid == 0) {
          a.res[q] = r;
          uint32_t fs = el & (kFailSlots - 1);
          s_fail_env[fs] = el; s_fail_mv[fs] = mv; s_fail_res[fs] = r;
        }
        if (multi) named_bar_sync(nthreads); else __syncwarp();
      }
    }
    if (multi) named_bar_sync(nthreads); else __syncwarp();
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
