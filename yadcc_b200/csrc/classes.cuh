// classes.cuh -- request classes and their sorted slot lists (slot-stream solver).
//
// A *class* is what UnsafeEnumerateEligibleServants keys on: (compiler digest,
// min_version) (task_dispatcher.cc:316-344).  All requests of a class see the same
// eligible servants, hence the same candidate slots in the same order; the only
// per-request difference is which servant is "self" (cc:372-379).
//
//   k_cls_insert   every request inserts its (digest id, min_version) into a small
//                  open-addressing table in HBM (atomicCAS on 8-byte keys) and flags
//                  its component if the requestor's IP matches a servant there
//   k_cls_number   one block numbers the occupied table slots in slot order
//                  (deterministic ids) and records digest / min_version / component
//   k_cls_elig     per class: number of eligible servants (0 => EnvironmentNotFound)
//   k_slot_decode  sorted slot -> (registry position, running_tasks value, component)
//   k_list_count / k_list_fill   per class, order-preserving filter of the sorted
//                  slots by class eligibility -> L[c] = (servant local index, r)
#pragma once
#include "common.cuh"

namespace yd {

constexpr uint32_t kClsTableSize = 2048;  // power of two
constexpr uint32_t kMaxClasses = 256;
constexpr unsigned long long kClsEmpty = ~0ull;

struct ClassTable {
  unsigned long long* keys;  // [kClsTableSize] digest id << 32 | min_version, or kClsEmpty
  uint32_t* slot_cls;        // [kClsTableSize] class id of the slot
  uint32_t* meta;            // [0] number of classes, [1] overflow flag (1: too many classes / table full / list
                             //     too long -> row-scan solver; 2: more classes than cls_bound -> retry with a bigger bound)
  uint32_t cls_bound;        // classes the per-class grids / tables of this solve are sized for (<= kMaxClasses)
  uint32_t* cls_env;         // [kMaxClasses]
  uint32_t* cls_mv;
  uint32_t* cls_comp;
  uint32_t* cls_nelig;       // eligible servants (max_tasks != 0, digest, version)
  uint32_t* cls_count;       // requests of the class in this batch
  uint32_t* comp_flags;      // [C] bit 0: some request's "self" servant lives in the component
  uint32_t* comp_ncls;       // [C] classes in the component
};

struct TopoView {  // the parts of the topology the class kernels need
  const uint32_t* env_comp;
  uint32_t n_envs;
  const uint32_t* sv_comp;
  const uint32_t* sv_local;
  const uint32_t* ip_off;
  const uint32_t* ip_sv;
  uint32_t n_ips;
  const uint32_t* sv_env_off;  // CSR: registry position -> digest ids it holds
  const uint32_t* sv_envs;
  const uint32_t* comp_sv_off;
  const uint32_t* comp_sv;
};

__device__ __forceinline__ uint32_t cls_hash(unsigned long long key) {
  key ^= key >> 33;
  key *= 0xff51afd7ed558ccdULL;
  key ^= key >> 29;
  return (uint32_t)key & (kClsTableSize - 1);
}

// Returns the table slot holding `key`, or kNone if absent (lookup only).
__device__ __forceinline__ uint32_t cls_find(const unsigned long long* __restrict__ keys, unsigned long long key) {
  uint32_t s = cls_hash(key);
  for (uint32_t probe = 0; probe < kClsTableSize; ++probe) {
    unsigned long long k = keys[s];
    if (k == key) return s;
    if (k == kClsEmpty) return kNone;
    s = (s + 1) & (kClsTableSize - 1);
  }
  return kNone;
}

__device__ __forceinline__ bool servant_has_env(const TopoView& t, uint32_t pos, uint32_t env) {
  for (uint32_t u = t.sv_env_off[pos], e = t.sv_env_off[pos + 1]; u < e; ++u) {
    if (t.sv_envs[u] == env) return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_cls_insert(const yd_task_req* __restrict__ reqs,
                                                    const DynParams* __restrict__ dp, TopoView t, ClassTable ct) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= dp->n) return;
  const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(reqs + q));
  const uint32_t env = w0.x, mv = w0.y;
  if (env >= t.n_envs) return;
  const uint32_t comp = t.env_comp[env];
  if (comp == kNone) return;
  const unsigned long long key = ((unsigned long long)env << 32) | mv;
  // one lane per distinct key of the warp does the probing (the batch has few classes)
  const uint32_t active = __activemask();
  const uint32_t peers = __match_any_sync(active, key);
  if ((uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31)) {
    uint32_t s = cls_hash(key);
    bool done = false;
    for (uint32_t probe = 0; probe < kClsTableSize && !done; ++probe) {
      unsigned long long k = ct.keys[s];
      if (k == kClsEmpty) k = atomicCAS(&ct.keys[s], kClsEmpty, key);
      if (k == key || k == kClsEmpty) done = true;
      else s = (s + 1) & (kClsTableSize - 1);
    }
    if (!done) atomicExch(&ct.meta[1], 1u);  // table full -> caller falls back to the row-scan solver
  }
  // does the requestor's IP belong to a servant of this component?
  const uint32_t ip = __ldg(reinterpret_cast<const uint2*>(reqs + q) + 1).x;
  if (ip < t.n_ips) {
    for (uint32_t u = t.ip_off[ip], e = t.ip_off[ip + 1]; u < e; ++u) {
      if (t.sv_comp[t.ip_sv[u]] == comp) {
        if (!(ct.comp_flags[comp] & 1u)) atomicOr(&ct.comp_flags[comp], 1u);
        break;
      }
    }
  }
}

// One block: deterministic class ids = rank of the occupied slot.
__global__ void __launch_bounds__(1024) k_cls_number(TopoView t, ClassTable ct) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < kClsTableSize; base += 1024) {
    const uint32_t s = base + tid;
    const unsigned long long k = ct.keys[s];
    const uint32_t v = k != kClsEmpty ? 1u : 0u;
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
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t id = carry_s + (warp ? warp_sums[warp - 1] : 0) + x - v;
    if (v) {
      if (id < kMaxClasses) {
        ct.slot_cls[s] = id;
        const uint32_t env = (uint32_t)(k >> 32);
        const uint32_t comp = t.env_comp[env];
        ct.cls_env[id] = env;
        ct.cls_mv[id] = (uint32_t)k;
        ct.cls_comp[id] = comp;
        ct.cls_nelig[id] = 0;
        ct.cls_count[id] = 0;
        atomicAdd(&ct.comp_ncls[comp], 1u);
      } else {
        ct.slot_cls[s] = kNone;
        ct.meta[1] = 1;
      }
      if (id >= ct.cls_bound && id < kMaxClasses) atomicMax(&ct.meta[1], 2u);  // 1 (hard overflow) wins below
    }
    __syncthreads();
    if (tid == 1023) carry_s += warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) {
    ct.meta[0] = carry_s < kMaxClasses ? carry_s : kMaxClasses;
    if (carry_s > kMaxClasses) ct.meta[1] = 1;
  }
}

// grid.y = class; threads stride over the servants of the class's component.
__global__ void __launch_bounds__(256) k_cls_elig(TopoView t, ClassTable ct, ServantArrays sv) {
  const uint32_t c = blockIdx.y;
  if (c >= ct.meta[0]) return;
  const uint32_t comp = ct.cls_comp[c], env = ct.cls_env[c], mv = ct.cls_mv[c];
  const uint32_t b = t.comp_sv_off[comp], e = t.comp_sv_off[comp + 1];
  int mine = 0;
  for (uint32_t i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
    const uint32_t pos = t.comp_sv[i];
    mine += (sv.max_tasks[pos] != 0 && (uint32_t)sv.version[pos] >= mv && servant_has_env(t, pos, env)) ? 1 : 0;
  }
  __shared__ uint32_t s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  if (mine) atomicAdd(&s_sum, (uint32_t)mine);
  __syncthreads();
  if (threadIdx.x == 0 && s_sum) atomicAdd(&ct.cls_nelig[c], s_sum);
}

// sorted slot i -> owner position, running_tasks value of the slot, component.
__global__ void __launch_bounds__(256) k_slot_decode(const uint32_t* __restrict__ sorted_orig,
                                                     const unsigned long long* __restrict__ m_ptr,
                                                     const uint32_t* __restrict__ slot_owner,
                                                     const uint32_t* __restrict__ row_off,
                                                     const uint32_t* __restrict__ row_len,
                                                     const uint32_t* __restrict__ run,
                                                     const uint32_t* __restrict__ sv_comp,
                                                     uint32_t* __restrict__ s_pos, uint32_t* __restrict__ s_r,
                                                     uint32_t* __restrict__ s_comp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)*m_ptr) return;
  const uint32_t orig = sorted_orig[i];
  const uint32_t pos = slot_owner[orig];
  const uint32_t k = orig - row_off[pos];
  s_pos[i] = pos;
  s_r[i] = run[pos] + k;
  s_comp[i] = (k < row_len[pos]) ? sv_comp[pos] : kNone;  // the row's sentinel is nobody's slot
}

constexpr int kListTile = 1024;

__device__ __forceinline__ bool slot_in_class(const TopoView& t, const ClassTable& ct, const ServantArrays& sv,
                                              uint32_t c_comp, uint32_t c_env, uint32_t c_mv, uint32_t pos,
                                              uint32_t comp) {
  return comp == c_comp && (uint32_t)sv.version[pos] >= c_mv && servant_has_env(t, pos, c_env);
}

// grid = (tiles, classes): per (class, tile) number of member slots.
__global__ void __launch_bounds__(kListTile) k_list_count(const unsigned long long* __restrict__ m_ptr,
                                                          const uint32_t* __restrict__ s_pos,
                                                          const uint32_t* __restrict__ s_comp, TopoView t,
                                                          ClassTable ct, ServantArrays sv, uint32_t n_tiles,
                                                          uint32_t* __restrict__ counts) {
  const uint32_t c = blockIdx.y;
  if (c >= ct.meta[0]) return;
  const uint32_t m = (uint32_t)*m_ptr;
  const uint32_t i = blockIdx.x * kListTile + threadIdx.x;
  bool in = false;
  if (i < m) in = slot_in_class(t, ct, sv, ct.cls_comp[c], ct.cls_env[c], ct.cls_mv[c], s_pos[i], s_comp[i]);
  const int cnt = __syncthreads_count(in);
  if (threadIdx.x == 0) counts[c * n_tiles + blockIdx.x] = (uint32_t)cnt;
}

// counts[] has been exclusive-scanned over (class-major, tile-minor); counts[C*n_tiles] = total.
__global__ void __launch_bounds__(kListTile) k_list_fill(const unsigned long long* __restrict__ m_ptr,
                                                         const uint32_t* __restrict__ s_pos,
                                                         const uint32_t* __restrict__ s_r,
                                                         const uint32_t* __restrict__ s_comp, TopoView t,
                                                         ClassTable ct, ServantArrays sv, uint32_t n_tiles,
                                                         const uint32_t* __restrict__ offs,
                                                         uint2* __restrict__ list, uint32_t list_cap) {
  __shared__ uint32_t warp_cnt[32];
  const uint32_t c = blockIdx.y;
  if (c >= ct.meta[0]) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t m = (uint32_t)*m_ptr;
  const uint32_t i = blockIdx.x * kListTile + tid;
  bool in = false;
  uint32_t pos = 0;
  if (i < m) {
    pos = s_pos[i];
    in = slot_in_class(t, ct, sv, ct.cls_comp[c], ct.cls_env[c], ct.cls_mv[c], pos, s_comp[i]);
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, in);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[w];
  before += __popc(bal & ((1u << lane) - 1));
  if (in) {
    const uint32_t dst = offs[c * n_tiles + blockIdx.x] + before;
    if (dst < list_cap) list[dst] = make_uint2(t.sv_local[pos], s_r[i]);
    else ct.meta[1] = 1;  // more (class, slot) pairs than provisioned: the host reruns with solver 1
  }
}

}  // namespace yd
