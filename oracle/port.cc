// port.cc -- TEST INFRASTRUCTURE: a plain, sequential CPU restatement of the
// reference scheduler hot path, exported through the C ABI of
// include/ydsched.h as `oracle/libydoracle.so`.
//
// It is the checker the CUDA backend is compared against when the verbatim
// reference build (oracle/_ref/libydref.so) is not at hand (e.g. on the GPU
// box, where /root/reference does not exist), and it is itself pinned against
// that verbatim build and against the reference's golden tests
// (tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py).
//
// Only tests/, bench.py's cpu_baseline / reference legs and
// __graft_entry__.smoke() may load this library.  Nothing under yadcc_b200/
// does.
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference).  The code is written from the behaviour described in
// SURVEY.md section 8(a), not transliterated: registries are flat vectors with
// integer ids, one pass computes eligibility, freeness and the pick.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "ydsched.h"

namespace {

// yadcc/common/parse_size.cc:25-45.  Optional K/M/G/B suffix, decimal body.
bool ParseSize(const std::string& text, std::uint64_t* out) {
  if (text.empty()) return false;
  std::uint64_t scale = 1;
  std::size_t len = text.size();
  switch (text.back()) {
    case 'G': scale = 1ull << 30; --len; break;
    case 'M': scale = 1ull << 20; --len; break;
    case 'K': scale = 1ull << 10; --len; break;
    case 'B': --len; break;
    default: break;
  }
  if (len == 0) return false;
  std::uint64_t v = 0;
  for (std::size_t i = 0; i != len; ++i) {
    char c = text[i];
    if (c < '0' || c > '9') return false;
    std::uint64_t nv = v * 10 + static_cast<std::uint64_t>(c - '0');
    if (nv / 10 != v) return false;  // overflow -> flare::TryParse fails
    v = nv;
  }
  *out = v * scale;
  return true;
}

struct ServantRec {
  // ServantPersonality, yadcc/scheduler/task_dispatcher.h:80-116.
  std::int32_t version = 0;
  std::string observed_location, reported_location;
  std::vector<std::string> environments;
  std::uint64_t num_processors = 0, current_load = 0;
  std::uint64_t total_memory = 0, memory_available = 0, max_tasks = 0;
  std::int32_t priority = 0, reason = 0;
  // ServantDesc, task_dispatcher.h:184-193.
  std::int64_t discovered_at = 0, expires_at = 0;
  std::uint64_t running_tasks = 0, ever_assigned_tasks = 0;
};

struct TaskRec {
  // TaskDesc, task_dispatcher.h:199-215.  `servant` plays the part of the
  // RefPtr<ServantDesc>: it keeps a detached servant record alive so that
  // decrements on it stay harmless (task_dispatcher.cc:177).
  std::shared_ptr<ServantRec> servant;
  std::int64_t started_at = 0, expires_at = 0;
  bool is_prefetch = false, zombie = false;
};

struct RunningTaskRec {
  std::uint64_t servant_task_id = 0, task_grant_id = 0;
  std::string servant_location, task_digest;
};

// task_dispatcher.cc:66-69.
bool IsNetworkAddressEqual(const std::string& ip_port, const std::string& ip) {
  return ip_port.size() > ip.size() && ip_port[ip.size()] == ':' &&
         ip_port.compare(0, ip.size(), ip) == 0;
}

}  // namespace

struct yd_sched {
  std::uint64_t min_memory_for_new_task = 0;
  std::vector<std::shared_ptr<ServantRec>> servants;  // discovery order == tie-break order
  std::unordered_map<std::uint64_t, TaskRec> tasks;
  std::uint64_t next_task_id = 0;  // task_dispatcher.h:218
  std::uint64_t id_stride = 1, id_offset = 0;
  // RunningTaskBookkeeper, running_task_bookkeeper.h:41-42.  Same container and
  // same operation sequence as the reference, so iteration order matches too.
  std::unordered_map<std::string, std::vector<RunningTaskRec>> running_tasks;

  std::vector<std::string> envs, ips;
  std::unordered_map<std::string, std::uint32_t> env_ids, ip_ids;
  std::vector<RunningTaskRec> running_cache;
  std::vector<const char*> personality_envs;
  struct Keeper {  // RunningTaskKeeper state (in-flight task index)
    std::vector<RunningTaskRec> snapshot;
    std::unordered_map<std::string, uint32_t> by_digest;  // digest -> winning snapshot index
  } keeper;
  yd_solve_stats stats{};
  bool have_stats = false;

  // GetCapacityAvailable, task_dispatcher.cc:283-313.
  std::uint64_t CapacityAvailable(const ServantRec& s) const {
    if (s.total_memory != 0 && s.memory_available < min_memory_for_new_task) {
      return s.running_tasks;  // low memory: pretend it is exactly full
    }
    // size_t subtraction reinterpreted as int64, clamped at 0 (cc:307-310).
    std::int64_t foreign =
        std::max<std::int64_t>(static_cast<std::int64_t>(s.current_load - s.running_tasks), 0);
    std::uint64_t avail = static_cast<std::uint64_t>(std::max<std::int64_t>(
        static_cast<std::int64_t>(s.num_processors - static_cast<std::uint64_t>(foreign)), 0));
    return std::min(s.max_tasks, avail);
  }

  // UnsafeFreeTasks, task_dispatcher.cc:172-188: an unknown id ends the call
  // (remaining ids are skipped).
  void FreeTasks(const std::vector<std::uint64_t>& ids) {
    for (auto id : ids) {
      auto it = tasks.find(id);
      if (it == tasks.end()) return;
      --it->second.servant->running_tasks;
      tasks.erase(it);
    }
  }

  // One WaitForStartingNewTask with timeout == now: task_dispatcher.cc:93-140
  // plus UnsafeEnumerateEligibleServants (:316-344), UnsafeEnumerateFreeServants
  // (:346-360), UnsafePickServantFor (:362-397), UnsafeTryPickServantFor
  // (:417-451) folded into one pass over the registry.
  void Decide(std::int64_t now, const std::string& requestor_ip, std::uint32_t min_version,
              const std::string& digest, std::int64_t expires_in, bool prefetch, yd_grant* out) {
    bool any_eligible = false, any_free = false;
    int self = -1;                    // first free servant on the requestor's IP (:372-379)
    int best_dedicated = -1, best_any = -1;
    double u_dedicated = 0, u_any = 0;
    for (int i = 0; i != static_cast<int>(servants.size()); ++i) {
      const ServantRec& s = *servants[i];
      bool has_env = false;  // ContainsEnvironmentSlow, :55-63
      for (auto&& e : s.environments) has_env |= (e == digest);
      if (!has_env || s.max_tasks == 0) continue;
      // `int < uint32_t` compares as unsigned (:333).
      if (static_cast<std::uint32_t>(s.version) < min_version) continue;
      any_eligible = true;
      std::uint64_t cap = CapacityAvailable(s);
      if (s.running_tasks >= cap) continue;  // :353
      any_free = true;
      if (self < 0 && IsNetworkAddressEqual(s.observed_location, requestor_ip)) {
        self = i;  // removed from the candidate list; used only as a last resort
        continue;
      }
      double u = static_cast<double>(s.running_tasks) / static_cast<double>(cap);  // :440-441
      // Dedicated tier (:399-410): priority DEDICATED and below 50% of its cores.
      if (s.priority == YD_PRIORITY_DEDICATED && s.running_tasks * 2 < s.num_processors) {
        if (best_dedicated < 0 || u < u_dedicated) { best_dedicated = i; u_dedicated = u; }
      }
      if (best_any < 0 || u < u_any) { best_any = i; u_any = u; }  // first minimum wins, :444
    }
    out->task_id = 0;
    out->servant_index = YD_NO_SERVANT;
    if (!any_eligible) { out->status = YD_STATUS_ENVIRONMENT_NOT_FOUND; return; }  // :105-108
    if (!any_free) { out->status = YD_STATUS_TIMEOUT; return; }                    // :116-118
    int pick = best_dedicated >= 0 ? best_dedicated : (best_any >= 0 ? best_any : self);
    if (pick < 0) std::abort();  // FLARE_CHECK(self), :394
    auto& s = servants[pick];
    ++s->running_tasks;  // :123-124
    ++s->ever_assigned_tasks;
    // :127; sharded deployments hand out local * stride + offset (yd_config.id_stride / id_offset):
    // the registry is keyed by the external id, so other shards' ids are simply unknown here
    std::uint64_t id = (next_task_id++) * id_stride + id_offset;
    TaskRec& t = tasks[id];
    t.servant = s;
    t.started_at = now;
    t.expires_at = now + expires_in;
    t.is_prefetch = prefetch;
    out->task_id = id;
    out->servant_index = static_cast<std::uint32_t>(pick);
    out->status = YD_STATUS_GRANTED;
  }
};

extern "C" {

const char* yd_backend_name(void) { return "oracle-port"; }

int yd_parse_size(const char* text, uint64_t* out_bytes) {
  if (!text) return 0;
  return ParseSize(text, out_bytes) ? 1 : 0;
}

yd_sched* yd_create(const yd_config* cfg) {
  if (!cfg || cfg->abi_version != YD_ABI_VERSION) return nullptr;
  auto* s = new yd_sched;
  const char* mm = cfg->servant_min_memory_for_accepting_new_task;
  if (!ParseSize(mm ? mm : "10G", &s->min_memory_for_new_task)) {  // cc:83-87
    delete s;
    return nullptr;
  }
  s->ips.emplace_back();  // id 0 == YD_IP_NONE == ""
  s->ip_ids.emplace("", 0);
  if (cfg->id_stride > 1) {
    if (cfg->id_offset >= cfg->id_stride) { delete s; return nullptr; }
    s->id_stride = cfg->id_stride;
    s->id_offset = cfg->id_offset;
  }
  return s;
}

void yd_destroy(yd_sched* s) { delete s; }

uint32_t yd_intern_env(yd_sched* s, const char* digest, size_t len) {
  std::string k(digest, len);
  auto it = s->env_ids.find(k);
  if (it != s->env_ids.end()) return it->second;
  auto id = static_cast<std::uint32_t>(s->envs.size());
  s->envs.push_back(k);
  s->env_ids.emplace(std::move(k), id);
  return id;
}

uint32_t yd_intern_ip(yd_sched* s, const char* ip, size_t len) {
  std::string k(ip, len);
  auto it = s->ip_ids.find(k);
  if (it != s->ip_ids.end()) return it->second;
  auto id = static_cast<std::uint32_t>(s->ips.size());
  s->ips.push_back(k);
  s->ip_ids.emplace(std::move(k), id);
  return id;
}

// KeepServantAlive, task_dispatcher.cc:190-220.
void yd_keep_servant_alive(yd_sched* s, int64_t now_ns, const yd_servant* sv,
                           int64_t expires_in_ns) {
  std::shared_ptr<ServantRec> rec;
  for (auto&& e : s->servants) {
    if (e->observed_location == sv->observed_location) { rec = e; break; }
  }
  if (!rec) {
    rec = s->servants.emplace_back(std::make_shared<ServantRec>());
    rec->discovered_at = now_ns;
    rec->running_tasks = 0;
  }
  // The whole personality is overwritten; running_tasks / ever_assigned /
  // discovered_at survive (cc:198-201).
  rec->version = sv->version;
  rec->observed_location = sv->observed_location;
  rec->reported_location = sv->reported_location;
  rec->environments.clear();
  for (std::uint32_t i = 0; i != sv->num_envs; ++i) rec->environments.emplace_back(sv->env_digests[i]);
  rec->num_processors = sv->num_processors;
  rec->current_load = sv->current_load;
  rec->total_memory = sv->total_memory_in_bytes;
  rec->memory_available = sv->memory_available_in_bytes;
  rec->max_tasks = sv->max_tasks;
  rec->priority = sv->priority;
  rec->reason = sv->not_accepting_task_reason;
  rec->expires_at = now_ns + expires_in_ns;
}

// NotifyServantRunningTasks, task_dispatcher.cc:222-277.
size_t yd_notify_servant_running_tasks(yd_sched* s, const char* servant_location,
                                       const yd_running_task* tasks, size_t n,
                                       uint64_t* unknown_out) {
  ServantRec* servant = nullptr;
  for (auto&& e : s->servants) {
    if (e->observed_location == servant_location) { servant = e.get(); break; }
  }
  if (!servant) {  // servant itself expired: every reported id is unknown (:243-245)
    for (size_t i = 0; i != n; ++i) unknown_out[i] = tasks[i].task_grant_id;
    return n;
  }
  std::unordered_set<std::uint64_t> reported;
  for (size_t i = 0; i != n; ++i) reported.insert(tasks[i].task_grant_id);
  // UnsafeSweepZombiesOf, cc:453-476: zombies of this servant it no longer reports.
  std::vector<std::uint64_t> sweeping;
  for (auto&& [id, t] : s->tasks) {
    if (t.servant.get() == servant && t.zombie && !reported.count(id)) sweeping.push_back(id);
  }
  s->FreeTasks(sweeping);
  // Reported ids that are not live, non-zombie grants on this servant (:257-273).
  size_t k = 0;
  std::vector<RunningTaskRec> kept;
  for (size_t i = 0; i != n; ++i) {
    auto it = s->tasks.find(tasks[i].task_grant_id);
    bool permitted = it != s->tasks.end() && it->second.servant.get() == servant && !it->second.zombie;
    if (!permitted) {
      unknown_out[k++] = tasks[i].task_grant_id;
    } else {
      kept.push_back(RunningTaskRec{tasks[i].servant_task_id, tasks[i].task_grant_id,
                                    tasks[i].servant_location ? tasks[i].servant_location : "",
                                    tasks[i].task_digest ? tasks[i].task_digest : ""});
    }
  }
  // RunningTaskBookkeeper::SetServantRunningTasks, running_task_bookkeeper.cc:24-29.
  s->running_tasks.erase(servant_location);
  s->running_tasks.emplace(servant_location, std::move(kept));
  return k;
}

// RunningTaskBookkeeper::GetRunningTasks, running_task_bookkeeper.cc:36-43:
// each servant's list is inserted at the FRONT of the result.
size_t yd_get_running_tasks(yd_sched* s, yd_running_task* out, size_t cap) {
  s->running_cache.clear();
  for (auto&& [k, v] : s->running_tasks) {
    s->running_cache.insert(s->running_cache.begin(), v.begin(), v.end());
  }
  for (size_t i = 0; i < s->running_cache.size() && i < cap; ++i) {
    auto&& t = s->running_cache[i];
    out[i] = yd_running_task{t.servant_task_id, t.task_grant_id, t.servant_location.c_str(),
                             t.task_digest.c_str()};
  }
  return s->running_cache.size();
}

// OnExpirationTimer, task_dispatcher.cc:498-536.
void yd_on_expiration_timer(yd_sched* s, int64_t now_ns) {
  // 1. Drop servants whose lease ran out (strictly before now), keeping order.
  std::vector<std::shared_ptr<ServantRec>> alive;
  for (auto&& e : s->servants) {
    if (e->expires_at < now_ns) {
      s->running_tasks.erase(e->observed_location);  // bookkeeper DropServant, :510-511
    } else {
      alive.push_back(e);
    }
  }
  s->servants.swap(alive);
  // 2. UnsafeSweepOrphans, cc:478-496: forget tasks whose servant left the registry.
  std::unordered_set<const ServantRec*> live;
  for (auto&& e : s->servants) live.insert(e.get());
  std::vector<std::uint64_t> orphans;
  for (auto&& [id, t] : s->tasks) {
    if (!live.count(t.servant.get())) orphans.push_back(id);
  }
  s->FreeTasks(orphans);
  // 3. Expired leases become zombies; they keep occupying their slot (:522-535).
  for (auto&& [id, t] : s->tasks) {
    if (t.expires_at < now_ns) t.zombie = true;
  }
}

void yd_wait_for_starting_new_tasks(yd_sched* s, int64_t now_ns, const yd_task_req* reqs, size_t n,
                                    yd_grant* out) {
  static const std::string kUnknownEnv = "<unknown env id>";
  static const std::string kEmpty;
  std::uint64_t granted = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; i != n; ++i) {
    const std::string& ip = reqs[i].requestor_ip < s->ips.size() ? s->ips[reqs[i].requestor_ip] : kEmpty;
    const std::string& env = reqs[i].env_id < s->envs.size() ? s->envs[reqs[i].env_id] : kUnknownEnv;
    s->Decide(now_ns, ip, reqs[i].min_version, env, reqs[i].expires_in_ns,
              (reqs[i].flags & YD_REQ_FLAG_PREFETCH) != 0, &out[i]);
    granted += out[i].status == YD_STATUS_GRANTED;
  }
  auto t1 = std::chrono::steady_clock::now();
  s->stats = yd_solve_stats{};
  s->stats.solve_ms = s->stats.total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  s->stats.decisions = n;
  s->stats.granted = granted;
  s->have_stats = true;
}

// KeepTaskAlive, task_dispatcher.cc:142-165.
void yd_keep_task_alive(yd_sched* s, int64_t now_ns, const uint64_t* ids, size_t n,
                        int64_t new_expires_in_ns, uint8_t* ok_out) {
  for (size_t i = 0; i != n; ++i) {
    auto it = s->tasks.find(ids[i]);
    if (it == s->tasks.end() || it->second.zombie) { ok_out[i] = 0; continue; }
    it->second.expires_at = now_ns + new_expires_in_ns;
    ok_out[i] = 1;
  }
}

// FreeTask, task_dispatcher.cc:167-170, once per id.
void yd_free_tasks(yd_sched* s, const uint64_t* ids, size_t n) {
  for (size_t i = 0; i != n; ++i) s->FreeTasks({ids[i]});
}

// The batched heartbeat entry points are DEFINED as the loop over the single-servant calls.
void yd_keep_servants_alive(yd_sched* s, int64_t now_ns, const yd_servant* servants, const int64_t* expires_in_ns, size_t n) {
  for (size_t i = 0; i != n; ++i) yd_keep_servant_alive(s, now_ns, &servants[i], expires_in_ns[i]);
}

size_t yd_notify_servants_running_tasks(yd_sched* s, const yd_heartbeat_item* items, size_t n, uint64_t* unknown_out,
                                        size_t* unknown_counts) {
  size_t total = 0;
  for (size_t i = 0; i != n; ++i) {
    const size_t k = yd_notify_servant_running_tasks(s, items[i].servant_location, items[i].tasks, items[i].n_tasks,
                                                     unknown_out + total);
    if (unknown_counts) unknown_counts[i] = k;
    total += k;
  }
  return total;
}

size_t yd_num_servants(yd_sched* s) { return s->servants.size(); }

uint64_t yd_grant_capacity_bound(yd_sched* s) {
  uint64_t b = 0;
  for (auto&& v : s->servants) b += std::min(v->num_processors, v->max_tasks);
  return b;
}

const char* yd_servant_location(yd_sched* s, uint32_t idx) {
  return idx < s->servants.size() ? s->servants[idx]->observed_location.c_str() : nullptr;
}

size_t yd_get_servant_state(yd_sched* s, yd_servant_state* out, size_t cap) {
  for (size_t i = 0; i < s->servants.size() && i < cap; ++i) {
    auto&& e = *s->servants[i];
    out[i] = yd_servant_state{e.running_tasks, e.ever_assigned_tasks, s->CapacityAvailable(e),
                              e.expires_at};
  }
  return s->servants.size();
}

int yd_get_servant_personality(yd_sched* s, uint32_t idx, yd_servant* out) {
  if (idx >= s->servants.size()) return 0;
  const ServantRec& v = *s->servants[idx];
  s->personality_envs.clear();
  for (auto&& e : v.environments) s->personality_envs.push_back(e.c_str());
  if (out) {
    *out = yd_servant{v.version, v.priority, v.reason, (uint32_t)v.environments.size(), v.observed_location.c_str(),
                      v.reported_location.c_str(), s->personality_envs.data(), (uint32_t)v.num_processors,
                      (uint32_t)v.current_load, (uint32_t)v.max_tasks, 0, v.total_memory, v.memory_available};
  }
  return 1;
}

uint64_t yd_next_task_id(yd_sched* s) { return s->next_task_id * s->id_stride + s->id_offset; }
uint64_t yd_num_tasks(yd_sched* s) { return s->tasks.size(); }

// DumpInternals summary, task_dispatcher.cc:540-547,581-584,603-612.
int yd_last_solve_stats(yd_sched* s, yd_solve_stats* out) {
  if (!s->have_stats) return 0;
  *out = s->stats;
  return 1;
}

void* yd_alloc_host(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void yd_free_host(void* p) { std::free(p); }

}  // extern "C"
#include "ydsched_rpc_impl.inc"
#include "yddump_impl.inc"
#include "ydservice_impl.inc"
#include "ydwire_impl.inc"

// ---- bloom pre-filter: restatement of flare's SaltedBloomFilter over XXH64 --------------
// XXH64 is restated from the published xxHash specification (Cyan4973/xxHash, doc/
// xxhash_spec.md; the reference vendors xxHash 0.8.0 under thirdparty/xxhash and calls
// XXH64(data, len, seed = 0) at flare/base/experimental/bloom_filter.cc:21-23).
namespace {
constexpr std::uint64_t kP1 = 11400714785074694791ull, kP2 = 14029467366897019727ull,
                        kP3 = 1609587929392839161ull, kP4 = 9650029242287828579ull,
                        kP5 = 2870177450012600261ull;
inline std::uint64_t Rotl(std::uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline std::uint64_t Rd64(const unsigned char* p) { std::uint64_t v; std::memcpy(&v, p, 8); return v; }
inline std::uint32_t Rd32(const unsigned char* p) { std::uint32_t v; std::memcpy(&v, p, 4); return v; }
inline std::uint64_t Round(std::uint64_t acc, std::uint64_t in) { return Rotl(acc + in * kP2, 31) * kP1; }
inline std::uint64_t Merge(std::uint64_t h, std::uint64_t v) { return (h ^ Round(0, v)) * kP1 + kP4; }

std::uint64_t Xxh64(const unsigned char* p, std::size_t len, std::uint64_t seed) {
  const unsigned char* end = p + len;
  std::uint64_t h;
  if (len >= 32) {
    std::uint64_t v1 = seed + kP1 + kP2, v2 = seed + kP2, v3 = seed, v4 = seed - kP1;
    do {
      v1 = Round(v1, Rd64(p)); v2 = Round(v2, Rd64(p + 8));
      v3 = Round(v3, Rd64(p + 16)); v4 = Round(v4, Rd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = Rotl(v1, 1) + Rotl(v2, 7) + Rotl(v3, 12) + Rotl(v4, 18);
    h = Merge(h, v1); h = Merge(h, v2); h = Merge(h, v3); h = Merge(h, v4);
  } else {
    h = seed + kP5;
  }
  h += len;
  while (p + 8 <= end) { h = Rotl(h ^ Round(0, Rd64(p)), 27) * kP1 + kP4; p += 8; }
  if (p + 4 <= end) { h = Rotl(h ^ (std::uint64_t(Rd32(p)) * kP1), 23) * kP2 + kP3; p += 4; }
  while (p < end) { h = Rotl(h ^ (*p * kP5), 11) * kP1; ++p; }
  h ^= h >> 33; h *= kP2; h ^= h >> 29; h *= kP3; h ^= h >> 32;
  return h;
}

struct Bloom {  // flare/base/experimental/bloom_filter.h:72-128
  std::uint32_t num_hashes = 0;
  std::uint64_t mask = 0;
  std::vector<unsigned char> bytes;
  // SaltedHashGenerator, :178-210: hash i = XXH64(le32(i) || key)
  template <class F> bool ForEachHash(const char* key, std::size_t len, F&& f) const {
    std::vector<unsigned char> buf(len + 4);
    std::memcpy(buf.data() + 4, key, len);
    for (std::uint32_t i = 0; i != num_hashes; ++i) {
      std::memcpy(buf.data(), &i, 4);
      if (!f(Xxh64(buf.data(), buf.size(), 0))) return false;
    }
    return true;
  }
};
std::unordered_map<yd_sched*, Bloom> g_blooms;
}  // namespace

extern "C" int yd_bloom_reset(yd_sched* s, uint64_t size_in_bits, uint32_t num_hashes) {
  if (size_in_bits == 0 || size_in_bits > (1ull << 30) || num_hashes == 0) return 1;
  std::uint64_t bits = 8;  // max(8, GetNextPowerOfTwo(m)), :214-219,:294-297
  while (bits < size_in_bits) bits <<= 1;
  Bloom& b = g_blooms[s];
  b.num_hashes = num_hashes; b.mask = bits - 1; b.bytes.assign(bits / 8, 0);
  return 0;
}
extern "C" int yd_bloom_load(yd_sched* s, const uint8_t* bytes, size_t n_bytes, uint32_t num_hashes) {
  if (n_bytes == 0 || ((n_bytes * 8) & (n_bytes * 8 - 1)) || num_hashes == 0) return 1;  // :225-235
  Bloom& b = g_blooms[s];
  b.num_hashes = num_hashes; b.mask = n_bytes * 8 - 1; b.bytes.assign(bytes, bytes + n_bytes);
  return 0;
}
extern "C" void yd_bloom_add(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride) {
  Bloom& b = g_blooms.at(s);
  for (size_t i = 0; i != n; ++i) {
    b.ForEachHash(keys + i * stride, key_len, [&](std::uint64_t h) {  // :249-254, SetBit :299-303
      std::uint64_t at = h & b.mask; b.bytes[at / 8] |= static_cast<unsigned char>(1u << (at % 8)); return true; });
  }
}
extern "C" void yd_bloom_possibly_contains(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride,
                                           uint8_t* out) {
  Bloom& b = g_blooms.at(s);
  for (size_t i = 0; i != n; ++i) {
    out[i] = b.ForEachHash(keys + i * stride, key_len, [&](std::uint64_t h) {  // :256-261, GetBit :305-309
      std::uint64_t at = h & b.mask; return (b.bytes[at / 8] & (1u << (at % 8))) != 0; });
  }
}
extern "C" size_t yd_bloom_get_bytes(yd_sched* s, uint8_t* out, size_t cap) {
  Bloom& b = g_blooms.at(s);
  if (out) std::memcpy(out, b.bytes.data(), std::min(cap, b.bytes.size()));
  return b.bytes.size();
}

// ---- in-flight task index: restatement of RunningTaskKeeper ---------------------------------
// yadcc/daemon/local/running_task_keeper.cc:40-65 (Refresh: tmp[digest] = desc over the
// GetRunningTasks answer, in order, so the LAST entry of a digest wins) and :67-75 (TryFindTask).

extern "C" size_t yd_running_index_refresh(yd_sched* s) {
  auto& k = s->keeper;
  k.snapshot.clear();
  for (auto&& [loc, v] : s->running_tasks) k.snapshot.insert(k.snapshot.begin(), v.begin(), v.end());  // bookkeeper.cc:36-43
  k.by_digest.clear();
  for (uint32_t i = 0; i < k.snapshot.size(); ++i) k.by_digest[k.snapshot[i].task_digest] = i;
  return k.snapshot.size();
}

extern "C" size_t yd_running_index_size(yd_sched* s) { return s->keeper.by_digest.size(); }

extern "C" void yd_running_index_find(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride,
                                      yd_running_hit* out) {
  if (!out) return;
  auto& k = s->keeper;
  for (size_t i = 0; i < n; ++i) {
    auto it = k.by_digest.find(std::string(keys + i * stride, key_len));
    if (it == k.by_digest.end()) out[i] = yd_running_hit{0, YD_NO_SERVANT, 0};
    else out[i] = yd_running_hit{k.snapshot[it->second].servant_task_id, it->second, 1};
  }
}

extern "C" int yd_running_index_entry(yd_sched* s, uint32_t i, yd_running_task* out) {
  auto& k = s->keeper;
  if (i >= k.snapshot.size()) return 0;
  auto&& t = k.snapshot[i];
  if (out) *out = yd_running_task{t.servant_task_id, t.task_grant_id, t.servant_location.c_str(), t.task_digest.c_str()};
  return 1;
}

#include "ydsched_filter_impl.inc"

// ---- packed interface (yd_wait_for_starting_new_tasks_packed): defined as unpack -> the plain call -> pack --------
extern "C" void yd_wait_for_starting_new_tasks_packed(yd_sched* s, int64_t now_ns, const yd_task_req16* reqs, size_t n,
                                                      yd_grant8* out, yd_packed_ids* ids) {
  yd_packed_ids local{yd_next_task_id(s), 1};
  if (ids) *ids = local;
  if (n == 0) return;
  std::vector<yd_task_req> r(n);
  std::vector<yd_grant> g(n);
  for (size_t i = 0; i != n; ++i) r[i] = yd_unpack_req(reqs[i]);
  yd_wait_for_starting_new_tasks(s, now_ns, r.data(), n, g.data());
  for (size_t i = 0; i != n; ++i) out[i] = yd_pack_grant(g[i], local);
}

// ---- staged queue (yd_stage_requests / yd_wait_for_staged_tasks): host-side copy ----------
namespace { std::unordered_map<yd_sched*, std::vector<yd_task_req>> g_staged; }
extern "C" void yd_stage_requests(yd_sched* s, const yd_task_req* reqs, size_t n) {
  g_staged[s].assign(reqs, reqs + n);
}
extern "C" void yd_wait_for_staged_tasks(yd_sched* s, int64_t now_ns, size_t n, yd_grant* out) {
  auto& q = g_staged[s];
  if (n > q.size()) { std::fprintf(stderr, "ydsched: %zu requests asked for, %zu staged\n", n, q.size()); std::abort(); }
  yd_wait_for_starting_new_tasks(s, now_ns, q.data(), n, out);
}
