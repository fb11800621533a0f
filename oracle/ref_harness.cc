// ref_harness.cc -- TEST INFRASTRUCTURE.  Puts the reference's own
// `yadcc::scheduler::TaskDispatcher` (compiled verbatim from /root/reference
// by oracle/Makefile against oracle/shim) behind the C ABI of
// include/ydsched.h, so parity tests and the bench's reference arm can drive
// it with exactly the event streams the CUDA backend sees.
//
// Nothing in yadcc_b200/ may link or load this; only tests/, bench.py's
// reference/cpu_baseline legs and __graft_entry__.smoke() do.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "flare/base/chrono.h"
#include "flare/fiber/timer.h"
#include "gflags/gflags.h"
#include "yadcc/common/parse_size.h"
#include "yadcc/scheduler/task_dispatcher.h"

#include "ydsched.h"

DECLARE_string(servant_min_memory_for_accepting_new_task);

using yadcc::scheduler::RunningTask;
using yadcc::scheduler::ServantPersonality;
using yadcc::scheduler::TaskDispatcher;
using yadcc::scheduler::TaskPersonality;
using yadcc::scheduler::WaitStatus;

// Granted friendship by the FRIEND_TEST shim (oracle/shim/gtest/gtest_prod.h).
struct yd_oracle_access {
  static auto& Servants(TaskDispatcher& d) { return d.servants_.servants; }
  static auto& Tasks(TaskDispatcher& d) { return d.tasks_.tasks; }
  static std::uint64_t NextTaskId(TaskDispatcher& d) { return d.tasks_.next_task_id.load(); }
  static std::size_t Capacity(TaskDispatcher& d, std::size_t i) {
    return d.GetCapacityAvailable(*d.servants_.servants[i]);
  }
  static Json::Value Dump(TaskDispatcher& d) { return d.DumpInternals(); }
  static std::uint64_t TimerId(TaskDispatcher& d) { return d.expiration_timer_; }
};

namespace {

inline void SetNow(std::int64_t now_ns) {
  yd_shim::g_now = std::chrono::steady_clock::time_point(std::chrono::nanoseconds(now_ns));
}

}  // namespace

struct yd_sched {
  std::unique_ptr<TaskDispatcher> d;
  std::vector<std::string> envs;  // env_id -> digest
  std::unordered_map<std::string, std::uint32_t> env_ids;
  std::vector<std::string> ips;  // ip id -> text (id 0 = "")
  std::unordered_map<std::string, std::uint32_t> ip_ids;
  std::vector<RunningTask> running_cache;  // backing store for yd_get_running_tasks
  std::vector<const char*> personality_envs;
  struct KeeperState {  // RunningTaskKeeper state (in-flight task index)
    std::vector<RunningTask> snapshot;
    struct TaskDesc { std::string servant_location; std::uint64_t servant_task_id; std::uint32_t index; };
    std::unordered_map<std::string, TaskDesc> running_tasks;
  } keeper;
  std::string location_cache;
  yd_solve_stats stats{};
  bool have_stats = false;
};

extern "C" {

const char* yd_backend_name(void) { return "reference"; }

int yd_parse_size(const char* text, uint64_t* out_bytes) {
  if (!text || !*text) return 0;  // reference calls .back() on the string: UB on empty
  auto v = yadcc::TryParseSize(text);
  if (!v) return 0;
  *out_bytes = *v;
  return 1;
}

yd_sched* yd_create(const yd_config* cfg) {
  if (!cfg || cfg->abi_version != YD_ABI_VERSION) return nullptr;
  const char* mm = cfg->servant_min_memory_for_accepting_new_task;
  uint64_t parsed;
  if (mm && !yd_parse_size(mm, &parsed)) return nullptr;  // reference would FLARE_CHECK
  FLAGS_servant_min_memory_for_accepting_new_task = mm ? mm : "10G";
  auto* s = new yd_sched;
  s->d = std::make_unique<TaskDispatcher>();
  s->ips.emplace_back();  // id 0 == YD_IP_NONE == ""
  s->ip_ids.emplace("", 0);
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

void yd_keep_servant_alive(yd_sched* s, int64_t now_ns, const yd_servant* sv,
                           int64_t expires_in_ns) {
  SetNow(now_ns);
  ServantPersonality p;
  p.version = sv->version;
  p.observed_location = sv->observed_location;
  p.reported_location = sv->reported_location;
  for (std::uint32_t i = 0; i != sv->num_envs; ++i) {
    p.environments.emplace_back().set_compiler_digest(sv->env_digests[i]);
  }
  p.num_processors = sv->num_processors;
  p.current_load = sv->current_load;
  p.total_memory_in_bytes = sv->total_memory_in_bytes;
  p.memory_available_in_bytes = sv->memory_available_in_bytes;
  p.max_tasks = sv->max_tasks;
  p.priority = static_cast<yadcc::scheduler::ServantPriority>(sv->priority);
  p.not_accepting_task_reason =
      static_cast<yadcc::scheduler::NotAcceptingTaskReason>(sv->not_accepting_task_reason);
  s->d->KeepServantAlive(p, std::chrono::nanoseconds(expires_in_ns));
}

size_t yd_notify_servant_running_tasks(yd_sched* s, const char* servant_location,
                                       const yd_running_task* tasks, size_t n,
                                       uint64_t* unknown_out) {
  std::vector<RunningTask> v(n);
  for (size_t i = 0; i != n; ++i) {
    v[i].set_servant_task_id(tasks[i].servant_task_id);
    v[i].set_task_grant_id(tasks[i].task_grant_id);
    v[i].set_servant_location(tasks[i].servant_location ? tasks[i].servant_location : "");
    v[i].set_task_digest(tasks[i].task_digest ? tasks[i].task_digest : "");
  }
  auto unknown = s->d->NotifyServantRunningTasks(servant_location, std::move(v));
  for (size_t i = 0; i != unknown.size(); ++i) unknown_out[i] = unknown[i];
  return unknown.size();
}

size_t yd_get_running_tasks(yd_sched* s, yd_running_task* out, size_t cap) {
  s->running_cache = s->d->GetRunningTasks();
  for (size_t i = 0; i < s->running_cache.size() && i < cap; ++i) {
    auto&& t = s->running_cache[i];
    out[i] = yd_running_task{t.servant_task_id(), t.task_grant_id(), t.servant_location().c_str(),
                             t.task_digest().c_str()};
  }
  return s->running_cache.size();
}

void yd_on_expiration_timer(yd_sched* s, int64_t now_ns) {
  SetNow(now_ns);
  // Fire the callback this dispatcher registered with fiber::SetTimer.
  auto it = yd_shim::g_timers.find(yd_oracle_access::TimerId(*s->d));
  if (it == yd_shim::g_timers.end()) std::abort();
  it->second();
}

void yd_wait_for_starting_new_tasks(yd_sched* s, int64_t now_ns, const yd_task_req* reqs, size_t n,
                                    yd_grant* out) {
  SetNow(now_ns);
  auto timeout = yd_shim::g_now;  // zero-wait
  // Personalities are materialised before the clock starts: in production they
  // arrive as strings from the RPC layer, so building them is not the
  // dispatcher's cost.
  std::vector<TaskPersonality> tp(n);
  for (size_t i = 0; i != n; ++i) {
    tp[i].requestor_ip = reqs[i].requestor_ip < s->ips.size() ? s->ips[reqs[i].requestor_ip] : "";
    tp[i].min_version = reqs[i].min_version;
    if (reqs[i].env_id < s->envs.size()) {
      tp[i].env_desc.set_compiler_digest(s->envs[reqs[i].env_id]);
    } else {
      tp[i].env_desc.set_compiler_digest("<unknown env id>");
    }
  }
  std::vector<std::string> locations(n);
  std::uint64_t granted = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; i != n; ++i) {
    auto r = s->d->WaitForStartingNewTask(tp[i], std::chrono::nanoseconds(reqs[i].expires_in_ns),
                                          timeout, (reqs[i].flags & YD_REQ_FLAG_PREFETCH) != 0);
    if (r) {
      out[i].task_id = r->task_id;
      out[i].status = YD_STATUS_GRANTED;
      locations[i] = std::move(r->servant_location);
      ++granted;
    } else {
      out[i].task_id = 0;
      out[i].servant_index = YD_NO_SERVANT;
      out[i].status = static_cast<std::uint32_t>(r.error());
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  // Resolve locations to registry positions outside the timed loop.  The
  // registry cannot change during the batch (no heartbeat / timer in between).
  std::unordered_map<std::string, std::uint32_t> pos;
  auto& sv = yd_oracle_access::Servants(*s->d);
  for (std::uint32_t i = 0; i != sv.size(); ++i) {
    pos.emplace(sv[i]->personality.observed_location, i);  // first wins, like find_if
  }
  for (size_t i = 0; i != n; ++i) {
    if (out[i].status == YD_STATUS_GRANTED) out[i].servant_index = pos.at(locations[i]);
  }
  s->stats = yd_solve_stats{};
  s->stats.solve_ms = s->stats.total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  s->stats.decisions = n;
  s->stats.granted = granted;
  s->have_stats = true;
}

void yd_keep_task_alive(yd_sched* s, int64_t now_ns, const uint64_t* ids, size_t n,
                        int64_t new_expires_in_ns, uint8_t* ok_out) {
  SetNow(now_ns);
  for (size_t i = 0; i != n; ++i) {
    ok_out[i] = s->d->KeepTaskAlive(ids[i], std::chrono::nanoseconds(new_expires_in_ns));
  }
}

void yd_free_tasks(yd_sched* s, const uint64_t* ids, size_t n) {
  for (size_t i = 0; i != n; ++i) s->d->FreeTask(ids[i]);
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

size_t yd_num_servants(yd_sched* s) { return yd_oracle_access::Servants(*s->d).size(); }

uint64_t yd_grant_capacity_bound(yd_sched* s) {
  uint64_t b = 0;
  for (auto&& v : yd_oracle_access::Servants(*s->d)) b += std::min(v->personality.num_processors, v->personality.max_tasks);
  return b;
}

// (sizing helper of the batched entry point; the loops below are the reference's own and stop at
// the first failure whatever the counts say)
size_t yd_rpc_expanded_requests(yd_sched* s, const yd_rpc_wait* rpcs, size_t n_rpcs) {
  const uint64_t bound = yd_grant_capacity_bound(s);
  size_t total = 0;
  for (size_t i = 0; i != n_rpcs; ++i) {
    if (rpcs[i].milliseconds_to_wait > 10000u || rpcs[i].next_keep_alive_ns > 30000000000ll) continue;
    total += (size_t)(std::min<uint64_t>(rpcs[i].immediate_reqs, bound + 1) + std::min<uint64_t>(rpcs[i].prefetch_reqs, bound + 1));
  }
  return total;
}

const char* yd_servant_location(yd_sched* s, uint32_t idx) {
  auto& sv = yd_oracle_access::Servants(*s->d);
  if (idx >= sv.size()) return nullptr;
  s->location_cache = sv[idx]->personality.observed_location;
  return s->location_cache.c_str();
}

size_t yd_get_servant_state(yd_sched* s, yd_servant_state* out, size_t cap) {
  auto& sv = yd_oracle_access::Servants(*s->d);
  for (size_t i = 0; i < sv.size() && i < cap; ++i) {
    out[i].running_tasks = sv[i]->running_tasks;
    out[i].ever_assigned_tasks = sv[i]->ever_assigned_tasks;
    out[i].capacity_available = yd_oracle_access::Capacity(*s->d, i);
    out[i].expires_at_ns =
        std::chrono::duration_cast<std::chrono::nanoseconds>(sv[i]->expires_at.time_since_epoch())
            .count();
  }
  return sv.size();
}

int yd_get_servant_personality(yd_sched* s, uint32_t idx, yd_servant* out) {
  auto& sv = yd_oracle_access::Servants(*s->d);
  if (idx >= sv.size()) return 0;
  const ServantPersonality& p = sv[idx]->personality;
  s->personality_envs.clear();
  for (auto&& e : p.environments) s->personality_envs.push_back(e.compiler_digest().c_str());
  if (out) {
    *out = yd_servant{p.version, (int32_t)p.priority, (int32_t)p.not_accepting_task_reason,
                      (uint32_t)p.environments.size(), p.observed_location.c_str(), p.reported_location.c_str(),
                      s->personality_envs.data(), (uint32_t)p.num_processors, (uint32_t)p.current_load,
                      (uint32_t)p.max_tasks, 0, p.total_memory_in_bytes, p.memory_available_in_bytes};
  }
  return 1;
}

uint64_t yd_next_task_id(yd_sched* s) { return yd_oracle_access::NextTaskId(*s->d); }
uint64_t yd_num_tasks(yd_sched* s) { return yd_oracle_access::Tasks(*s->d).size(); }

// TaskDispatcher::DumpInternals (task_dispatcher.cc:538-614), the reference's own Json::Value written out as text:
// everything it holds except the wall-clock strings ("discovered_at", "expires_at") and the "tasks" map.
static void WriteJson(const Json::Value& v, std::string* o) {
  char num[64];
  switch (v.kind()) {
    case Json::Value::kNull: *o += "null"; break;
    case Json::Value::kBool: *o += v.asInt64() ? "true" : "false"; break;
    case Json::Value::kInt: std::snprintf(num, sizeof num, "%lld", (long long)v.asInt64()); *o += num; break;
    case Json::Value::kUInt: std::snprintf(num, sizeof num, "%llu", (unsigned long long)v.asUInt64()); *o += num; break;
    case Json::Value::kString: {
      o->push_back('"');
      for (unsigned char c : v.asString()) {
        if (c == '"' || c == '\\') { o->push_back('\\'); o->push_back((char)c); }
        else if (c < 0x20) { std::snprintf(num, sizeof num, "\\u%04x", c); *o += num; }
        else o->push_back((char)c);
      }
      o->push_back('"');
      break;
    }
    case Json::Value::kArray: {
      o->push_back('[');
      bool first = true;
      for (auto&& e : v.array()) { if (!first) o->push_back(','); first = false; WriteJson(e, o); }
      o->push_back(']');
      break;
    }
    case Json::Value::kObject: {
      o->push_back('{');
      bool first = true;
      for (auto&& [k, e] : v.object()) {
        if (k == "discovered_at" || k == "expires_at" || k == "tasks") continue;
        if (!first) o->push_back(',');
        first = false;
        o->push_back('"'); *o += k; *o += "\":";
        WriteJson(e, o);
      }
      o->push_back('}');
      break;
    }
  }
}

size_t yd_dump_internals_json(yd_sched* s, char* buf, size_t cap) {
  Json::Value j = yd_oracle_access::Dump(*s->d);
  std::string o;
  WriteJson(j, &o);
  if (buf && cap) std::snprintf(buf, cap, "%s", o.c_str());
  return o.size();
}

int yd_last_solve_stats(yd_sched* s, yd_solve_stats* out) {
  if (!s->have_stats) return 0;
  *out = s->stats;
  return 1;
}

void* yd_alloc_host(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void yd_free_host(void* p) { std::free(p); }

}  // extern "C"

// SchedulerServiceImpl::WaitForStartingTask's body (scheduler_service_impl.cc:209-271),
// loop for loop, over the verbatim TaskDispatcher.  This is the oracle for the batched
// expansion in include/ydsched_rpc_impl.inc.
extern "C" size_t yd_wait_for_starting_task_rpcs(yd_sched* s, int64_t now_ns, const yd_rpc_wait* rpcs, size_t n_rpcs,
                                                 yd_rpc_wait_result* results, yd_grant* grants_out, size_t cap) {
  using namespace std::chrono_literals;
  SetNow(now_ns);
  size_t written = 0;
  std::unordered_map<std::string, std::uint32_t> pos;
  auto& sv = yd_oracle_access::Servants(*s->d);
  for (std::uint32_t i = 0; i != sv.size(); ++i) pos.emplace(sv[i]->personality.observed_location, i);
  for (size_t r = 0; r != n_rpcs; ++r) {
    const yd_rpc_wait& q = rpcs[r];
    results[r] = yd_rpc_wait_result{YD_RPC_OK, 0, static_cast<std::uint32_t>(written), 0};
    auto max_wait = q.milliseconds_to_wait * 1ms;
    auto next_keep_alive = std::chrono::nanoseconds(q.next_keep_alive_ns);
    if (max_wait > 10s || next_keep_alive > 30s) {  // :221-226
      results[r].status = YD_RPC_INVALID_ARGUMENT;
      continue;
    }
    TaskPersonality task;  // :228-231
    task.requestor_ip = q.requestor_ip < s->ips.size() ? s->ips[q.requestor_ip] : "";
    task.min_version = q.min_version;
    task.env_desc.set_compiler_digest(q.env_id < s->envs.size() ? s->envs[q.env_id] : "<unknown env id>");
    auto now = yd_shim::g_now;
    std::uint32_t granted = 0;
    bool failed_env = false;
    auto emit = [&](const yadcc::scheduler::TaskAllocation& a) {
      if (written == cap) std::abort();
      grants_out[written++] = yd_grant{a.task_id, pos.at(a.servant_location), YD_STATUS_GRANTED};
      ++granted;
    };
    for (std::uint32_t i = 0; i != q.immediate_reqs; ++i) {  // :234-252
      auto result = s->d->WaitForStartingNewTask(task, next_keep_alive, now /* zero-wait */, false);
      if (!result) {
        if (result.error() == WaitStatus::EnvironmentNotFound) { failed_env = true; }
        break;
      }
      emit(*result);
    }
    if (failed_env) {  // :242-246 SetFailed(STATUS_ENVIRONMENT_NOT_AVAILABLE); return;
      results[r].status = YD_RPC_ENVIRONMENT_NOT_AVAILABLE;
      results[r].n_grants = granted;
      continue;
    }
    for (std::uint32_t i = 0; i != q.prefetch_reqs; ++i) {  // :254-264
      auto result = s->d->WaitForStartingNewTask(task, next_keep_alive, now, true);
      if (!result) break;
      emit(*result);
    }
    results[r].n_grants = granted;
    if (granted == 0) results[r].status = YD_RPC_NO_QUOTA_AVAILABLE;  // :266-270
  }
  return written;
}

// ---- bloom pre-filter: flare's own SaltedBloomFilter + the vendored xxHash -----------
#include "flare/base/experimental/bloom_filter.h"

namespace {
std::unordered_map<yd_sched*, std::unique_ptr<flare::experimental::SaltedBloomFilter>> g_blooms;
}

extern "C" int yd_bloom_reset(yd_sched* s, uint64_t size_in_bits, uint32_t num_hashes) {
  if (size_in_bits == 0 || size_in_bits > (1ull << 30) || num_hashes == 0) return 1;
  g_blooms[s] = std::make_unique<flare::experimental::SaltedBloomFilter>(size_in_bits, num_hashes);
  return 0;
}
extern "C" int yd_bloom_load(yd_sched* s, const uint8_t* bytes, size_t n_bytes, uint32_t num_hashes) {
  if (n_bytes == 0 || ((n_bytes * 8) & (n_bytes * 8 - 1)) || num_hashes == 0) return 1;  // the ctor FLARE_CHECKs this
  g_blooms[s] = std::make_unique<flare::experimental::SaltedBloomFilter>(
      std::string_view(reinterpret_cast<const char*>(bytes), n_bytes), num_hashes);
  return 0;
}
extern "C" void yd_bloom_add(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride) {
  auto& f = *g_blooms.at(s);
  for (size_t i = 0; i != n; ++i) f.Add(std::string_view(keys + i * stride, key_len));
}
extern "C" void yd_bloom_possibly_contains(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride,
                                           uint8_t* out) {
  auto& f = *g_blooms.at(s);
  for (size_t i = 0; i != n; ++i) out[i] = f.PossiblyContains(std::string_view(keys + i * stride, key_len));
}
extern "C" size_t yd_bloom_get_bytes(yd_sched* s, uint8_t* out, size_t cap) {
  auto b = g_blooms.at(s)->GetBytes();
  if (out) std::memcpy(out, b.data(), std::min(cap, b.size()));
  return b.size();
}

// ---- in-flight task index: RunningTaskKeeper's two loops, literally ----------------------------
// running_task_keeper.cc itself needs flare's RPC client (scheduler_stub_), so it cannot be
// compiled here; what it does with the GetRunningTasks answer is restated line by line on top
// of the verbatim TaskDispatcher::GetRunningTasks (Refresh cc:56-64, TryFindTask cc:67-75).

extern "C" size_t yd_running_index_refresh(yd_sched* s) {
  auto& k = s->keeper;
  k.snapshot = s->d->GetRunningTasks();
  std::unordered_map<std::string, yd_sched::KeeperState::TaskDesc> tmp_running_tasks;
  std::uint32_t i = 0;
  for (auto&& running_task : k.snapshot) {
    yd_sched::KeeperState::TaskDesc task_desc = {running_task.servant_location(), running_task.servant_task_id(), i++};
    tmp_running_tasks[running_task.task_digest()] = std::move(task_desc);
  }
  k.running_tasks.swap(tmp_running_tasks);
  return k.snapshot.size();
}

extern "C" size_t yd_running_index_size(yd_sched* s) { return s->keeper.running_tasks.size(); }

extern "C" void yd_running_index_find(yd_sched* s, const char* keys, size_t n, size_t key_len, size_t stride,
                                      yd_running_hit* out) {
  if (!out) return;
  auto& k = s->keeper;
  for (size_t i = 0; i < n; ++i) {
    auto result = k.running_tasks.find(std::string(keys + i * stride, key_len));
    if (result != k.running_tasks.end()) out[i] = yd_running_hit{result->second.servant_task_id, result->second.index, 1};
    else out[i] = yd_running_hit{0, YD_NO_SERVANT, 0};
  }
}

extern "C" int yd_running_index_entry(yd_sched* s, uint32_t i, yd_running_task* out) {
  auto& k = s->keeper;
  if (i >= k.snapshot.size()) return 0;
  auto&& t = k.snapshot[i];
  if (out) *out = yd_running_task{t.servant_task_id(), t.task_grant_id(), t.servant_location().c_str(), t.task_digest().c_str()};
  return 1;
}

// SchedulerServiceImpl's other handlers: host logic over the ABI above (the handlers themselves need
// flare's RPC controller and protobuf, so scheduler_service_impl.cc cannot be compiled here).
#include "ydservice_impl.inc"
#include "ydwire_impl.inc"

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
