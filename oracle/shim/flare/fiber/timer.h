// flare::fiber::SetTimer / KillTimer (task_dispatcher.cc:81-82,90-92): the
// callback is captured, and the harness fires it when the stream says `Tick`.
#pragma once
#include <chrono>
#include <cstdint>
#include <functional>
#include <map>
#include "flare/base/chrono.h"
namespace yd_shim {
inline std::map<std::uint64_t, std::function<void()>> g_timers;
inline std::uint64_t g_next_timer_id = 1;
}  // namespace yd_shim
namespace flare::fiber {
template <class TP, class D, class F>
std::uint64_t SetTimer(TP, D, F&& cb) {
  auto id = yd_shim::g_next_timer_id++;
  yd_shim::g_timers.emplace(id, std::forward<F>(cb));
  return id;
}
inline void KillTimer(std::uint64_t id) { yd_shim::g_timers.erase(id); }
}  // namespace flare::fiber
