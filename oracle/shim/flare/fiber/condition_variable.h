// flare::fiber::ConditionVariable (task_dispatcher.h:290) under the zero-wait
// discipline: wait_until never blocks and always reports timeout, so a request
// that finds no free servant fails with WaitStatus::Timeout at once
// (task_dispatcher.cc:116-118).
#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
namespace yd_shim { inline unsigned long g_notify_all_calls = 0; }
namespace flare::fiber {
class ConditionVariable {
 public:
  template <class Lock, class TP>
  std::cv_status wait_until(Lock&, const TP&) { return std::cv_status::timeout; }
  void notify_all() noexcept { ++yd_shim::g_notify_all_calls; }
  void notify_one() noexcept {}
};
}  // namespace flare::fiber
