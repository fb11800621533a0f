// Only `SystemClockView(steady time_point).Get()` is used, by FormatTime
// (task_dispatcher.cc:46-53), for the debug dump.
#pragma once
#include <chrono>
#include "flare/base/chrono.h"
namespace flare::internal {
class SystemClockView {
 public:
  /* implicit */ SystemClockView(std::chrono::steady_clock::time_point tp)
      : tp_(ReadSystemClock() + std::chrono::duration_cast<std::chrono::system_clock::duration>(
                                    tp - ReadSteadyClock())) {}
  const std::chrono::system_clock::time_point& Get() const noexcept { return tp_; }

 private:
  std::chrono::system_clock::time_point tp_;
};
}  // namespace flare::internal
