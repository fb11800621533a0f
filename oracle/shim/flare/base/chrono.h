// Virtual steady clock replacing flare::ReadCoarseSteadyClock
// (flare/base/chrono.h:60-71).  The harness sets it before every call.
#pragma once
#include <chrono>
namespace yd_shim {
inline std::chrono::steady_clock::time_point g_now{};
}
namespace flare {
inline std::chrono::steady_clock::time_point ReadCoarseSteadyClock() { return yd_shim::g_now; }
inline std::chrono::steady_clock::time_point ReadSteadyClock() { return yd_shim::g_now; }
inline std::chrono::system_clock::time_point ReadSystemClock() {
  return std::chrono::system_clock::time_point(
      std::chrono::duration_cast<std::chrono::system_clock::duration>(
          yd_shim::g_now.time_since_epoch()));
}
}  // namespace flare
