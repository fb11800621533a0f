// FLARE_CHECK* abort like the real ones; the FLARE_LOG* family is silenced
// (arguments are not evaluated -- none of the reference's log arguments have
// side effects: task_dispatcher.cc:105,151,156,177,212-218,267,339-342,...).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
namespace yd_shim {
[[noreturn]] inline void CheckFailed(const char* expr, const char* file, int line) {
  std::fprintf(stderr, "FLARE_CHECK failed: %s  (%s:%d)\n", expr, file, line);
  std::abort();
}
}  // namespace yd_shim
#define FLARE_CHECK(expr, ...) \
  do { if (!(expr)) ::yd_shim::CheckFailed(#expr, __FILE__, __LINE__); } while (0)
#define FLARE_CHECK_EQ(a, b, ...) FLARE_CHECK((a) == (b))
#define FLARE_CHECK_NE(a, b, ...) FLARE_CHECK((a) != (b))
#define FLARE_CHECK_GT(a, b, ...) FLARE_CHECK((a) > (b))
#define FLARE_CHECK_GE(a, b, ...) FLARE_CHECK((a) >= (b))
#define FLARE_CHECK_LT(a, b, ...) FLARE_CHECK((a) < (b))
#define FLARE_CHECK_LE(a, b, ...) FLARE_CHECK((a) <= (b))
#define FLARE_LOG_INFO(...) ((void)0)
#define FLARE_LOG_WARNING(...) ((void)0)
#define FLARE_LOG_ERROR(...) ((void)0)
#define FLARE_LOG_WARNING_EVERY_SECOND(...) ((void)0)
#define FLARE_LOG_ERROR_EVERY_SECOND(...) ((void)0)
#define FLARE_LOG_WARNING_IF(cond, ...) ((void)0)
#define FLARE_LOG_ERROR_IF_EVERY_SECOND(cond, ...) ((void)0)
#define FLARE_VLOG(...) ((void)0)
