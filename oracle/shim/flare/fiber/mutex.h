// flare::fiber::Mutex (task_dispatcher.h:289).  The harness is single-threaded.
#pragma once
#include <mutex>
namespace flare::fiber { using Mutex = std::mutex; }
