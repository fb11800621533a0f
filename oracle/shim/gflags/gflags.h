// Stand-in for gflags: a flag is a plain global (task_dispatcher.cc:35-38).
#pragma once
#include <string>
#define DEFINE_string(name, def, desc) std::string FLAGS_##name = (def)
#define DECLARE_string(name) extern std::string FLAGS_##name
