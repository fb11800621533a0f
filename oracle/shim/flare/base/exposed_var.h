// flare::ExposedVarDynamic<T>: remembers the getter so the harness can pull
// DumpInternals the way GET /inspect/vars/yadcc/task_dispatcher would
// (task_dispatcher.cc:79-80).
#pragma once
#include <functional>
#include <string>
#include <utility>
namespace flare {
template <class T>
class ExposedVarDynamic {
 public:
  ExposedVarDynamic(std::string path, std::function<T()> getter)
      : path_(std::move(path)), getter_(std::move(getter)) {}
  T Read() const { return getter_(); }
  const std::string& path() const { return path_; }

 private:
  std::string path_;
  std::function<T()> getter_;
};
}  // namespace flare
