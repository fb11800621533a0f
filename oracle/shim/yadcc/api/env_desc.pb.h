// Plain-struct stand-in for protoc output of yadcc/api/env_desc.proto:20-42.
#pragma once
#include <string>
namespace yadcc {
class EnvironmentDesc {
 public:
  const std::string& compiler_digest() const { return compiler_digest_; }
  void set_compiler_digest(std::string v) { compiler_digest_ = std::move(v); }

 private:
  std::string compiler_digest_;
};
}  // namespace yadcc
