// Minimal functional stand-in for jsoncpp's Json::Value: just what
// TaskDispatcher::DumpInternals (task_dispatcher.cc:538-614) touches.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace Json {
using UInt64 = std::uint64_t;
using Int64 = std::int64_t;
class Value {
 public:
  enum Kind { kNull, kInt, kUInt, kBool, kString, kArray, kObject };
  Value() = default;
  Value& operator[](int i) {
    kind_ = kArray;
    if (arr_.size() <= static_cast<std::size_t>(i)) arr_.resize(i + 1);
    return arr_[i];
  }
  Value& operator[](const std::string& k) { kind_ = kObject; return obj_[k]; }
  Value& operator[](const char* k) { kind_ = kObject; return obj_[k]; }
  Value& operator=(bool v) { kind_ = kBool; i_ = v; return *this; }
  Value& operator=(int v) { kind_ = kInt; i_ = v; return *this; }
  Value& operator=(Int64 v) { kind_ = kInt; i_ = v; return *this; }
  Value& operator=(UInt64 v) { kind_ = kUInt; u_ = v; return *this; }
  Value& operator=(const std::string& v) { kind_ = kString; s_ = v; return *this; }
  Value& operator=(const char* v) { kind_ = kString; s_ = v; return *this; }
  void append(const std::string& v) { kind_ = kArray; arr_.emplace_back() = v; }
  Kind kind() const { return kind_; }
  std::int64_t asInt64() const { return kind_ == kUInt ? static_cast<std::int64_t>(u_) : i_; }
  std::uint64_t asUInt64() const { return kind_ == kUInt ? u_ : static_cast<std::uint64_t>(i_); }
  const std::string& asString() const { return s_; }
  const std::vector<Value>& array() const { return arr_; }
  const std::map<std::string, Value>& object() const { return obj_; }
  bool isMember(const std::string& k) const { return obj_.count(k) != 0; }

 private:
  Kind kind_ = kNull;
  std::int64_t i_ = 0;
  std::uint64_t u_ = 0;
  std::string s_;
  std::vector<Value> arr_;
  std::map<std::string, Value> obj_;
};
}  // namespace Json
