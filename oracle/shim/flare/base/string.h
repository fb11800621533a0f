// flare::StartsWith (used by IsNetworkAddressEqual, task_dispatcher.cc:66-69)
// and flare::TryParse<std::size_t> (used by TryParseSize, parse_size.cc:41).
#pragma once
#include <charconv>
#include <cstdint>
#include <cstddef>
#include <optional>
#include <string_view>
namespace flare {
inline bool StartsWith(std::string_view s, std::string_view prefix) {
  return s.size() >= prefix.size() && s.substr(0, prefix.size()) == prefix;
}
template <class T>
std::optional<T> TryParse(std::string_view s) {
  // Whole-string decimal parse; anything else (empty, sign, trailing junk,
  // overflow) is nullopt, which is what the reference's tests pin
  // (yadcc/common/parse_size_test.cc:23-29: "3A" -> nullopt).
  T v{};
  auto* b = s.data();
  auto* e = s.data() + s.size();
  if (b == e) return std::nullopt;
  auto [p, ec] = std::from_chars(b, e, v, 10);
  if (ec != std::errc() || p != e) return std::nullopt;
  return v;
}
}  // namespace flare
