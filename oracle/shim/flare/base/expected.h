// flare::Expected<T, E>: value-or-error, as returned by WaitForStartingNewTask
// (task_dispatcher.h:138-140).
#pragma once
#include <utility>
#include <variant>
namespace flare {
template <class T, class E>
class Expected {
 public:
  /* implicit */ Expected(T v) : v_(std::in_place_index<0>, std::move(v)) {}
  /* implicit */ Expected(E e) : v_(std::in_place_index<1>, std::move(e)) {}
  explicit operator bool() const noexcept { return v_.index() == 0; }
  T* operator->() { return &std::get<0>(v_); }
  const T* operator->() const { return &std::get<0>(v_); }
  T& operator*() { return std::get<0>(v_); }
  const T& operator*() const { return std::get<0>(v_); }
  T& value() { return std::get<0>(v_); }
  const E& error() const { return std::get<1>(v_); }

 private:
  std::variant<T, E> v_;
};
}  // namespace flare
