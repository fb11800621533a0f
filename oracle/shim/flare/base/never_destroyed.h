// flare::NeverDestroyed<T> as used by TaskDispatcher::Instance
// (task_dispatcher.cc:73-76): construct in place, never run the destructor.
#pragma once
#include <new>
#include <utility>
namespace flare {
template <class T>
class NeverDestroyed {
 public:
  template <class... A>
  explicit NeverDestroyed(A&&... a) { new (buf_) T(std::forward<A>(a)...); }
  NeverDestroyed(const NeverDestroyed&) = delete;
  NeverDestroyed& operator=(const NeverDestroyed&) = delete;
  T* Get() noexcept { return reinterpret_cast<T*>(buf_); }
  T* operator->() noexcept { return Get(); }
  T& operator*() noexcept { return *Get(); }

 private:
  alignas(T) unsigned char buf_[sizeof(T)];
};
}  // namespace flare
