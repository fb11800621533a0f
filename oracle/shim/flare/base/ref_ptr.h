// Intrusive ref counting: flare::RefCounted / RefPtr / MakeRefCounted, as used
// for ServantDesc (task_dispatcher.h:184-197, .cc:132,205-206).
#pragma once
#include <cstddef>
#include <utility>
namespace flare {
struct ref_ptr_t { explicit ref_ptr_t() = default; };
struct adopt_ptr_t { explicit adopt_ptr_t() = default; };
inline constexpr ref_ptr_t ref_ptr{};
inline constexpr adopt_ptr_t adopt_ptr{};

template <class T>
class RefCounted {
 public:
  void Ref() noexcept { ++refs_; }
  void Deref() noexcept { if (--refs_ == 0) delete static_cast<T*>(this); }
  std::size_t UnsafeRefCount() const noexcept { return refs_; }

 protected:
  RefCounted() = default;
  ~RefCounted() = default;

 private:
  std::size_t refs_ = 1;  // Single-threaded harness: no atomics needed.
};

template <class T>
class RefPtr {
 public:
  constexpr RefPtr() noexcept = default;
  /* implicit */ constexpr RefPtr(std::nullptr_t) noexcept {}
  RefPtr(ref_ptr_t, T* p) noexcept : p_(p) { if (p_) p_->Ref(); }
  RefPtr(adopt_ptr_t, T* p) noexcept : p_(p) {}
  RefPtr(const RefPtr& o) noexcept : p_(o.p_) { if (p_) p_->Ref(); }
  RefPtr(RefPtr&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  ~RefPtr() { if (p_) p_->Deref(); }
  RefPtr& operator=(const RefPtr& o) noexcept { RefPtr t(o); std::swap(p_, t.p_); return *this; }
  RefPtr& operator=(RefPtr&& o) noexcept { RefPtr t(std::move(o)); std::swap(p_, t.p_); return *this; }
  T* Get() const noexcept { return p_; }
  T* operator->() const noexcept { return p_; }
  T& operator*() const noexcept { return *p_; }
  explicit operator bool() const noexcept { return p_ != nullptr; }

 private:
  T* p_ = nullptr;
};

template <class T, class... A>
RefPtr<T> MakeRefCounted(A&&... a) { return RefPtr<T>(adopt_ptr, new T(std::forward<A>(a)...)); }
}  // namespace flare
