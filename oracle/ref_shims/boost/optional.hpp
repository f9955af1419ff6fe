// Minimal stand-in for boost::optional<T&>, only so that the reference's
// ElemType.hpp compiles in a container without boost.  TEST INFRASTRUCTURE ONLY.
#pragma once
namespace boost {
struct none_t {};
static const none_t none{};
template <class T> class optional;
template <class T> class optional<T&> {
  T* p_;
 public:
  optional() : p_(nullptr) {}
  optional(none_t) : p_(nullptr) {}
  optional(T& r) : p_(&r) {}
  explicit operator bool() const { return p_ != nullptr; }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
};
}  // namespace boost
