// Stand-in for XLA's "xla/ffi/api/ffi.h" (absent from this image): just enough of the
// public API to type-check jax_md_b200/csrc/jmd_ffi.cc -- every handler's C++ signature is
// static_assert'ed against the parameter list its Bind() chain declares, which is what
// the real XLA_FFI_DEFINE_HANDLER_SYMBOL enforces.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <type_traits>

namespace xla::ffi {

enum class ErrorCode { kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string message) : failed_(true), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool failure() const { return failed_; }
  bool success() const { return !failed_; }

 private:
  bool failed_ = false;
  std::string message_;
};

template <typename T>
class ErrorOr {
 public:
  bool has_value() const { return ok_; }
  T& operator*() { return value_; }
  T* operator->() { return &value_; }
  Error error() const { return Error(ErrorCode::kInternal, "stub"); }

 private:
  bool ok_ = true;
  T value_;
};

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T>
  ErrorOr<T> get(size_t) const { return {}; }
};

class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T>
  ErrorOr<Result<T>> get(size_t) const { return {}; }
};

template <typename S>
struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C>
  auto Ctx() const { return CtxOf<C>(); }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  Binding<Ts..., xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  Binding<Ts..., xla::ffi::RemainingRets> RemainingRets() const { return {}; }

  template <typename Fn>
  static constexpr bool Matches = std::is_invocable_r_v<Error, Fn, Ts...>;

 private:
  template <typename C>
  struct CtxHelper;
  template <typename S>
  struct CtxHelper<PlatformStream<S>> { using type = Binding<Ts..., S>; };
  template <typename C>
  typename CtxHelper<C>::type CtxOf() const { return {}; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, fn, binding)                                          \
  static_assert(decltype(binding)::template Matches<decltype(&fn)>,                              \
                #sym ": handler signature does not match its Bind() chain");                     \
  extern "C" void* sym(void* call_frame) { return call_frame; }
