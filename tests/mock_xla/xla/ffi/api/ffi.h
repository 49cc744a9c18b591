// MOCK of the slice of xla/ffi/api/ffi.h that turbozero_b200/csrc/tz_jax_ffi.cc uses -- test infrastructure only
// (tests/test_abi.py compiles the adapter against it to catch errors in OUR code; it proves nothing about XLA's API).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
namespace xla { namespace ffi {
enum class ErrorCode { kInvalidArgument, kInternal };
struct Error {
  bool ok = true;
  Error() {}
  Error(ErrorCode, std::string) : ok(false) {}
  static Error Success() { return Error(); }
  bool failure() const { return !ok; }
};
template <typename T> struct Span { const T* p; size_t n; size_t size() const { return n; } const T& operator[](size_t i) const { return p[i]; } };
struct AnyBuffer {
  void* data = nullptr; std::vector<int64_t> dims; size_t bytes = 0;
  void* untyped_data() const { return data; }
  Span<int64_t> dimensions() const { return {dims.data(), dims.size()}; }
  size_t size_bytes() const { return bytes; }
};
template <typename T> struct Result { T v; T* operator->() { return &v; } T& operator*() { return v; } };
template <typename T> struct ErrorOr { T v; T* operator->() { return &v; } T& operator*() { return v; } bool has_value() const { return true; } };
struct RemainingArgs { template <typename T> ErrorOr<T> get(size_t) const { return {}; } size_t size() const { return 0; } };
struct RemainingRets { template <typename T> ErrorOr<Result<T>> get(size_t) const { return {}; } size_t size() const { return 0; } };
template <typename S> struct PlatformStream {};
struct Binding {
  template <typename T> Binding Ctx() const { return *this; }
  template <typename T> Binding Attr(const char*) const { return *this; }
  template <typename T> Binding Arg() const { return *this; }
  template <typename T> Binding Ret() const { return *this; }
  Binding RemainingArgs() const { return *this; }
  Binding RemainingRets() const { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}}  // namespace xla::ffi
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" void* name() { (void)(binding); return (void*)&impl; }
