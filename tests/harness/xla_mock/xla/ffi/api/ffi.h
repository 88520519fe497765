// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h -- just enough surface for a SYNTAX / type check of
// vivsim_b200/csrc/xla/vivsim_b200_xla.cc against the C ABI in this image, where jaxlib is absent
// (tests/test_abi_cpu.py).  Nothing here executes; the real header comes from jax.ffi.include_dir().
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

namespace xla {
namespace ffi {

enum DataType { F32, S32, U8 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

template <typename T> struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T* begin() const { return ptr; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType D> struct Buffer {
  using T = typename NativeOf<D>::type;
  T* data = nullptr;
  Span<const int64_t> dims;
  T* typed_data() const { return data; }
  Span<const int64_t> dimensions() const { return dims; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};
template <DataType D> struct ResultBuffer {
  Buffer<D> b;
  Buffer<D>* operator->() { return &b; }
};

enum class ErrorCode { kInternal };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};

struct Binding {
  template <typename T> Binding Ctx() { return *this; }
  template <typename T> Binding Arg() { return *this; }
  template <typename T> Binding Ret() { return *this; }
  template <typename T> Binding Attr(const char*) { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro also checks the handler's signature against the binding; the mock only keeps both alive
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding) \
  extern "C" void* symbol() { (void)(binding); return reinterpret_cast<void*>(&fn); }
