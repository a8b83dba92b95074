// Force-included (-include) when compiling the UNMODIFIED reference .cu files with today's libtorch
// (oracle/build_ref.py).  The reference dispatches on `tensor.type()` (a DeprecatedTypeProperties,
// e.g. channelnorm_kernel.cu:111, correlation_cuda_kernel.cu:386); libtorch 2.x removed the
// `::detail::scalar_type(const DeprecatedTypeProperties&)` overload its AT_DISPATCH macros call.
// Restoring that one overload is the whole shim — no reference source is edited or copied.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <ATen/core/DeprecatedTypeProperties.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
