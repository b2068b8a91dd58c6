// torch >= 2.5 dropped the AT_DISPATCH_*(tensor.type(), ...) overload that the reference's
// altcorr/correlation_kernel.cu:211,273,299,325 relies on.  Force-included (nvcc -include) so the reference
// source compiles unmodified.  Test infrastructure only.
#pragma once
#include <ATen/core/DeprecatedTypeProperties.h>
#include <c10/core/ScalarType.h>
namespace detail {
inline c10::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
