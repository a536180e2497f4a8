// Force-included (-include) when compiling the UNMODIFIED reference sources
// /root/reference/models/ops/src/** against torch 2.11 (TEST INFRASTRUCTURE, see oracle/__init__.py).
//
// The reference calls AT_DISPATCH_FLOATING_TYPES(value.type(), ...) (ms_deform_attn_cuda.cu:64, :134);
// torch >= 2.x only accepts an at::ScalarType there.  Instead of patching a copy of the source we
// re-define the macro so that either a ScalarType or a DeprecatedTypeProperties is accepted.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace msda_ref_compat {
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace msda_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
    AT_DISPATCH_SWITCH(msda_ref_compat::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
