// wfm_math.cuh — correctly-rounded, never-contracted fp64 primitives.
// NumPy / SciPy evaluate a*b+c as two rounded operations; these wrappers keep
// the compiler from fusing them (see wfm_basis.cuh, wfm_iir.cu).
#pragma once

namespace wfm {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

}  // namespace wfm
