// kmath.cuh -- the floating-point expressions of the apply / MINRES kernels with EXPLICIT rounding
// (__fma_rn / __dmul_rn), shared by the multi-launch kernels (apply.cu, krylov.cu) and the persistent
// MINRES kernel.  The compiler's own FMA contraction depends on the surrounding code; writing the
// operations out makes the two schedules of the same loop produce identical bits (tested), which is what
// the "bit-identical for any number of GPUs" property rests on once one GPU runs the persistent kernel
// and several GPUs run the multi-launch one.
#pragma once
#include <cuda_runtime.h>

namespace nosh {

// complex multiply-accumulate  acc += v * x   (two fused multiply-adds per component)
__device__ __forceinline__ void cfma(double2 &acc, double2 v, double2 x) {
  acc.x = __fma_rn(-v.y, x.y, __fma_rn(v.x, x.x, acc.x));
  acc.y = __fma_rn(v.y, x.x, __fma_rn(v.x, x.y, acc.y));
}
// fp32 matrix value (mixed-precision V-cycle): widened at the FMA, so that a thread's values in flight cost half the
// registers
__device__ __forceinline__ void cfma(double2 &acc, float2 v, double2 x) { cfma(acc, make_double2((double)v.x, (double)v.y), x); }
__device__ __forceinline__ double2 scaled(double2 x, double s) {
  return make_double2(__dmul_rn(x.x, s), __dmul_rn(x.y, s));
}
// jacobian_operator::apply epilogue (src/jacobian_operator.cpp:95-100): acc + [[d0.x, d1],[d1, d0.y]] x
__device__ __forceinline__ double2 diag_epilogue(double2 acc, double2 d0, double d1, double2 x) {
  return make_double2(__fma_rn(d1, x.y, __fma_rn(d0.x, x.x, acc.x)), __fma_rn(d0.y, x.y, __fma_rn(d1, x.x, acc.y)));
}
// y - f r
__device__ __forceinline__ double2 sub_scaled(double2 y, double f, double2 r) {
  return make_double2(__fma_rn(-f, r.x, y.x), __fma_rn(-f, r.y, y.y));
}
// Re <a, b> of one complex entry
__device__ __forceinline__ double cdot(double2 a, double2 b) { return __fma_rn(a.y, b.y, __dmul_rn(a.x, b.x)); }
// MINRES direction update: w = ((r ib - oe a) - de b) ig ;  x += ph w
__device__ __forceinline__ double2 minres_w(double2 r, double2 a, double2 b, double ib, double oe, double de, double ig) {
  return make_double2(__dmul_rn(__fma_rn(-de, b.x, __fma_rn(-oe, a.x, __dmul_rn(r.x, ib))), ig),
                      __dmul_rn(__fma_rn(-de, b.y, __fma_rn(-oe, a.y, __dmul_rn(r.y, ib))), ig));
}
__device__ __forceinline__ double2 axpy2(double a, double2 x, double2 y) {
  return make_double2(__fma_rn(a, x.x, y.x), __fma_rn(a, x.y, y.y));
}

}  // namespace nosh
