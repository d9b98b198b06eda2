// Anchor decoding shared by rpn_decode_kernel (anchors.cu) and the fused S2 kernel
// (anchor_fused.cu): one definition, so that both produce the same bits.
//
// Reference: avod/core/anchor_encoder.py:99-150 (offset_to_anchor), avod/core/anchor_projector.py:
// 13-69 (project_to_bev) + :254-273 (reorder_projected_boxes), as chained by
// avod/core/models/dt_rpn_model.py:573-591. Float64 throughout, one IEEE operation per NumPy
// operation (no FMA contraction), rounded to float32 at the end.
//
// *_f32: the tf.Tensor branches of the same two functions, which is what the reference's INFERENCE
// graph executes on its float32 placeholders (dt_rpn_model.py:568-591): anchors rounded to float32
// by the feed, then one float32 operation per TF op. exp / log are evaluated in float64 and rounded,
// i.e. correctly rounded float32 results (TF's own GPU expf / logf are within 2 ulp of that).
#pragma once
#include <math.h>

namespace dodt {

// regressed anchor r[6] = offset_to_anchor(anchor a[6], offsets o[6])
__device__ __forceinline__ void decode_anchor(const double *__restrict__ a, const float *__restrict__ o,
                                              double r[6]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // x = dx * dim_x + x_anchor ; dim = exp(log(dim) + d)
    r[k] = __dadd_rn(__dmul_rn(static_cast<double>(__ldg(o + k)), __ldg(a + 3 + k)), __ldg(a + k));
    r[3 + k] = exp(__dadd_rn(log(__ldg(a + 3 + k)), static_cast<double>(__ldg(o + 3 + k))));
  }
}

// [z1, x1, z2, x2] normalised BEV box of a regressed anchor (project_to_bev + reorder)
__device__ __forceinline__ float4 bev_box_of(const double r[6], double x_min, double x_max, double z_min,
                                             double z_max) {
  const double hx = __ddiv_rn(r[3], 2.0), hz = __ddiv_rn(r[5], 2.0);
  const double xr = __dsub_rn(x_max, x_min), zr = __dsub_rn(z_max, z_min);
  const double x1 = __ddiv_rn(__dsub_rn(__dsub_rn(r[0], hx), x_min), xr);
  const double x2 = __ddiv_rn(__dsub_rn(__dadd_rn(r[0], hx), x_min), xr);
  const double z1 = __ddiv_rn(__dsub_rn(__dsub_rn(z_max, __dadd_rn(r[2], hz)), z_min), zr);
  const double z2 = __ddiv_rn(__dsub_rn(__dsub_rn(z_max, __dsub_rn(r[2], hz)), z_min), zr);
  return make_float4(__double2float_rn(z1), __double2float_rn(x1), __double2float_rn(z2), __double2float_rn(x2));
}

__device__ __forceinline__ float exp_f32(float x) { return __double2float_rn(exp(static_cast<double>(x))); }
__device__ __forceinline__ float log_f32(float x) { return __double2float_rn(log(static_cast<double>(x))); }

__device__ __forceinline__ void decode_anchor_f32(const double *__restrict__ a, const float *__restrict__ o,
                                                  float r[6]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float pos = __double2float_rn(__ldg(a + k)), dim = __double2float_rn(__ldg(a + 3 + k));
    r[k] = __fadd_rn(__fmul_rn(__ldg(o + k), dim), pos);
    r[3 + k] = exp_f32(__fadd_rn(log_f32(dim), __ldg(o + 3 + k)));
  }
}

// the extents enter the TF graph as Python floats: their differences are formed in float64 on the
// host and become float32 constants
struct BevExtentsF32 {
  float x_min, z_min, z_max, x_range, z_range;
};
__host__ __device__ __forceinline__ BevExtentsF32 bev_extents_f32(double x_min, double x_max, double z_min,
                                                                  double z_max) {
  BevExtentsF32 e;
  e.x_min = static_cast<float>(x_min); e.z_min = static_cast<float>(z_min); e.z_max = static_cast<float>(z_max);
  e.x_range = static_cast<float>(x_max - x_min); e.z_range = static_cast<float>(z_max - z_min);
  return e;
}

__device__ __forceinline__ float4 bev_box_of_f32(const float r[6], const BevExtentsF32 e) {
  const float hx = __fdiv_rn(r[3], 2.0f), hz = __fdiv_rn(r[5], 2.0f);
  const float x1 = __fdiv_rn(__fsub_rn(__fsub_rn(r[0], hx), e.x_min), e.x_range);
  const float x2 = __fdiv_rn(__fsub_rn(__fadd_rn(r[0], hx), e.x_min), e.x_range);
  const float z1 = __fdiv_rn(__fsub_rn(__fsub_rn(e.z_max, __fadd_rn(r[2], hz)), e.z_min), e.z_range);
  const float z2 = __fdiv_rn(__fsub_rn(__fsub_rn(e.z_max, __fsub_rn(r[2], hz)), e.z_min), e.z_range);
  return make_float4(z1, x1, z2, x2);
}

}  // namespace dodt
