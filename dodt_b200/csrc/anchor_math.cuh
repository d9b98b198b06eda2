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

// The anchor grid of avod/core/anchor_generators/grid_anchor_3d_generator.py:39-108 (tile_anchors_3d)
// + box_3d_encoder.py:85-132 (box_3d_to_anchor) as a function of the anchor index: what
// grid_anchors_kernel writes out and what the fused S2 kernel evaluates in place of a table read.
constexpr int kMaxSizes = 16;
struct GridGeom {
  double x_start, x_delta, z_start, z_delta;   // centres: float32(start + i * delta), as np.arange fills
  double a, b, c, d;                           // ground plane
  double dims[kMaxSizes * 2][3];               // [size * 2 + rotation] -> dim_x, dim_y, dim_z
  int nx, nz, n_sizes;
};
// host: fills g from the arguments of dodt_grid_anchors (defined in anchors.cu); DODT_OK or DODT_EINVAL
int fill_grid_geom(const double ext[6], const double *sizes, int n_sizes, const double stride[2],
                   const double plane[4], GridGeom *g);

__device__ __forceinline__ void grid_anchor(const GridGeom &g, long long i, double a[6]) {
  // meshgrid(x, z, size, rotation) with 'xy' indexing, reshaped row-major: z slowest, then x,
  // then size, then rotation (grid_anchor_3d_generator.py:80-85)
  const int combo = static_cast<int>(i % (2 * g.n_sizes));
  const long long cell = i / (2 * g.n_sizes);
  const int xi = static_cast<int>(cell % g.nx);
  const int zi = static_cast<int>(cell / g.nx);
  a[0] = static_cast<double>(__double2float_rn(__dadd_rn(g.x_start, __dmul_rn(static_cast<double>(xi), g.x_delta))));
  a[2] = static_cast<double>(__double2float_rn(__dadd_rn(g.z_start, __dmul_rn(static_cast<double>(zi), g.z_delta))));
  // all_y = -(a * all_x + c * all_z + d) / b
  a[1] = __ddiv_rn(-__dadd_rn(__dadd_rn(__dmul_rn(g.a, a[0]), __dmul_rn(g.c, a[2])), g.d), g.b);
  a[3] = g.dims[combo][0]; a[4] = g.dims[combo][1]; a[5] = g.dims[combo][2];
}

__device__ __forceinline__ void load_anchor(const double *__restrict__ p, double a[6]) {
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = __ldg(p + k);
}

// regressed anchor r[6] = offset_to_anchor(anchor a[6] (values), offsets o[6] (device memory))
__device__ __forceinline__ void decode_anchor(const double a[6], const float *__restrict__ o, double r[6]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // x = dx * dim_x + x_anchor ; dim = exp(log(dim) + d)
    r[k] = __dadd_rn(__dmul_rn(static_cast<double>(__ldg(o + k)), a[3 + k]), a[k]);
    r[3 + k] = exp(__dadd_rn(log(a[3 + k]), static_cast<double>(__ldg(o + 3 + k))));
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

__device__ __forceinline__ void decode_anchor_f32(const double a[6], const float *__restrict__ o, float r[6]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float pos = __double2float_rn(a[k]), dim = __double2float_rn(a[3 + k]);
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
