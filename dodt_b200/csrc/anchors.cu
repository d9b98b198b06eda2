// Anchor geometry on the device: the host-side NumPy helpers that sit between
// the stages of the front end in the reference — anchor grid generation, projection of anchors
// into the BEV map and the image, decoding of regressed offsets — so that a frame never has to
// upload 89 600 x 6 doubles or their projections.
//
// Reference behaviour reproduced (paths relative to the Guoxs/DODT checkout):
//   avod/core/anchor_generators/grid_anchor_3d_generator.py:39-108   tile_anchors_3d
//   avod/core/box_3d_encoder.py:85-132                               box_3d_to_anchor
//   avod/core/anchor_projector.py:13-69                              project_to_bev
//   avod/core/anchor_projector.py:72-156                             project_to_image_space
//       (+ wavedata/wavedata/tools/core/calib_utils.py:394-410       project_to_image)
//   avod/core/anchor_encoder.py:99-150                               offset_to_anchor
//
// Exactness. The grid and the BEV projection are sums, differences, products and quotients of
// float64 values in a fixed order: they are evaluated here with the same IEEE operations
// (no FMA contraction) and are bit-identical to NumPy. The image projection goes through
// np.dot (BLAS: summation order and FMA use are implementation defined) and the offset decoding
// through exp/log (libm): those two agree to float64 rounding noise, far inside the 1e-5 bar, and
// the float32 boxes they feed are identical except where a value sits on a float32 rounding edge.
#include <math.h>

#include "anchor_math.cuh"
#include "common.cuh"

namespace dodt {
namespace {

__global__ void __launch_bounds__(256)
grid_anchors_kernel(const GridGeom g, long long n, double *__restrict__ anchors) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  double a[6];
  grid_anchor(g, i, a);
  double *o = anchors + i * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = a[k];
}

template <typename T>
__global__ void __launch_bounds__(256)
project_bev_kernel(const T *__restrict__ anchors, long long n, double x_min, double x_max,
                   double z_min, double z_max, int tf_order, float *__restrict__ norm,
                   float *__restrict__ metres) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const T *a = anchors + i * 6;
  const double x = a[0], z = a[2];
  const double hx = __ddiv_rn(static_cast<double>(a[3]), 2.0), hz = __ddiv_rn(static_cast<double>(a[5]), 2.0);
  const double xr = __dsub_rn(x_max, x_min), zr = __dsub_rn(z_max, z_min);
  // corners relative to the top-left of the map (z flipped), then divided by the extent ranges
  const double x1 = __dsub_rn(__dsub_rn(x, hx), x_min);
  const double x2 = __dsub_rn(__dadd_rn(x, hx), x_min);
  const double z1 = __dsub_rn(__dsub_rn(z_max, __dadd_rn(z, hz)), z_min);
  const double z2 = __dsub_rn(__dsub_rn(z_max, __dsub_rn(z, hz)), z_min);
  const double c[4] = {x1, z1, x2, z2};
  const double r[4] = {xr, zr, xr, zr};
  // tf_order: [y1, x1, y2, x2] = [z1, x1, z2, x2] (anchor_projector.reorder_projected_boxes)
  const int perm[4] = {1, 0, 3, 2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int s = tf_order ? perm[k] : k;
    if (norm) norm[i * 4 + k] = __double2float_rn(__ddiv_rn(c[s], r[s]));
    if (metres) metres[i * 4 + k] = __double2float_rn(c[s]);
  }
}

struct Calib {
  double p[12];   // 3 x 4 row-major
};

template <typename T>
__global__ void __launch_bounds__(256)
project_image_kernel(const T *__restrict__ anchors, long long n, const Calib P, double img_h,
                     double img_w, int tf_order, float *__restrict__ norm, float *__restrict__ pixels) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const T *a = anchors + i * 6;
  const double x = a[0], y = a[1], z = a[2];
  const double hx = __ddiv_rn(static_cast<double>(a[3]), 2.0), dy = a[4], hz = __ddiv_rn(static_cast<double>(a[5]), 2.0);
  const double xs[2] = {__dadd_rn(x, hx), __dsub_rn(x, hx)};
  const double ys[2] = {y, __dsub_rn(y, dy)};
  const double zs[2] = {__dadd_rn(z, hz), __dsub_rn(z, hz)};
  double u_min = 0, u_max = 0, v_min = 0, v_max = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {   // the 8 cuboid corners (any order: only min / max are kept)
    const double cx = xs[k & 1], cy = ys[(k >> 1) & 1], cz = zs[(k >> 2) & 1];
    const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[8], cx), __dmul_rn(P.p[9], cy)), __dmul_rn(P.p[10], cz)), P.p[11]);
    const double u = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[0], cx), __dmul_rn(P.p[1], cy)), __dmul_rn(P.p[2], cz)), P.p[3]), w);
    const double v = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[4], cx), __dmul_rn(P.p[5], cy)), __dmul_rn(P.p[6], cz)), P.p[7]), w);
    if (k == 0) { u_min = u_max = u; v_min = v_max = v; }
    else { u_min = fmin(u_min, u); u_max = fmax(u_max, u); v_min = fmin(v_min, v); v_max = fmax(v_max, v); }
  }
  const double c[4] = {u_min, v_min, u_max, v_max};
  const double r[4] = {img_w, img_h, img_w, img_h};
  const int perm[4] = {1, 0, 3, 2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int s = tf_order ? perm[k] : k;
    if (norm) norm[i * 4 + k] = __double2float_rn(__ddiv_rn(c[s], r[s]));
    if (pixels) pixels[i * 4 + k] = __double2float_rn(c[s]);
  }
}

template <typename TA, typename TO>
__global__ void __launch_bounds__(256)
offset_to_anchor_kernel(const TA *__restrict__ anchors, const TO *__restrict__ offsets, long long n,
                        double *__restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const TA *a = anchors + i * 6;
  const TO *o = offsets + i * 6;
  double *r = out + i * 6;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // x = dx * dim_x + x_anchor ; dim = exp(log(dim) + d)
    r[k] = __dadd_rn(__dmul_rn(static_cast<double>(o[k]), static_cast<double>(a[3 + k])), static_cast<double>(a[k]));
    r[3 + k] = exp(__dadd_rn(log(static_cast<double>(a[3 + k])), static_cast<double>(o[3 + k])));
  }
}

// What the reference does between the RPN head and NMS / the second-stage crops
// (avod/core/models/dt_rpn_model.py:573-591,618-660): regressed anchors of the kept anchors,
// projected into the BEV map and the image. One thread per kept anchor, float64 throughout.
__global__ void __launch_bounds__(128)
rpn_decode_kernel(const double *__restrict__ anchors, const float *__restrict__ offsets,
                  const int *__restrict__ idx, const int *__restrict__ idx2,
                  const int *__restrict__ count, int n_max,
                  double x_min, double x_max, double z_min, double z_max, const Calib P,
                  double img_h, double img_w, int decode_f32, float *__restrict__ bev_boxes,
                  float *__restrict__ img_boxes) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n_max || i >= __ldg(count)) return;
  const size_t src = static_cast<size_t>(__ldg(idx + (idx2 ? __ldg(idx2 + i) : i)));
  double a[6];
  load_anchor(anchors + src * 6, a);
  const float *o = offsets + src * 6;
  double r[6];
  if (decode_f32) {
    // the TF graph's float32 chain; the image projection below (a float32 matmul in TF, whose
    // summation order is cuBLAS's) continues in float64 from the float32 regressed anchor
    float rf[6];
    decode_anchor_f32(a, o, rf);
    if (bev_boxes)
      reinterpret_cast<float4 *>(bev_boxes)[i] = bev_box_of_f32(rf, bev_extents_f32(x_min, x_max, z_min, z_max));
#pragma unroll
    for (int k = 0; k < 6; ++k) r[k] = rf[k];
  } else {
    decode_anchor(a, o, r);
    if (bev_boxes)   // [z1, x1, z2, x2] normalised (project_to_bev + reorder_projected_boxes)
      reinterpret_cast<float4 *>(bev_boxes)[i] = bev_box_of(r, x_min, x_max, z_min, z_max);
  }
  const double hx = __ddiv_rn(r[3], 2.0), hz = __ddiv_rn(r[5], 2.0);
  if (img_boxes) {   // [y1, x1, y2, x2] normalised (project_to_image_space + reorder)
    const double xs[2] = {__dadd_rn(r[0], hx), __dsub_rn(r[0], hx)};
    const double ys[2] = {r[1], __dsub_rn(r[1], r[4])};
    const double zs[2] = {__dadd_rn(r[2], hz), __dsub_rn(r[2], hz)};
    double u_min = 0, u_max = 0, v_min = 0, v_max = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double cx = xs[k & 1], cy = ys[(k >> 1) & 1], cz = zs[(k >> 2) & 1];
      const double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[8], cx), __dmul_rn(P.p[9], cy)), __dmul_rn(P.p[10], cz)), P.p[11]);
      const double u = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[0], cx), __dmul_rn(P.p[1], cy)), __dmul_rn(P.p[2], cz)), P.p[3]), w);
      const double v = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.p[4], cx), __dmul_rn(P.p[5], cy)), __dmul_rn(P.p[6], cz)), P.p[7]), w);
      if (k == 0) { u_min = u_max = u; v_min = v_max = v; }
      else { u_min = fmin(u_min, u); u_max = fmax(u_max, u); v_min = fmin(v_min, v); v_max = fmax(v_max, v); }
    }
    reinterpret_cast<float4 *>(img_boxes)[i] =
        make_float4(__double2float_rn(__ddiv_rn(v_min, img_h)), __double2float_rn(__ddiv_rn(u_min, img_w)),
                    __double2float_rn(__ddiv_rn(v_max, img_h)), __double2float_rn(__ddiv_rn(u_max, img_w)));
  }
}

// np.arange(start, stop, step): length ceil((stop - start) / step), element i = start + i * delta
// with delta = (start + step) - start
int arange_len(double start, double stop, double step) {
  const double len = ceil((stop - start) / step);
  return len > 0 ? static_cast<int>(len) : 0;
}

int fill_grid(const double ext[6], const double stride[2], int n_sizes, GridGeom *g) {
  if (!ext || !stride || !(stride[0] > 0.0) || !(stride[1] > 0.0) || n_sizes <= 0 || n_sizes > kMaxSizes)
    return DODT_EINVAL;
  volatile double xs = ext[0] + stride[0] / 2.0;
  volatile double zs = ext[5] - stride[1] / 2.0;
  volatile double xn = xs + stride[0], zn = zs + (-stride[1]);
  g->x_start = xs;
  g->z_start = zs;
  g->x_delta = xn - xs;
  g->z_delta = zn - zs;
  g->nx = arange_len(xs, ext[1], stride[0]);
  g->nz = arange_len(zs, ext[4], -stride[1]);
  g->n_sizes = n_sizes;
  return DODT_OK;
}

}  // namespace

int fill_grid_geom(const double ext[6], const double *sizes, int n_sizes, const double stride[2],
                   const double plane[4], GridGeom *g) {
  const int rc = fill_grid(ext, stride, n_sizes, g);
  if (rc != DODT_OK) return rc;
  if (!sizes || !plane || plane[1] == 0.0) return DODT_EINVAL;
  g->a = plane[0]; g->b = plane[1]; g->c = plane[2]; g->d = plane[3];
  const double rot[2] = {0.0, M_PI / 2.0};
  for (int s = 0; s < n_sizes; ++s)
    for (int r = 0; r < 2; ++r) {
      // box_3d_encoder.py:120-130: dim_x = l*|cos| + w*|sin|, dim_y = h, dim_z = w*|cos| + l*|sin|
      volatile double cr = fabs(cos(rot[r])), sr = fabs(sin(rot[r]));
      const double l = sizes[3 * s], w = sizes[3 * s + 1], h = sizes[3 * s + 2];
      volatile double lc = l * cr, ws = w * sr, wc = w * cr, ls = l * sr;
      g->dims[2 * s + r][0] = lc + ws;
      g->dims[2 * s + r][1] = h;
      g->dims[2 * s + r][2] = wc + ls;
    }
  return DODT_OK;
}

}  // namespace dodt

extern "C" {

int dodt_grid_anchor_shape(const double extents[6], const double stride[2], int32_t n_sizes,
                           int32_t shape[4]) {
  using namespace dodt;
  GridGeom g;
  const int rc = fill_grid(extents, stride, n_sizes, &g);
  if (rc != DODT_OK) return rc;
  if (!shape) return DODT_EINVAL;
  shape[0] = g.nz; shape[1] = g.nx; shape[2] = n_sizes; shape[3] = 2;
  return DODT_OK;
}

int dodt_grid_anchors(const double extents[6], const double *sizes, int32_t n_sizes,
                      const double stride[2], const double plane[4], double *anchors,
                      dodt_stream_t stream_) {
  using namespace dodt;
  GridGeom g;
  const int rc = fill_grid_geom(extents, sizes, n_sizes, stride, plane, &g);
  if (rc != DODT_OK) return rc;
  const long long n = static_cast<long long>(g.nx) * g.nz * n_sizes * 2;
  if (n == 0) return DODT_OK;
  if (!anchors) return DODT_EINVAL;
  if ((n + 255) / 256 > 0x7FFFFFFFll) return DODT_ECAPACITY;
  grid_anchors_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream_)>>>(g, n, anchors);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_project_to_bev(const void *anchors, int32_t dtype, int64_t n, const double bev_extents[4],
                        int32_t tf_order, float *boxes_norm, float *boxes_metres,
                        dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || !bev_extents || (dtype != DODT_F32 && dtype != DODT_F64)) return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  if (!anchors || (!boxes_norm && !boxes_metres)) return DODT_EINVAL;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (dtype == DODT_F64)
    project_bev_kernel<double><<<blocks, 256, 0, as_stream(stream_)>>>(
        static_cast<const double *>(anchors), n, bev_extents[0], bev_extents[1], bev_extents[2],
        bev_extents[3], tf_order, boxes_norm, boxes_metres);
  else
    project_bev_kernel<float><<<blocks, 256, 0, as_stream(stream_)>>>(
        static_cast<const float *>(anchors), n, bev_extents[0], bev_extents[1], bev_extents[2],
        bev_extents[3], tf_order, boxes_norm, boxes_metres);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_project_to_image_space(const void *anchors, int32_t dtype, int64_t n, const double p2[12],
                                int32_t image_h, int32_t image_w, int32_t tf_order,
                                float *boxes_norm, float *boxes_pixels, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || !p2 || image_h <= 0 || image_w <= 0 || (dtype != DODT_F32 && dtype != DODT_F64))
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  if (!anchors || (!boxes_norm && !boxes_pixels)) return DODT_EINVAL;
  Calib P;
  for (int k = 0; k < 12; ++k) P.p[k] = p2[k];
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (dtype == DODT_F64)
    project_image_kernel<double><<<blocks, 256, 0, as_stream(stream_)>>>(
        static_cast<const double *>(anchors), n, P, image_h, image_w, tf_order, boxes_norm, boxes_pixels);
  else
    project_image_kernel<float><<<blocks, 256, 0, as_stream(stream_)>>>(
        static_cast<const float *>(anchors), n, P, image_h, image_w, tf_order, boxes_norm, boxes_pixels);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_offset_to_anchor(const void *anchors, int32_t anchors_dtype, const void *offsets,
                          int32_t offsets_dtype, int64_t n, double *out, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || (anchors_dtype != DODT_F32 && anchors_dtype != DODT_F64) ||
      (offsets_dtype != DODT_F32 && offsets_dtype != DODT_F64))
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  if (!anchors || !offsets || !out) return DODT_EINVAL;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  cudaStream_t stream = as_stream(stream_);
  if (anchors_dtype == DODT_F64 && offsets_dtype == DODT_F64)
    offset_to_anchor_kernel<double, double><<<blocks, 256, 0, stream>>>(static_cast<const double *>(anchors), static_cast<const double *>(offsets), n, out);
  else if (anchors_dtype == DODT_F64)
    offset_to_anchor_kernel<double, float><<<blocks, 256, 0, stream>>>(static_cast<const double *>(anchors), static_cast<const float *>(offsets), n, out);
  else if (offsets_dtype == DODT_F64)
    offset_to_anchor_kernel<float, double><<<blocks, 256, 0, stream>>>(static_cast<const float *>(anchors), static_cast<const double *>(offsets), n, out);
  else
    offset_to_anchor_kernel<float, float><<<blocks, 256, 0, stream>>>(static_cast<const float *>(anchors), static_cast<const float *>(offsets), n, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_rpn_decode(const double *anchors, const float *offsets, const int32_t *idx,
                    const int32_t *idx2, const int32_t *count, int64_t n_max,
                    const double bev_extents[4],
                    const double p2[12], int32_t image_h, int32_t image_w, int32_t decode_f32,
                    float *bev_boxes, float *img_boxes, dodt_stream_t stream_) {
  using namespace dodt;
  if (n_max < 0 || n_max > 0x7FFFFFFF || !count || !bev_extents) return DODT_EINVAL;
  if (img_boxes && (!p2 || image_h <= 0 || image_w <= 0)) return DODT_EINVAL;
  if (n_max == 0) return DODT_OK;
  if (!anchors || !offsets || !idx || (!bev_boxes && !img_boxes)) return DODT_EINVAL;
  if (reinterpret_cast<uintptr_t>(bev_boxes) % 16 != 0 || reinterpret_cast<uintptr_t>(img_boxes) % 16 != 0)
    return DODT_EALIGN;
  Calib P;
  for (int k = 0; k < 12; ++k) P.p[k] = p2 ? p2[k] : 0.0;
  rpn_decode_kernel<<<ceil_div(n_max, 128), 128, 0, as_stream(stream_)>>>(
      anchors, offsets, idx, idx2, count, static_cast<int>(n_max), bev_extents[0], bev_extents[1],
      bev_extents[2], bev_extents[3], P, image_h, image_w, decode_f32 ? 1 : 0, bev_boxes, img_boxes);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // extern "C"
