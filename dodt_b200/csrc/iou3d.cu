// (f3) Tracking-association IoU on the device: all-pairs "3-D IoU" of upright
// boxes that may rotate about the vertical axis, the consumer of the front end's detections.
//
// Reference (paths relative to the Guoxs/DODT checkout):
//   wavedata/wavedata/tools/obj_detection/evaluation.py:44-92   three_d_iou (sphere reject, height
//                                                               overlap x base overlap / union)
//   wavedata/.../evaluation.py:95-128                           height_metrics
//   wavedata/.../evaluation.py:131-161                          get_rotated_3d_bb (corner order)
//   wavedata/.../evaluation.py:164-261                          get_rectangular_metrics
//   called per (track, detection) pair by avod/experiments/video_detection*.py (iou_3d).
// The reference finds the base overlap by rasterising both rectangles at 0.01 m with PIL's polygon
// fill ("minor precision loss due to discretization") inside a Python loop over the pairs. Here one
// thread per pair clips the two rectangles exactly (Sutherland-Hodgman, float64); the result
// differs from the rasterised value by the discretisation only: <= 0.01 in IoU on car-sized boxes
// (tests state 0.02 against outputs of the reference itself, tests/golden/f3_three_d_iou.npz).
#include <math.h>

#include "common.cuh"

namespace dodt {
namespace {

struct Pt { double x, z; };

// the four base corners in the order of get_rotated_3d_bb
__device__ __forceinline__ void corners_of(const double *b, Pt c[4]) {
  const double cs = cos(b[0]), sn = sin(b[0]);
  const double hl = b[1] / 2, hw = b[3] / 2;
  const double xc[4] = {hl, hl, -hl, -hl};
  const double zc[4] = {hw, -hw, -hw, hw};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c[k].x = cs * xc[k] + sn * zc[k] + b[4];
    c[k].z = -sn * xc[k] + cs * zc[k] + b[6];
  }
}

// exact area of the intersection of two convex quadrilaterals
__device__ double quad_intersection_area(const Pt a[4], const Pt b_in[4]) {
  Pt clip[4];
  double area2 = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) area2 += b_in[k].x * b_in[(k + 1) & 3].z - b_in[(k + 1) & 3].x * b_in[k].z;
#pragma unroll
  for (int k = 0; k < 4; ++k) clip[k] = area2 < 0 ? b_in[3 - k] : b_in[k];   // counter-clockwise
  Pt poly[8], next[8];
  int n = 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) poly[k] = a[k];
  for (int e = 0; e < 4 && n > 0; ++e) {
    const Pt p0 = clip[e], p1 = clip[(e + 1) & 3];
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const Pt p = poly[i], q = poly[i + 1 == n ? 0 : i + 1];
      const double sp = (p1.x - p0.x) * (p.z - p0.z) - (p1.z - p0.z) * (p.x - p0.x);
      const double sq = (p1.x - p0.x) * (q.z - p0.z) - (p1.z - p0.z) * (q.x - p0.x);
      if (sp >= 0 && m < 8) next[m++] = p;
      if ((sp >= 0) != (sq >= 0) && m < 8) {
        const double t = sp / (sp - sq);
        next[m].x = p.x + t * (q.x - p.x);
        next[m].z = p.z + t * (q.z - p.z);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) poly[i] = next[i];
  }
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    const Pt p = poly[i], q = poly[i + 1 == n ? 0 : i + 1];
    s += p.x * q.z - q.x * p.z;
  }
  return fabs(s) / 2;
}

__global__ void __launch_bounds__(128)
three_d_iou_kernel(const double *__restrict__ a, int na, const double *__restrict__ b, int nb,
                   double *__restrict__ iou) {
  const long long t = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (t >= static_cast<long long>(na) * nb) return;
  const double *p = a + (t / nb) * 7, *q = b + (t % nb) * 7;
  double pa[7], qb[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { pa[k] = __ldg(p + k); qb[k] = __ldg(q + k); }
  // spheres around the boxes (evaluation.py:60-74)
  const double da = sqrt(pa[1] * pa[1] + pa[2] * pa[2] + pa[3] * pa[3]) / 2;
  const double db = sqrt(qb[1] * qb[1] + qb[2] * qb[2] + qb[3] * qb[3]) / 2;
  const double dx = qb[4] - pa[4], dy = qb[5] - pa[5], dz = qb[6] - pa[6];
  double r = 0.0;
  if (da + db >= sqrt(dx * dx + dy * dy + dz * dz)) {
    // height_metrics: y points down and ty is the bottom face
    const double h_int = fmax(0.0, fmin(pa[5], qb[5]) - fmax(pa[5] - pa[2], qb[5] - qb[2]));
    Pt ca[4], cb[4];
    corners_of(pa, ca);
    corners_of(qb, cb);
    const double base = fmin(100.0, quad_intersection_area(ca, cb));   // evaluation.py:254 caps at 100
    const double inter = h_int * base;
    r = inter / (pa[1] * pa[2] * pa[3] + qb[1] * qb[2] * qb[3] - inter);
  }
  iou[t] = r;
}

}  // namespace
}  // namespace dodt

extern "C" int dodt_three_d_iou_matrix(const double *boxes_a, int32_t na, const double *boxes_b, int32_t nb,
                                       double *iou, dodt_stream_t stream_) {
  using namespace dodt;
  if (na < 0 || nb < 0) return DODT_EINVAL;
  if (na == 0 || nb == 0) return DODT_OK;
  if (!boxes_a || !boxes_b || !iou) return DODT_EINVAL;
  const long long total = static_cast<long long>(na) * nb;
  if ((total + 127) / 128 > 0x7FFFFFFFll) return DODT_ECAPACITY;
  three_d_iou_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, as_stream(stream_)>>>(boxes_a, na, boxes_b,
                                                                                                 nb, iou);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}
