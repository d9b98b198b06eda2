// LiDAR ingest: raw velodyne scan -> camera-frame points inside the image
// frustum, the step in front of S1. Host NumPy in the reference:
//   wavedata/wavedata/tools/core/calib_utils.py:484-523          lidar_to_cam_frame
//   wavedata/wavedata/tools/core/calib_utils.py:394-410          project_to_image
//   wavedata/wavedata/tools/obj_detection/tracking_utils.py:152-203  get_lidar_point_cloud
//       (points with camera z > 0 whose projection lies strictly inside the image, input order)
//   avod/datasets/kitti/kitti_tracking_dataset.py:317-328           point_cloud_transform: the scan of
//       frame t+tau moved into frame t's LiDAR frame, (xyz + trans) @ matrix in float64, stored
//       back into the float32 scan (dodt_lidar_to_camera_aligned; trans / matrix are the host OXTS
//       arithmetic of kitti_tracking_utils.py:129-216)
//
// One pass computes the rectified camera coordinates (float64, as the reference: np.dot of float64
// calibration matrices) and the keep flag of every point; the ordered compaction reuses the
// front end's compact kernels; a gather writes the (3, M) structure-of-arrays cloud S1 consumes.
// np.dot's summation order / FMA use is BLAS-defined, so coordinates agree with NumPy to float64
// rounding noise (the keep decision could differ only for a point whose projection is within that
// noise of the image border).
#include <string.h>

#include "common.cuh"

namespace dodt {
namespace {

struct LidarGeom {
  double m[12];   // rows 0..2 of R0_rect(4x4) . Tr_velo_to_cam(4x4)
  double p[12];   // camera matrix P2
  double im_w, im_h;
  double et[3];   // ego-motion alignment: translation ...
  double er[9];   // ... and 3x3 matrix (row-major), applied as (xyz + et) @ er
  int filter;     // 0: keep every point (im_size not given)
  int ego;        // 1: align the scan first (the result is rounded to float32 like the reference's store)
};

__global__ void __launch_bounds__(256)
lidar_transform(const float4 *__restrict__ velo, long long n, const LidarGeom g,
                double *__restrict__ cam /* (3, n) */, unsigned char *__restrict__ keep,
                float4 *__restrict__ aligned /* [n] or null */) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 v = __ldg(velo + i);
  double x = v.x, y = v.y, z = v.z;
  if (g.ego) {
    // kitti_tracking_dataset.py:326: pc_next[:, :3] = (pc_next[:, :3] + trans) @ matrix — float64
    // arithmetic assigned into the float32 scan
    const double ax = __dadd_rn(x, g.et[0]), ay = __dadd_rn(y, g.et[1]), az = __dadd_rn(z, g.et[2]);
    float al[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      al[j] = __double2float_rn(fma(az, g.er[6 + j], fma(ay, g.er[3 + j], __dmul_rn(ax, g.er[j]))));
    x = al[0]; y = al[1]; z = al[2];
    if (aligned) aligned[i] = make_float4(al[0], al[1], al[2], v.w);
  }
  double c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.m[4 * r], x), __dmul_rn(g.m[4 * r + 1], y)),
                               __dmul_rn(g.m[4 * r + 2], z)), g.m[4 * r + 3]);
  cam[i] = c[0];
  cam[n + i] = c[1];
  cam[2 * n + i] = c[2];
  bool k = true;
  if (g.filter) {
    k = c[2] > 0.0;   // in front of the camera
    double q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      q[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.p[4 * r], c[0]), __dmul_rn(g.p[4 * r + 1], c[1])),
                                 __dmul_rn(g.p[4 * r + 2], c[2])), g.p[4 * r + 3]);
    const double u = __ddiv_rn(q[0], q[2]), w = __ddiv_rn(q[1], q[2]);
    k = k && u > 0.0 && u < g.im_w && w > 0.0 && w < g.im_h;
  }
  keep[i] = k ? 1 : 0;
}

// point_cloud_transform on its own (no rectification wanted): the moved scan only
__global__ void __launch_bounds__(256)
lidar_align(const float4 *__restrict__ velo, long long n, const LidarGeom g, float4 *__restrict__ aligned) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 v = __ldg(velo + i);
  const double ax = __dadd_rn(static_cast<double>(v.x), g.et[0]), ay = __dadd_rn(static_cast<double>(v.y), g.et[1]),
               az = __dadd_rn(static_cast<double>(v.z), g.et[2]);
  float al[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    al[j] = __double2float_rn(fma(az, g.er[6 + j], fma(ay, g.er[3 + j], __dmul_rn(ax, g.er[j]))));
  aligned[i] = make_float4(al[0], al[1], al[2], v.w);
}

template <typename T>
__global__ void __launch_bounds__(256)
lidar_gather(const double *__restrict__ cam, long long n, const int *__restrict__ idx,
             const int *__restrict__ count, long long out_stride, T *__restrict__ out) {
  const long long j = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (j >= n || j >= __ldg(count)) return;
  const long long src = __ldg(idx + j);
#pragma unroll
  for (int r = 0; r < 3; ++r) out[r * out_stride + j] = static_cast<T>(cam[r * n + src]);
}

}  // namespace
}  // namespace dodt

extern "C" {

size_t dodt_lidar_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  // camera coordinates (3n doubles), keep flags, kept indices, the compaction's own scratch
  return 3 * nn * sizeof(double) + ((nn + 255) & ~static_cast<size_t>(255)) + nn * sizeof(int32_t) + 256 +
         dodt_compact_workspace_bytes(n) + 256;
}

int dodt_lidar_to_camera(const float *velo, int64_t n, const double rectified[12],
                         const double p2[12], int32_t image_w, int32_t image_h, void *points,
                         int32_t points_dtype, int64_t row_stride, int32_t *count, void *workspace,
                         size_t workspace_bytes, dodt_stream_t stream_) {
  return dodt_lidar_to_camera_aligned(velo, n, nullptr, nullptr, nullptr, rectified, p2, image_w, image_h, points,
                                      points_dtype, row_stride, count, workspace, workspace_bytes, stream_);
}

int dodt_lidar_to_camera_aligned(const float *velo, int64_t n, const double ego_trans[3],
                                 const double ego_matrix[9], float *aligned, const double rectified[12],
                                 const double p2[12], int32_t image_w, int32_t image_h, void *points,
                                 int32_t points_dtype, int64_t row_stride, int32_t *count, void *workspace,
                                 size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if ((ego_trans == nullptr) != (ego_matrix == nullptr) || (aligned && !ego_trans)) return DODT_EINVAL;
  if (aligned && reinterpret_cast<uintptr_t>(aligned) % 16 != 0) return DODT_EALIGN;
  if (aligned && !points) {
    // alignment only: aligned <- the moved scan; rectified / p2 / points / count / workspace are not used
    if (n < 0 || n > 0x7FFFFFFF || (n > 0 && !velo)) return DODT_EINVAL;
    if (reinterpret_cast<uintptr_t>(velo) % 16 != 0) return DODT_EALIGN;
    if (n == 0) return DODT_OK;
    LidarGeom ga;
    memset(&ga, 0, sizeof(ga));
    for (int k = 0; k < 3; ++k) ga.et[k] = ego_trans[k];
    for (int k = 0; k < 9; ++k) ga.er[k] = ego_matrix[k];
    lidar_align<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream_)>>>(
        reinterpret_cast<const float4 *>(velo), n, ga, reinterpret_cast<float4 *>(aligned));
    DODT_AFTER_LAUNCH();
    return DODT_OK;
  }
  if (n < 0 || n > 0x7FFFFFFF || !rectified || !count || row_stride < n) return DODT_EINVAL;
  if (points_dtype != DODT_F32 && points_dtype != DODT_F64) return DODT_EINVAL;
  const bool filter = image_w > 0 && image_h > 0;
  if (filter && !p2) return DODT_EINVAL;
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    DODT_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), stream));
    return DODT_OK;
  }
  if (!velo || !points) return DODT_EINVAL;
  if (reinterpret_cast<uintptr_t>(velo) % 16 != 0) return DODT_EALIGN;
  if (!workspace || workspace_bytes < dodt_lidar_workspace_bytes(n)) return DODT_ECAPACITY;
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return DODT_EALIGN;
  char *ws = static_cast<char *>(workspace);
  const size_t nn = static_cast<size_t>(n);
  double *cam = reinterpret_cast<double *>(ws);
  size_t off = 3 * nn * sizeof(double);
  unsigned char *keep = reinterpret_cast<unsigned char *>(ws + off);
  off += (nn + 255) & ~static_cast<size_t>(255);
  int32_t *idx = reinterpret_cast<int32_t *>(ws + off);
  off = (off + nn * sizeof(int32_t) + 255) & ~static_cast<size_t>(255);
  void *cws = ws + off;
  LidarGeom g;
  for (int k = 0; k < 12; ++k) { g.m[k] = rectified[k]; g.p[k] = p2 ? p2[k] : 0.0; }
  g.im_w = image_w; g.im_h = image_h; g.filter = filter ? 1 : 0;
  g.ego = ego_trans ? 1 : 0;
  for (int k = 0; k < 3; ++k) g.et[k] = ego_trans ? ego_trans[k] : 0.0;
  for (int k = 0; k < 9; ++k) g.er[k] = ego_matrix ? ego_matrix[k] : 0.0;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  lidar_transform<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(velo), n, g, cam, keep,
                                              reinterpret_cast<float4 *>(aligned));
  DODT_AFTER_LAUNCH();
  const int rc = dodt_compact_mask(keep, n, idx, count, cws, workspace_bytes - off, stream_);
  if (rc != DODT_OK) return rc;
  if (points_dtype == DODT_F64)
    lidar_gather<double><<<blocks, 256, 0, stream>>>(cam, n, idx, count, row_stride, static_cast<double *>(points));
  else
    lidar_gather<float><<<blocks, 256, 0, stream>>>(cam, n, idx, count, row_stride, static_cast<float *>(points));
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // extern "C"
