/*
 * dodt_fe.h — C ABI of libdodt_fe.so, the B200 (sm_100a) implementation of the DODT/AVOD
 * per-frame proposal front end:
 *
 *   S1  LiDAR points -> BEV height-slice + density maps   (dodt_bev_slices)
 *   S2  integral image + empty-anchor filter               (dodt_integral_image_2d,
 *                                                           dodt_anchor_filter_2d, dodt_map_to_index)
 *   S3  crop_and_resize of NHWC feature maps               (dodt_crop_and_resize)
 *   S4  inter-frame feature correlation (forward)          (dodt_correlation)
 *   S5  greedy axis-aligned NMS                            (dodt_nms)
 *
 * Conventions (precedent: the reference's only FFI, the ctypes binding of
 * wavedata/wavedata/tools/core/lib/src/integral_images_3d.cpp:66-77 loaded by
 * wavedata/wavedata/tools/core/integral_image.py:96-120 — caller-allocated outputs, plain
 * pointers and sizes, no ownership transfer):
 *
 *  - every data pointer is a DEVICE pointer unless the parameter comment says "host";
 *  - the caller owns every buffer including workspaces; the library never allocates or frees
 *    device memory and keeps no state between calls (re-entrant, one host thread per GPU);
 *  - every compute entry point enqueues work on `stream` (a cudaStream_t passed as void*)
 *    and returns without synchronising; all of them are CUDA-graph capturable;
 *  - return value: DODT_OK (0) or a negative DODT_E* code; nothing is printed;
 *  - there is no CPU fallback: without a CUDA device the compute entry points return DODT_ECUDA.
 *
 * Reference paths are relative to the Guoxs/DODT checkout.
 */
#ifndef DODT_FE_H_
#define DODT_FE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DODT_FE_VERSION 201 /* major*100 + minor */

typedef void *dodt_stream_t; /* cudaStream_t */

enum dodt_status {
  DODT_OK = 0,
  DODT_EINVAL = -1,    /* bad argument value (NULL pointer, negative size, even kernel_size, ...) */
  DODT_ESHAPE = -2,    /* shapes inconsistent / output would be empty                              */
  DODT_ECAPACITY = -3, /* workspace too small or size beyond what the packed keys can index        */
  DODT_ECUDA = -4,     /* a CUDA runtime call failed; see dodt_last_cuda_error()                   */
  DODT_EALIGN = -5     /* pointer not aligned as documented                                        */
};

enum dodt_dtype { DODT_F32 = 0, DODT_F64 = 1 };

#define DODT_MAX_SLICES 15
#define DODT_MAX_DENSITY_LUT 64
/* layout of the int32 stats block written by dodt_bev_slices */
#define DODT_BEV_STATS_LEN 24
#define DODT_BEV_STAT_DENSITY 16  /* number of points inside the density slice                    */
#define DODT_BEV_STAT_OCC 17      /* number of points inside the occupancy (anchor filter) slice  */
#define DODT_BEV_STAT_TOUCHED 18  /* number of (map, cell) pairs touched by at least one point    */
#define DODT_BEV_STAT_OVERFLOW 19 /* !=0: touched list overflowed the workspace (result invalid)  */
#define DODT_BEV_STAT_OOB 20      /* number of points binned outside the grid (no-filter mode)    */

const char *dodt_strerror(int code);
const char *dodt_last_cuda_error(void); /* thread-local text of the last CUDA failure */
int dodt_version(void);
/* number of kernels this library has launched in this process since load
 * (memset nodes are not counted); used by bench.py for its "gpu_launches" claim */
int64_t dodt_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * S1 — BEV height-slice + density maps, with the S2 occupancy grid produced in the same pass.
 * Replaces avod/core/bev_generators/bev_slices.py:33-150 (BevSlices.generate_bev), which calls
 * avod/datasets/kitti/kitti_utils.py:81-109 (create_slice_filter),
 * wavedata/wavedata/tools/obj_detection/obj_utils.py:453-500 (get_point_filter),
 * wavedata/wavedata/tools/core/voxel_grid_2d.py:43-160 (VoxelGrid2D.voxelize_2d),
 * wavedata/wavedata/tools/core/geometry_utils.py:25-40 (dist_to_plane) and
 * avod/core/bev_generators/bev_generator.py:23-41 (_create_density_map); the occupancy grid
 * replaces avod/datasets/kitti/kitti_utils.py:212-277 (_apply_slice_filter +
 * create_sliced_voxel_grid_2d, leaf_layout_2d).
 * ---------------------------------------------------------------------------------------- */
typedef struct dodt_bev_params {
  double plane[4];     /* ground plane a,b,c,d                                                    */
  double extents[6];   /* x_min,x_max,y_min,y_max,z_min,z_max (open intervals, obj_utils.py:474)  */
  double voxel_size;   /* the fp32-rounded proto value, e.g. (double)0.1f                          */
  double height_lo;    /* BevSlices config (bev_slices.py:24-31)                                  */
  double height_hi;
  int32_t num_slices;  /* 0..DODT_MAX_SLICES                                                      */
  int32_t filter_mode; /* 1: slice filters as in generate_bev; 0: every point belongs to slice 0
                          and to the density map (plain VoxelGrid2D.voxelize_2d of given points)  */
  double occ_lo;       /* occupancy slice above the plane, kitti_utils.py:212-213 (0.2 .. 2.0)    */
  double occ_hi;
  int32_t density_lut_len;                   /* entries used in density_lut, <= 64                */
  int32_t reserved;
  double density_lut[DODT_MAX_DENSITY_LUT];  /* density value for n points, n < lut_len; 1.0 for
                                                n >= lut_len (min(1, log(n+1)/norm))              */
} dodt_bev_params;

/* grid[0..5] = nx, ny, nz, min_x, min_y, min_z in voxel units: floor(ext_min/voxel) and
 * ceil(ext_max/voxel - 1) as voxel_grid_2d.py:125-130,142-143 (ny is the number of y bins, used
 * only for the winner key; the grid itself is collapsed along y). Host only. */
int dodt_bev_grid(const double extents[6], double voxel_size, int32_t grid[6]);

size_t dodt_bev_workspace_bytes(int64_t n_points, int32_t num_slices, int32_t nx, int32_t nz);

/*
 * pts       : (3, n) structure-of-arrays as the reference passes it (point_cloud (3,N)); element
 *             type pts_dtype; row r starts at pts + r*row_stride elements.
 * maps      : out float32 [(num_slices+1), nz, nx]; maps[s] is height map s already rotated as
 *             bev_slices.py:115-116 (row 0 = far z), maps[num_slices] is the density map.
 * occ       : out uint8 [nx, nz] (1 = VOXEL_FILLED, 0 = VOXEL_EMPTY) or NULL to skip.
 * stats     : out int32 [DODT_BEV_STATS_LEN]; stats[s] = points in height slice s.
 * winner_idx: optional out int32 [num_slices, nz, nx] — index of the winning point per cell
 *             (-1 = empty); NULL to skip (parity/diagnostic output, integer exact).
 * counts    : optional out int32 [nz, nx] — points per cell of the density slice; NULL to skip.
 * n_dev     : optional device int32 — only the first min(n, *n_dev) points are used (a count
 *             produced on the device, e.g. by dodt_lidar_to_camera); NULL: all n.
 */
int dodt_bev_slices(const void *pts, int32_t pts_dtype, int64_t n, const int32_t *n_dev,
                    int64_t row_stride,
                    const dodt_bev_params *params /* host */, float *maps, uint8_t *occ,
                    int32_t *stats, int32_t *winner_idx, int32_t *counts, void *workspace,
                    size_t workspace_bytes, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * S2 — integral image of the occupancy grid and the O(1) box-sum anchor filter.
 * Replaces wavedata/wavedata/tools/core/integral_image_2d.py:7-87 (IntegralImage2D),
 * wavedata/wavedata/tools/core/voxel_grid_2d.py:162-186 (map_to_index) and
 * avod/core/anchor_filter.py:64-119 (get_empty_anchor_filter_2d).
 * ---------------------------------------------------------------------------------------- */
size_t dodt_integral_workspace_bytes(int32_t nx, int32_t nz);

/* ii: out int32 [(nx+1), (nz+1)], zero first row/column (integral_image_2d.py:30-37) */
int dodt_integral_image_2d(const uint8_t *occ, int32_t nx, int32_t nz, int32_t *ii,
                           void *workspace, size_t workspace_bytes, dodt_stream_t stream);

/* Banded form of the integral image for consumers that add the band offsets themselves (the fused
 * kernel below): ONE launch writes the band-local image ii_local (zero row / column included) and,
 * by the last band to finish, the exclusive band offsets; the full image of
 * dodt_integral_image_2d is ii_local[X][Z] + bandoff[(X-1) / band_rows][Z-1] for X, Z >= 1.
 * workspace: dodt_integral_banded_workspace_bytes, ZERO before its first use (the call leaves its
 * counter zero); *bandoff_out (host pointer, optional) receives the device address of the offsets
 * [ceil(nx / band_rows), nz] inside the workspace. */
size_t dodt_integral_banded_workspace_bytes(int32_t nx, int32_t nz);
int32_t dodt_integral_band_rows(void);
int dodt_integral_image_2d_banded(const uint8_t *occ, int32_t nx, int32_t nz, int32_t *ii_local,
                                  void *workspace, size_t workspace_bytes, int32_t **bandoff_out,
                                  dodt_stream_t stream);

/* S2 for the frame stream in one launch: the empty-anchor filter (avod/core/anchor_filter.py:64-119)
 * on float64 anchors, the ordered compaction of the kept anchors (dt_rpn_model.py:952-958), the
 * gather of their precomputed BEV / image crop boxes [n,4] and RPN scores [n], and the decoded,
 * BEV-projected boxes of their regressed anchors (dt_rpn_model.py:573-591; same bits as
 * dodt_rpn_decode). ii: the full integral image (bandoff NULL) or the band-local one with bandoff /
 * band_rows from dodt_integral_image_2d_banded. Outputs: keep [n] u8, kept_idx [n] ascending,
 * n_kept [1] (device), k_* [n,4] / [n] at the compacted positions (any of the k_* with its source
 * may be NULL). workspace: dodt_anchor_filter_fused_workspace_bytes(n), ZERO before its first use
 * (every call leaves it zero). Same results as dodt_anchor_filter_2d + dodt_compact_mask +
 * dodt_gather_rows_multi + dodt_rpn_decode. decode_f32: as in dodt_rpn_decode. */
/* anchors = NULL with grid != NULL: the anchors are the grid of dodt_grid_anchors(grid->...) and are
 * evaluated from the anchor index inside the kernel instead of being read ("anchors
 * become a function of index"); n must be the size of that grid. Same bits as with the table. */
typedef struct dodt_anchor_grid {
  double extents[6];      /* x_min, x_max, y_min, y_max, z_min, z_max                              */
  double stride[2];       /* x, z                                                                   */
  double plane[4];        /* a, b, c, d with b != 0                                                 */
  double sizes[16 * 3];   /* n_sizes rows (l, w, h)                                                 */
  int32_t n_sizes;        /* 1..16                                                                  */
  int32_t reserved;
} dodt_anchor_grid;
size_t dodt_anchor_filter_fused_workspace_bytes(int64_t n);
int dodt_anchor_filter_fused(const double *anchors, const dodt_anchor_grid *grid /* host */, int64_t n,
                             const int32_t *ii, const int32_t *bandoff,
                             int32_t band_rows, int32_t nx, int32_t nz, int32_t min_x, int32_t min_z,
                             double voxel_size, double density_threshold, const float *anchor_bev_boxes,
                             const float *anchor_img_boxes, const float *rpn_scores, const float *rpn_offsets,
                             const double bev_extents[4], int32_t decode_f32, uint8_t *keep, int32_t *kept_idx,
                             int32_t *n_kept, float *k_bev_boxes, float *k_img_boxes, float *k_scores,
                             float *k_rpn_boxes, void *workspace, size_t workspace_bytes, dodt_stream_t stream);

/* coords (n,2) [x,z] of dtype -> idx int32 (n,2); division in the coordinate dtype, truncation
 * toward zero, shift by the grid minimum, clip to [0, ndiv] (voxel_grid_2d.py:182-184) */
int dodt_map_to_index(const void *coords, int32_t dtype, int64_t n, double voxel_size,
                      int32_t min_x, int32_t min_z, int32_t nx, int32_t nz, int32_t *idx,
                      dodt_stream_t stream);

/* anchors (n,6) [x,y,z,dim_x,dim_y,dim_z] row-major of dtype; keep: out uint8 [n];
 * scores: optional out int32 [n] box sums (NULL to skip) */
int dodt_anchor_filter_2d(const void *anchors, int32_t dtype, int64_t n, const int32_t *ii,
                          int32_t nx, int32_t nz, int32_t min_x, int32_t min_z, double voxel_size,
                          double density_threshold, uint8_t *keep, int32_t *scores,
                          dodt_stream_t stream);

/* Ordered compaction of a keep mask — what `anchors[anchor_filter]` (NumPy boolean indexing,
 * avod/core/models/dt_rpn_model.py:952-958) does on the host in the reference. idx: out int32 [n]
 * (first *count entries valid, ascending), count: out int32 [1] on the device. */
size_t dodt_compact_workspace_bytes(int64_t n);
int dodt_compact_mask(const uint8_t *keep, int64_t n, int32_t *idx, int32_t *count,
                      void *workspace, size_t workspace_bytes, dodt_stream_t stream);
/* dst[i, :] = src[idx[i], :] for i < *count (rows of `width` float32); count is a device pointer */
int dodt_gather_rows(const float *src, int32_t width, const int32_t *idx, const int32_t *count,
                     int64_t n_max, float *dst, dodt_stream_t stream);
/* the same for up to DODT_MAX_GATHER arrays in ONE launch (specs is a host array of device ptrs) */
#define DODT_MAX_GATHER 8
typedef struct dodt_gather_spec {
  const float *src; /* [m, width] */
  float *dst;       /* [n_max, width] */
  int32_t width;
  int32_t reserved;
} dodt_gather_spec;
int dodt_gather_rows_multi(const dodt_gather_spec *specs /* host */, int32_t n_specs,
                           const int32_t *idx, const int32_t *count, int64_t n_max,
                           dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LiDAR ingest in front of S1: wavedata tracking_utils.py:152-203
 * (get_lidar_point_cloud) = calib_utils.py:484-523 (lidar_to_cam_frame) + the image-frustum filter
 * through calib_utils.py:394-410 (project_to_image).
 * velo: device float32 [n, 4] (x, y, z, intensity: the KITTI .bin layout), 16-byte aligned.
 * rectified: host, rows 0..2 of R0_rect(4x4) . Tr_velo_to_cam(4x4), row-major 3x4; p2: host 3x4.
 * image_w / image_h > 0: keep points with camera z > 0 whose projection is strictly inside the
 * image; 0: keep all. points: out (3, *count) structure of arrays of points_dtype, row r at
 * points + r*row_stride, input order kept; count: out device int32.
 * ---------------------------------------------------------------------------------------- */
size_t dodt_lidar_workspace_bytes(int64_t n);
int dodt_lidar_to_camera(const float *velo, int64_t n, const double rectified[12],
                         const double p2[12], int32_t image_w, int32_t image_h, void *points,
                         int32_t points_dtype, int64_t row_stride, int32_t *count, void *workspace,
                         size_t workspace_bytes, dodt_stream_t stream);
/* The same with the ego-motion alignment of DODT's frame pairs in front of it
 * (avod/datasets/kitti/kitti_tracking_dataset.py:317-328, point_cloud_transform): the scan of frame
 * t+tau is moved into frame t's LiDAR frame, xyz <- float32((xyz + ego_trans) @ ego_matrix)
 * (float64 arithmetic stored back into the float32 scan, as the reference's assignment does), then
 * rectified / filtered as above. ego_trans[3], ego_matrix[9] (row-major 3x3): host, from the OXTS
 * records (kitti_tracking_utils.py:189-207; dodt_b200.lidar.coordinate_transform); both NULL = no
 * alignment. aligned: optional out, device float32 [n, 4] (moved x, y, z, intensity unchanged),
 * 16-byte aligned, or NULL. With aligned != NULL and points == NULL only the alignment is done
 * (point_cloud_transform on its own): rectified, p2, count and workspace are not used then. */
int dodt_lidar_to_camera_aligned(const float *velo, int64_t n, const double ego_trans[3],
                                 const double ego_matrix[9], float *aligned, const double rectified[12],
                                 const double p2[12], int32_t image_w, int32_t image_h, void *points,
                                 int32_t points_dtype, int64_t row_stride, int32_t *count, void *workspace,
                                 size_t workspace_bytes, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Anchor geometry between the stages — host NumPy in the reference.
 * ---------------------------------------------------------------------------------------- */
/* shape[4] = {nz, nx, n_sizes, 2 rotations} of the anchor grid that
 * avod/core/anchor_generators/grid_anchor_3d_generator.py:39-108 (tile_anchors_3d) tiles over
 * extents (x_min,x_max,y_min,y_max,z_min,z_max) with stride (x, z); the grid has their product
 * anchors, ordered z row (far to near), x column, size, rotation. Host only. */
int dodt_grid_anchor_shape(const double extents[6], const double stride[2], int32_t n_sizes,
                           int32_t shape[4]);
/* anchors: out device float64 [N, 6] = box_3d_to_anchor(tile_anchors_3d(...))
 * (avod/core/box_3d_encoder.py:85-132), bit-identical to NumPy. sizes: host [n_sizes, 3] (l, w, h),
 * n_sizes <= 16; plane: host a, b, c, d with b != 0. */
int dodt_grid_anchors(const double extents[6], const double *sizes, int32_t n_sizes,
                      const double stride[2], const double plane[4], double *anchors,
                      dodt_stream_t stream);
/* avod/core/anchor_projector.py:13-69 (project_to_bev). anchors [n, 6] of dtype; bev_extents
 * x_min, x_max, z_min, z_max (host); outputs float32 [n, 4] (either may be NULL): corners
 * normalised by the extent ranges, and in metres from the top-left of the map. tf_order = 0:
 * [x1, z1, x2, z2] as the reference returns them; 1: [z1, x1, z2, x2], the order
 * tf.image.crop_and_resize wants (anchor_projector.py:254-273 reorder_projected_boxes). */
int dodt_project_to_bev(const void *anchors, int32_t dtype, int64_t n, const double bev_extents[4],
                        int32_t tf_order, float *boxes_norm, float *boxes_metres,
                        dodt_stream_t stream);
/* avod/core/anchor_projector.py:72-156 (project_to_image_space): min / max of the eight cuboid
 * corners projected with the 3x4 camera matrix p2 (host, row-major); outputs float32 [n, 4]
 * [x1, y1, x2, y2] (tf_order: [y1, x1, y2, x2]) normalised by the image size, and in pixels. */
int dodt_project_to_image_space(const void *anchors, int32_t dtype, int64_t n, const double p2[12],
                                int32_t image_h, int32_t image_w, int32_t tf_order,
                                float *boxes_norm, float *boxes_pixels, dodt_stream_t stream);
/* avod/core/anchor_encoder.py:99-150 (offset_to_anchor): out float64 [n, 6]. */
int dodt_offset_to_anchor(const void *anchors, int32_t anchors_dtype, const void *offsets,
                          int32_t offsets_dtype, int64_t n, double *out, dodt_stream_t stream);

/* The decode chain between the RPN head and NMS / the second-stage crops
 * (avod/core/models/dt_rpn_model.py:573-591,618-660) for the anchors a device-side filter kept:
 * for i < *count: j = idx[idx2 ? idx2[i] : i]; regressed = offset_to_anchor(anchors[j], offsets[j]);
 * bev_boxes[i] = normalised BEV corners [z1, x1, z2, x2]; img_boxes[i] = normalised image corners
 * [y1, x1, y2, x2] (either output may be NULL). anchors [m, 6] float64, offsets [m, 6] float32.
 * idx2 (optional) selects among the kept anchors, e.g. the NMS survivors: the image projection
 * (eight corners in float64) is only needed for those. n_max bounds i.
 * decode_f32 = 0: the NumPy branches of offset_to_anchor / project_to_bev, float64 throughout,
 * rounded to float32 at the end. 1: their tf.Tensor branches as the reference's inference graph runs
 * them on float32 placeholders (dt_rpn_model.py:568-591): anchors rounded to float32, one float32
 * operation per TF op, exp / log correctly rounded (TF's GPU expf / logf are within 2 ulp of that);
 * the image projection continues in float64 from the float32 regressed anchor. */
int dodt_rpn_decode(const double *anchors, const float *offsets, const int32_t *idx,
                    const int32_t *idx2, const int32_t *count, int64_t n_max,
                    const double bev_extents[4],
                    const double p2[12], int32_t image_h, int32_t image_w, int32_t decode_f32,
                    float *bev_boxes, float *img_boxes, dodt_stream_t stream);

/* Multi-GPU hand-off (the per-frame rows the reference writes with np.savetxt,
 * avod/core/dt_evaluator.py:1098-1147, and that a sharded run gathers once per shard): appends the
 * detections keep[0 .. *n_keep) of one frame to a block of fixed-size lists.
 * rows [max_frames, max_det, 6] f32 = box (4), score, index into boxes; counts [max_frames];
 * frame_ids [max_frames, 2] = the two int32 at frame_id (sequence, frame; NULL: -2, row);
 * cursor: device int32, the next free row (incremented by the call; rows past max_frames are
 * dropped, the cursor keeps counting). row_io: optional device int32 — with rewrite == 0 the row the
 * frame was given is stored there; with rewrite != 0 the frame's list REPLACES row *row_io (the
 * cursor is untouched): how a frame whose RPN NMS had to be resumed corrects its entry.
 * All device pointers; graph-capturable. */
int dodt_emit_detections(const float *boxes, const float *scores, const int32_t *keep,
                         const int32_t *n_keep, int32_t max_det, const int32_t *frame_id,
                         float *rows, int32_t *counts, int32_t *frame_ids, int32_t *cursor,
                         int32_t max_frames, int32_t *row_io, int32_t rewrite, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Tracking-association IoU: three_d_iou of
 * wavedata/wavedata/tools/obj_detection/evaluation.py:44-92 for ALL pairs of two box sets, as the
 * greedy linker of avod/experiments/video_detection*.py needs it per frame (tracks x detections).
 * boxes_a [na,7], boxes_b [nb,7] f64 rows [ry, l, h, w, tx, ty, tz] (y down, ty = bottom face);
 * iou [na,nb] f64. The base overlap is clipped exactly; the reference rasterises it at 0.01 m
 * (evaluation.py:164-261), which moves an IoU by at most ~0.01. All device pointers. */
int dodt_three_d_iou_matrix(const double *boxes_a, int32_t na, const double *boxes_b, int32_t nb,
                            double *iou, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * S3 — tf.image.crop_and_resize (TensorFlow 1.3.0 core/kernels/crop_and_resize_op.cc, bilinear),
 * called at avod/core/models/dt_rpn_model.py:418-428 and dt_avod_model.py:253-273.
 * image [batch,H,W,C] f32 NHWC; boxes [n,4] normalised [y1,x1,y2,x2]; box_ind [n] (rows whose
 * index is outside [0,batch) are left untouched, as TF does; NULL means all zero);
 * crops out [n,crop_h,crop_w,C]. n_dev: optional device int32 — only the first min(n, *n_dev)
 * boxes are processed (a box count produced on the device, e.g. by dodt_compact_mask).
 * ---------------------------------------------------------------------------------------- */
int dodt_crop_and_resize(const float *image, int32_t batch, int32_t height, int32_t width,
                         int32_t channels, const float *boxes, const int32_t *box_ind, int64_t n,
                         const int32_t *n_dev, int32_t crop_h, int32_t crop_w,
                         float extrapolation_value, float *crops, dodt_stream_t stream);

/* Several feature maps cropped with per-map boxes in ONE launch (the reference issues one
 * tf.image.crop_and_resize per map: BEV + image at dt_rpn_model.py:418-428, BEV + image + corr at
 * dt_avod_model.py:253-273). All maps share batch, box_ind (NULL = zeros), n, n_dev, crop size. */
#define DODT_MAX_CROP_MAPS 4
typedef struct dodt_crop_spec {
  const float *image; /* [batch, height, width, channels] */
  const float *boxes; /* [n, 4] */
  float *crops;       /* [n, crop_h, crop_w, channels] */
  int32_t height, width, channels;
  int32_t reserved;
} dodt_crop_spec;
int dodt_crop_and_resize_multi(const dodt_crop_spec *specs /* host */, int32_t n_specs,
                               int32_t batch, const int32_t *box_ind, int64_t n,
                               const int32_t *n_dev, int32_t crop_h, int32_t crop_w,
                               float extrapolation_value, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * S4 — FlowNet correlation forward. Replaces the TF custom op
 * avod/core/ops/correlation/correlation_op.cc:53-62 (REGISTER_OP "Correlation"),
 * correlation_kernel.cc:26-124 (shape math, padding) and correlation_kernel.cu.cc:21-119
 * (CorrelateData) + pad.cu.cc:14-74 (PadData; padding is never materialised here).
 * a, b [batch,H,W,C] f32 NHWC -> out [batch,out_h,out_w,(2*(max_disp/stride_2)+1)^2].
 * ---------------------------------------------------------------------------------------- */
int dodt_correlation_out_shape(int32_t height, int32_t width, int32_t kernel_size,
                               int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                               int32_t pad, int32_t out_hwc[3]);
int dodt_correlation(const float *a, const float *b, int32_t batch, int32_t height, int32_t width,
                     int32_t channels, int32_t kernel_size, int32_t max_displacement,
                     int32_t stride_1, int32_t stride_2, int32_t pad, float *out,
                     dodt_stream_t stream);
/* The same, for a caller that runs other work next to it: max_ctas > 0 caps the number of
 * (persistent) CTAs of the launch, e.g. one per SM instead of two, which leaves half of every SM's
 * registers and shared memory to kernels of other streams. The block scheduler does not promise one
 * CTA per SM for such a launch: when it co-locates some of them the launch takes up to 1.5x, which is
 * why the frame-stream runner leaves the cap off by default. 0 = no cap. */
int dodt_correlation_shared(const float *a, const float *b, int32_t batch, int32_t height,
                            int32_t width, int32_t channels, int32_t kernel_size,
                            int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                            int32_t pad, float *out, int32_t max_ctas, dodt_stream_t stream);

/* Frame-stream form of S4 (DODT correlates EVERY pair of adjacent key frames of a sequence,
 * avod/core/models/dt_rpn_model.py:324-331 called once per sample pair by the inference loop):
 * outs[j] = correlation(maps[j], maps[j + 1]) for j = 0 .. n_maps - 2, bit-identical to n_maps - 1
 * calls of dodt_correlation with batch 1. Up to DODT_CORR_STREAM_MAX_PAIRS pairs share one launch
 * whose tiles are interleaved over the pairs, so that the map two neighbouring pairs have in common
 * is read from HBM once. maps, outs: HOST arrays of n_maps / n_maps - 1 device pointers
 * ([1,H,W,C] inputs, [1,out_h,out_w,out_c] outputs). max_ctas as for dodt_correlation_shared. */
#define DODT_CORR_STREAM_MAX_PAIRS 8
int dodt_correlation_stream(const float *const *maps, int32_t n_maps, float *const *outs,
                            int32_t height, int32_t width, int32_t channels, int32_t kernel_size,
                            int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                            int32_t pad, int32_t max_ctas, dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * S4 backward — gradients of the correlation w.r.t. both inputs. Replaces the
 * TF custom op "CorrelationGrad": avod/core/corr_layers/correlation.py:30-48
 * (@tf.RegisterGradient), correlation_grad_kernel.cc:28-151 and correlation_grad_kernel.cu.cc:20-189
 * (CorrelateDataBackward0 / CorrelateDataBackward1). Same attributes as the forward op.
 * grad [batch,out_h,out_w,out_c] f32 (the gradient of the forward output), a, b [batch,H,W,C] ->
 * grad_a, grad_b [batch,H,W,C] (either may be NULL: that gradient is skipped).
 * workspace: not used any more (dodt_correlation_grad_workspace_bytes returns 0: the displacement
 * flip that grad_b needs is gathered per tile inside the kernel); the parameters stay in the ABI and
 * may be NULL / 0.
 * ---------------------------------------------------------------------------------------- */
size_t dodt_correlation_grad_workspace_bytes(int32_t batch, int32_t height, int32_t width,
                                             int32_t channels, int32_t kernel_size,
                                             int32_t max_displacement, int32_t stride_1,
                                             int32_t stride_2, int32_t pad);
int dodt_correlation_grad(const float *grad, const float *a, const float *b, int32_t batch,
                          int32_t height, int32_t width, int32_t channels, int32_t kernel_size,
                          int32_t max_displacement, int32_t stride_1, int32_t stride_2, int32_t pad,
                          float *grad_a, float *grad_b, void *workspace, size_t workspace_bytes,
                          dodt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * S5 — tf.image.non_max_suppression (TensorFlow 1.3.0 core/kernels/non_max_suppression_op.cc),
 * called at avod/core/models/dt_rpn_model.py:587-591 and dt_avod_model.py:609-613.
 * boxes [n,4] f32 (any corner order), scores [n] f32; keep: out int32 [max_out] (entries past
 * *n_keep are set to -1); n_keep: out int32 [1] on the device; suppress iff IoU > iou_threshold.
 * Equal scores are ordered by ascending index (TF leaves that order unspecified).
 * n_dev: optional device int32 — the candidate count is min(n, *n_dev) (entries past it are
 * ignored). Candidates are visited in windows of DODT_NMS_WINDOW in score order, ordered lazily
 * in chunks of DODT_NMS_CHUNK_WINDOWS windows. max_windows: 0 = enqueue enough windows for any
 * input; k > 0 = enqueue at most k windows (fixed launch count for CUDA graphs); if the selection
 * is not complete after them n_keep[1] is 0, else 1. n_keep: out int32 [2].
 * first_window: 0 = new selection; a positive multiple of DODT_NMS_CHUNK_WINDOWS = continue the
 * selection a previous, incomplete call left in the same workspace / keep / n_keep, starting at
 * that window (same boxes, scores, n, max_out, threshold).
 * ---------------------------------------------------------------------------------------- */
#define DODT_NMS_WINDOW 1536
#define DODT_NMS_CHUNK_WINDOWS 2
size_t dodt_nms_workspace_bytes(int64_t n);
/* byte offset, inside the workspace, of the diagnostic block of the last solved window:
 * int32 {n_kept, done, ticket, sweeps} then uint64 globaltimer[6] {kernel start, IoU tiles done,
 * masks in shared memory, recurrence solved, emitted, -} */
size_t dodt_nms_state_offset(int64_t n);
int dodt_nms(const float *boxes, const float *scores, int64_t n, const int32_t *n_dev,
             int32_t max_out, float iou_threshold, int32_t first_window, int32_t max_windows,
             int32_t *keep, int32_t *n_keep, void *workspace, size_t workspace_bytes,
             dodt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DODT_FE_H_ */
