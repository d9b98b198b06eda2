"""CPU oracle of the DODT proposal front end — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under dodt_b200/ does. It restates in NumPy what the reference
computes for each stage, each function citing the reference file:line it follows (paths relative
to the Guoxs/DODT checkout):

  (ingest) lidar_to_cam_frame / lidar_in_camera_view   (wavedata calib_utils / tracking_utils)
           oxts_coordinate_transform / point_cloud_transform (ego-motion alignment of DODT's frame pairs)
  S1  bev_slices / voxelize_2d / point_filter / dist_to_plane / density
  S2  integral_image_2d / map_to_index / empty_anchor_filter_2d
  S3  crop_and_resize       (TensorFlow 1.3.0 core/kernels/crop_and_resize_op.cc — NOT vendored
                             in the reference; published algorithm restated)
  S4  correlation           (avod/core/ops/correlation/*.cu.cc)
  S5  non_max_suppression   (TensorFlow 1.3.0 core/kernels/non_max_suppression_op.cc — NOT
                             vendored; published algorithm restated)

Pinning status
  S1, S2: PINNED. tests/test_oracle_vs_reference.py runs these functions against the reference's
          own Python (imported from /root/reference through oracle/ref_shim.py) where that
          checkout exists, and tests/golden/*.npz holds outputs generated from the reference by
          oracle/make_golden.py (real KITTI fixture frames + the reference's unit-test vectors).
  S4:     the reference op is GPU-only TensorFlow 1.3 C++ with no asserting test
          (corr_layers/correlation_test.py prints only) and cannot be built here (needs TF
          headers). PARITY UNPINNED by reference outputs; the restatement follows the kernel's
          indexing and summation order literally and is cross-checked against an independent
          torch formulation (unfold/shift) in tests.
  S3, S5: TensorFlow 1.3.0 is a third-party dependency absent from the checkout and this image.
          PARITY UNPINNED by reference outputs; cross-checked against independent
          implementations (torch.nn.functional.grid_sample align_corners=True, torchvision.ops.nms)
          whose outputs are frozen in tests/golden/, and box_list_ops_test.py:86-98's IoU vectors.
"""
import numpy as np

# ------------------------------------------------------------------------------------------
# LiDAR ingest
# ------------------------------------------------------------------------------------------


def lidar_to_cam_frame(xyz_lidar, r0_rect, tr_velodyne_to_cam):
    """wavedata/wavedata/tools/core/calib_utils.py:484-523: N x 3 lidar -> N x 3 rectified camera."""
    r0 = np.pad(np.asarray(r0_rect, dtype=np.float64), ((0, 1), (0, 1)), 'constant')
    r0[3, 3] = 1
    tf = np.pad(np.asarray(tr_velodyne_to_cam, dtype=np.float64), ((0, 1), (0, 0)), 'constant')
    tf[3, 3] = 1
    xyz1 = np.append(xyz_lidar, np.ones(len(xyz_lidar)).reshape(-1, 1), axis=1)
    return np.dot(np.dot(r0, tf), xyz1.T)[0:3].T


def lidar_in_camera_view(velo, r0_rect, tr_velodyne_to_cam, p2, im_size=None):
    """wavedata/wavedata/tools/obj_detection/tracking_utils.py:152-203 (get_lidar_point_cloud) on
    an already loaded scan: velo N x 4 (x, y, z, intensity) -> (3, M) camera-frame points that lie
    in front of the camera and project strictly inside an image of im_size = [w, h]."""
    pts = lidar_to_cam_frame(np.asarray(velo)[:, :3], r0_rect, tr_velodyne_to_cam)
    if not im_size:
        return pts.T
    pts = pts[pts[:, 2] > 0]
    pc = pts.T
    uvw = np.dot(np.asarray(p2), np.append(pc, np.ones((1, pc.shape[1])), axis=0))   # calib_utils.py:394-410
    u, v = uvw[0] / uvw[2], uvw[1] / uvw[2]
    keep = (u > 0) & (u < im_size[0]) & (v > 0) & (v < im_size[1])
    return pts[keep].T


def oxts_coordinate_transform(cur, nxt):
    """avod/datasets/kitti/kitti_tracking_dataset.py:300-315 (coordinate_transform) with the Oxts
    arithmetic of kitti_tracking_utils.py:141-216. cur / nxt: the first six values of the two oxts
    lines (latitude, longitude [deg], altitude, roll, pitch, yaw [rad]). Returns (trans [3],
    matrix [3, 3], yaw difference)."""
    lat1, lon1 = cur[0] * np.pi / 180.00, cur[1] * np.pi / 180.00
    lat2, lon2 = nxt[0] * np.pi / 180.00, nxt[1] * np.pi / 180.00
    a = lat2 - lat1
    b = lon2 - lon1
    d = abs(2 * 6378137.0 * np.arcsin(np.sqrt(np.power(np.sin(a / 2), 2) +
                                               np.cos(lat1) * np.cos(lat2) * np.power(np.sin(b / 2), 2))))
    d_roll, d_pitch, d_yaw = cur[3] - nxt[3], cur[4] - nxt[4], cur[5] - nxt[5]
    trans = np.array([d * np.cos(d_yaw), d * np.sin(d_yaw), d * np.sin(d_pitch)])
    c, s = np.cos(d_pitch), np.sin(d_pitch)
    rz = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])       # "rotz": pitch difference, x-z plane
    c, s = np.cos(d_roll), np.sin(d_roll)
    rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    c, s = np.cos(d_yaw), np.sin(d_yaw)
    ry = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])       # "roty": yaw difference, x-y plane
    return trans, rz @ rx @ ry, d_yaw


def point_cloud_transform(pc_next, trans, matrix):
    """kitti_tracking_dataset.py:317-328: pc_next (4, N) float32 scan of frame t+tau (x, y, z,
    intensity rows) -> the same array with xyz = float32((xyz + trans) @ matrix)."""
    out = np.array(pc_next, dtype=np.float32).T.copy()
    out[:, :3] = (out[:, :3] + trans) @ matrix              # float64 product stored as float32
    return out.T


# ------------------------------------------------------------------------------------------
# S1
# ------------------------------------------------------------------------------------------


def point_filter(point_cloud, extents, ground_plane=None, offset_dist=2.0):
    """wavedata/wavedata/tools/obj_detection/obj_utils.py:453-500 (get_point_filter).

    point_cloud (3, N); extents [[x0,x1],[y0,y1],[z0,z1]] as OPEN intervals; with a plane the point
    must also satisfy dot(plane - [0,0,0,offset], [x,y,z,1]) < 0.
    """
    pc = np.asarray(point_cloud)
    inside = np.ones(pc.shape[1], dtype=bool)
    for axis in range(3):
        inside &= (pc[axis] > extents[axis][0]) & (pc[axis] < extents[axis][1])
    if ground_plane is None:
        return inside
    shifted = np.array(ground_plane) + [0, 0, 0, -offset_dist]
    homogeneous = np.vstack([pc, np.ones(pc.shape[1])])
    return inside & (np.dot(shifted, homogeneous) < 0)


def slice_filter(point_cloud, extents, ground_plane, lo, hi):
    """avod/datasets/kitti/kitti_utils.py:81-109 (create_slice_filter): xor of the two filters."""
    return np.logical_xor(point_filter(point_cloud, extents, ground_plane, hi),
                          point_filter(point_cloud, extents, ground_plane, lo))


def dist_to_plane(plane, points):
    """wavedata/wavedata/tools/core/geometry_utils.py:25-40."""
    a, b, c, d = plane
    pts = np.array(points)
    return (a * pts[:, 0] + b * pts[:, 1] + c * pts[:, 2] + d) / np.sqrt(a ** 2 + b ** 2 + c ** 2)


def voxelize_2d(pts, voxel_size, extents=None, ground_plane=None):
    """wavedata/wavedata/tools/core/voxel_grid_2d.py:43-160 (VoxelGrid2D.voxelize_2d).

    The reference lexsorts by (x-bin, z-bin, y-bin) and keeps the first row of every (x,z) group
    (np.unique(return_index)); group sizes are the point counts. Returned as a dict:
      voxel_coords (V,3) absolute bins with y zeroed, order = ascending (x, z)
      first (V,) index INTO `pts` of the point that represents each voxel
      heights (V,), counts (V,), min_coord (3,), num_divisions (3,) int32
    Raises ValueError exactly where the reference does.
    """
    pts = np.asarray(pts)
    if pts.shape[1] != 3:
        raise ValueError("Points have the wrong shape: {}".format(pts.shape))
    bins = np.floor(pts / voxel_size).astype(np.int32)
    order = np.lexsort((bins[:, 1], bins[:, 2], bins[:, 0]))   # stable: ties keep file order
    sb = bins[order]
    new_group = np.ones(len(sb), dtype=bool)
    new_group[1:] = (sb[1:, 0] != sb[:-1, 0]) | (sb[1:, 2] != sb[:-1, 2])
    starts = np.flatnonzero(new_group)
    counts = np.diff(np.append(starts, len(sb)))
    first = order[starts]
    coords = sb[starts].copy()
    coords[:, 1] = 0
    if ground_plane is None:
        heights = pts[first, 1]
    else:
        heights = dist_to_plane(ground_plane, pts[first])
    if extents is not None:
        ext_t = np.array(extents).transpose()
        if ext_t.shape != (2, 3):
            raise ValueError("Extents are the wrong shape {}".format(ext_t.shape))
        min_coord = np.floor(ext_t[0] / voxel_size)
        max_coord = np.ceil((ext_t[1] / voxel_size) - 1)
        min_coord[1] = 0
        max_coord[1] = 0
        if not (min_coord <= np.amin(coords, axis=0)).all():
            raise ValueError("Extents are smaller than min_voxel_coord")
        if not (max_coord >= np.amax(coords, axis=0)).all():
            raise ValueError("Extents are smaller than max_voxel_coord")
    else:
        min_coord = np.amin(coords, axis=0)
        max_coord = np.amax(coords, axis=0)
    num_divisions = ((max_coord - min_coord) + 1).astype(np.int32)
    return dict(voxel_coords=coords, first=first, heights=heights, counts=counts,
                min_coord=min_coord, num_divisions=num_divisions,
                voxel_indices=(coords - min_coord).astype(int))


def leaf_layout_2d(vox):
    """voxel_grid_2d.py:152-160: -1 empty / 0 filled, shape (nx, 1, nz)."""
    layout = -1.0 * np.ones(vox["num_divisions"].astype(int))
    layout[vox["voxel_indices"][:, 0], 0, vox["voxel_indices"][:, 2]] = 0
    return layout


def bev_slices(point_cloud, ground_plane, area_extents, voxel_size, height_lo, height_hi,
               num_slices, norm_value=np.log(16), return_debug=False):
    """avod/core/bev_generators/bev_slices.py:33-150 (BevSlices.generate_bev) and
    avod/core/bev_generators/bev_generator.py:23-41 (_create_density_map).

    Returns {'height_maps': [S x (nz, nx) float64], 'density_map': (nz, nx) float64}; with
    return_debug also 'winner' [S x (nz,nx) int64 index into the cloud, -1 empty], 'counts'
    (nz,nx) int64, 'slice_counts' (S+1,).
    """
    point_cloud = np.asarray(point_cloud)
    all_points = np.transpose(point_cloud)
    hpd = (height_hi - height_lo) / num_slices if num_slices else 0.0
    height_maps, winners, slice_counts = [], [], []
    all_idx = np.arange(all_points.shape[0])
    for s in range(num_slices):
        lo = height_lo + s * hpd
        hi = lo + hpd
        mask = slice_filter(point_cloud, area_extents, ground_plane, lo, hi)
        pts = all_points[mask]
        src = all_idx[mask]
        slice_counts.append(len(pts))
        if len(pts) > 1:
            vox = voxelize_2d(pts, voxel_size, area_extents, ground_plane)
            win_src = src[vox["first"]]
        else:
            # bev_slices.py:86-99 — 0 or 1 points: a single origin point is voxelised instead
            vox = voxelize_2d(np.zeros((1, 3)), voxel_size, area_extents, ground_plane)
            win_src = np.array([-1])
        vi = vox["voxel_indices"][:, [0, 2]]
        hmap = np.zeros((vox["num_divisions"][0], vox["num_divisions"][2]))
        hmap[vi[:, 0], vi[:, 1]] = np.asarray(vox["heights"] - lo) / hpd
        wmap = -np.ones(hmap.shape, dtype=np.int64)
        wmap[vi[:, 0], vi[:, 1]] = win_src
        height_maps.append(np.flip(hmap.transpose(), axis=0))       # bev_slices.py:115-116
        winners.append(np.flip(wmap.transpose(), axis=0))
    dmask = slice_filter(point_cloud, area_extents, ground_plane, height_lo, height_hi)
    dpts = all_points[dmask]
    slice_counts.append(len(dpts))
    dvox = voxelize_2d(dpts, voxel_size, area_extents, ground_plane)   # IndexError-free only if >0
    dvi = dvox["voxel_indices"][:, [0, 2]]
    dmap = np.zeros((dvox["num_divisions"][0], dvox["num_divisions"][2]))
    dmap[dvi[:, 0], dvi[:, 1]] = np.minimum(1.0, np.log(dvox["counts"] + 1) / norm_value)
    cmap = np.zeros(dmap.shape, dtype=np.int64)
    cmap[dvi[:, 0], dvi[:, 1]] = dvox["counts"]
    out = {"height_maps": height_maps, "density_map": np.flip(dmap.transpose(), axis=0)}
    if return_debug:
        out["winner"] = winners
        out["counts"] = np.flip(cmap.transpose(), axis=0)
        out["slice_counts"] = np.array(slice_counts)
    return out


def occupancy_grid(point_cloud, ground_plane, area_extents, voxel_size, height_lo=0.2,
                   height_hi=2.0):
    """avod/datasets/kitti/kitti_utils.py:212-277 (_apply_slice_filter + create_sliced_voxel_grid_2d):
    leaf_layout_2d + 1 as a (nx, nz) uint8 grid (1 filled)."""
    point_cloud = np.asarray(point_cloud)
    mask = slice_filter(point_cloud, area_extents, ground_plane, height_lo, height_hi)
    vox = voxelize_2d(np.transpose(point_cloud)[mask], voxel_size, area_extents, ground_plane)
    return (np.squeeze(leaf_layout_2d(vox)) + 1).astype(np.uint8), vox


# ------------------------------------------------------------------------------------------
# S2
# ------------------------------------------------------------------------------------------


def integral_image_2d(img):
    """wavedata/wavedata/tools/core/integral_image_2d.py:17-37: padded double cumsum."""
    img = np.asarray(img)
    if img.ndim != 2:
        raise ValueError("Not a 2D image for integral image: input dim {}".format(img.ndim))
    out = np.zeros((img.shape[0] + 1, img.shape[1] + 1))
    out[1:, 1:] = np.cumsum(np.cumsum(img, 0), 1)
    return out


def integral_query(ii, boxes):
    """integral_image_2d.py:39-87: boxes (4, N) uint32 [x1, z1, x2, z2]."""
    boxes = np.asarray(boxes)
    if boxes.shape[0] != 4:
        raise ValueError("Incorrect number of dimensions for query")
    if boxes.dtype != np.uint32:
        raise TypeError("boxes must be type of np.uint32")
    lim = np.array([ii.shape[0], ii.shape[1], ii.shape[0], ii.shape[1]]) - 1
    b = np.minimum(boxes, lim.reshape(4, -1)).astype(np.uint32)
    x1, z1, x2, z2 = b
    return ii[x2, z2] + ii[x1, z1] - ii[x2, z1] - ii[x1, z2]


def map_to_index(coords, voxel_size, min_coord_2d, num_div_2d):
    """wavedata/wavedata/tools/core/voxel_grid_2d.py:162-186: divide in the dtype of `coords`,
    truncate toward zero (np.int32), shift, clip to [0, ndiv]."""
    coords = np.asarray(coords)
    idx = np.int32(coords / voxel_size) - np.asarray(min_coord_2d, dtype=np.float64)
    idx[:, 0] = np.clip(idx[:, 0], 0, num_div_2d[0])
    idx[:, 1] = np.clip(idx[:, 1], 0, num_div_2d[1])
    return idx


def empty_anchor_filter_2d(anchors, occ, voxel_size, min_coord_2d, density_threshold=1,
                           return_scores=False):
    """avod/core/anchor_filter.py:64-119 (get_empty_anchor_filter_2d).

    anchors (N,6) [x,y,z,dx,dy,dz]; occ (nx,nz) {0,1} = leaf_layout_2d + 1.
    """
    anchors = np.asarray(anchors)
    a2 = anchors[:, [0, 2, 3, 5]]
    ii = integral_image_2d(np.asarray(occ, dtype=np.float64))
    tl = np.zeros([len(a2), 2]).astype(np.float32)
    br = np.zeros([len(a2), 2]).astype(np.float32)
    tl[:, 0] = a2[:, 0] - (a2[:, 2] / 2.)
    tl[:, 1] = a2[:, 1] - (a2[:, 3] / 2.)
    br[:, 0] = a2[:, 0] + (a2[:, 2] / 2.)
    br[:, 1] = a2[:, 1] + (a2[:, 3] / 2.)
    nd = (occ.shape[0], occ.shape[1])
    box = np.zeros([len(a2), 4]).astype(np.uint32)
    box[:, :2] = map_to_index(tl, voxel_size, min_coord_2d, nd)
    box[:, 2:] = map_to_index(br, voxel_size, min_coord_2d, nd)
    scores = integral_query(ii, box.T)
    keep = scores >= density_threshold
    return (keep, scores) if return_scores else keep


# ------------------------------------------------------------------------------------------
# S3
# ------------------------------------------------------------------------------------------


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    """TensorFlow 1.3.0 tensorflow/core/kernels/crop_and_resize_op.cc, CropAndResize<CPUDevice>
    (bilinear). image [B,H,W,C] f32, boxes [N,4] f32 normalised [y1,x1,y2,x2], box_ind [N] int32,
    crop_size (ch, cw). Every fp32 operation is rounded individually, as the C++ loop does.
    Rows whose box_ind is outside [0,B) are left at zero (TF skips them)."""
    f32 = np.float32
    image = np.asarray(image, dtype=f32)
    boxes = np.asarray(boxes, dtype=f32)
    box_ind = np.asarray(box_ind)
    B, H, W, C = image.shape
    ch, cw = int(crop_size[0]), int(crop_size[1])
    N = boxes.shape[0]
    out = np.zeros((N, ch, cw, C), dtype=f32)
    if N == 0:
        return out
    y1, x1, y2, x2 = (boxes[:, k] for k in range(4))

    def coords(lo, hi, size, crop):
        sm1 = f32(size - 1)
        if crop > 1:
            scale = (hi - lo) * sm1 / f32(crop - 1)                     # (N,)
            steps = np.arange(crop, dtype=f32)
            return (lo * sm1)[:, None] + steps[None, :] * scale[:, None]   # (N, crop)
        mid = (0.5 * (lo + hi).astype(np.float64)) * np.float64(size - 1)  # double, then narrowed
        return mid.astype(f32)[:, None]

    in_y = coords(y1, y2, H, ch)         # (N, ch)
    in_x = coords(x1, x2, W, cw)         # (N, cw)
    ok_y = ~((in_y < 0) | (in_y > f32(H - 1)))
    ok_x = ~((in_x < 0) | (in_x > f32(W - 1)))
    ty = np.floor(np.where(ok_y, in_y, 0)).astype(np.int64)
    by = np.ceil(np.where(ok_y, in_y, 0)).astype(np.int64)
    lx = np.floor(np.where(ok_x, in_x, 0)).astype(np.int64)
    rx = np.ceil(np.where(ok_x, in_x, 0)).astype(np.int64)
    yl = (in_y - ty.astype(f32)).astype(f32)
    xl = (in_x - lx.astype(f32)).astype(f32)
    valid_b = (box_ind >= 0) & (box_ind < B)
    bsel = np.where(valid_b, box_ind, 0).astype(np.int64)
    bi = bsel[:, None, None]
    TY, BY = ty[:, :, None], by[:, :, None]
    LX, RX = lx[:, None, :], rx[:, None, :]
    tl = image[bi, TY, LX]               # (N, ch, cw, C)
    tr = image[bi, TY, RX]
    bl = image[bi, BY, LX]
    br = image[bi, BY, RX]
    XL = xl[:, None, :, None]
    YL = yl[:, :, None, None]
    top = tl + (tr - tl) * XL
    bot = bl + (br - bl) * XL
    val = (top + (bot - top) * YL).astype(f32)
    ok = (ok_y[:, :, None] & ok_x[:, None, :])[..., None]
    val = np.where(ok, val, f32(extrapolation_value)).astype(f32)
    out[valid_b] = val[valid_b]
    return out


# ------------------------------------------------------------------------------------------
# S4
# ------------------------------------------------------------------------------------------


def correlation_out_shape(H, W, kernel_size, max_displacement, stride_1, stride_2, pad):
    """avod/core/ops/correlation/correlation_kernel.cc:39-57 / correlation_op.cc:27-46."""
    if kernel_size % 2 == 0:
        raise ValueError("kernel_size must be odd")                  # correlation_kernel.cc:23
    border = max_displacement + (kernel_size - 1) // 2
    oh = int(np.ceil(np.float32(H + 2 * pad - 2 * border) / np.float32(stride_1)))
    ow = int(np.ceil(np.float32(W + 2 * pad - 2 * border) / np.float32(stride_1)))
    if oh < 1:
        raise ValueError("Neighborhood and kernel don't fit in input height.")
    if ow < 1:
        raise ValueError("Neighborhood and kernel don't fit in input width.")
    r = max_displacement // stride_2
    return oh, ow, (2 * r + 1) ** 2


def correlation(input_a, input_b, kernel_size=1, max_displacement=20, stride_1=1, stride_2=2,
                padding=20):
    """avod/core/ops/correlation/correlation_kernel.cu.cc:21-119 (CorrelateData) on inputs padded
    as pad.cu.cc:14-74 does; signature of avod/core/corr_layers/correlation.py:7.

    Summation order of the kernel is kept: lane t of the 32-thread block accumulates, over the
    kernel window (j, i) and channels ch = t, t+32, ..., its products in fp32 (fused multiply-adds,
    as the compiled reference kernel does); thread 0 then adds
    the 32 partial sums in lane order and divides by kernel_size^2 * C.
    """
    f32 = np.float32
    a = np.asarray(input_a, dtype=f32)
    b = np.asarray(input_b, dtype=f32)
    if a.ndim != 4 or b.ndim != 4:
        raise ValueError("inputs must have rank 4")
    if a.shape != b.shape:
        raise ValueError("inputs must have the same shape")
    N, H, W, C = a.shape
    ks, md, s1, s2, pad = kernel_size, max_displacement, stride_1, stride_2, padding
    oh, ow, oc = correlation_out_shape(H, W, ks, md, s1, s2, pad)
    r = md // s2
    wn = 2 * r + 1
    # extra margin so that displaced windows that leave the padded image read zeros
    ex = max(0, r * s2 + md + ks - pad)
    P = pad + ex
    ap = np.zeros((N, H + 2 * P, W + 2 * P, C), dtype=f32)
    bp = np.zeros_like(ap)
    ap[:, P:P + H, P:P + W] = a
    bp[:, P:P + H, P:P + W] = b
    ys = np.arange(oh) * s1 + md + ex          # padded-with-margin coordinates of the patch corner
    xs = np.arange(ow) * s1 + md + ex
    out = np.zeros((N, oh, ow, oc), dtype=f32)
    lanes = 32
    for k in range(oc):
        s2o = (k % wn - r) * s2
        s2p = (k // wn - r) * s2
        partial = np.zeros((N, oh, ow, lanes), dtype=f32)
        for j in range(ks):
            for i in range(ks):
                pa = ap[:, ys[:, None] + j, xs[None, :] + i]                     # (N,oh,ow,C)
                pb = bp[:, ys[:, None] + j + s2p, xs[None, :] + i + s2o]
                # `sum[ch_off] += patch * b` (correlation_kernel.cu.cc:93) compiles to one FFMA (nvcc's
                # default -fmad=true; pinned by tests/golden/s4_reference_kernel.npz): the fp32 product
                # is exact in float64, so one float64 add + one rounding to fp32 is the fused result
                prod = pa.astype(np.float64) * pb.astype(np.float64)
                for c0 in range(0, C, lanes):
                    chunk = prod[..., c0:c0 + lanes]
                    w = chunk.shape[-1]
                    partial[..., :w] = (chunk + partial[..., :w].astype(np.float64)).astype(f32)
        total = np.zeros((N, oh, ow), dtype=f32)
        for t in range(lanes):
            total = total + partial[..., t]
        out[..., k] = total / f32(ks * ks * C)
    return out


# ------------------------------------------------------------------------------------------
# S5
# ------------------------------------------------------------------------------------------


def iou_matrix_row(box, area, boxes, areas):
    """ComputeIOU of non_max_suppression_op.cc (fp32): IoU of one normalised box with many."""
    f32 = np.float32
    iy0 = np.maximum(box[0], boxes[:, 0])
    ix0 = np.maximum(box[1], boxes[:, 1])
    iy1 = np.minimum(box[2], boxes[:, 2])
    ix1 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(iy1 - iy0, f32(0)) * np.maximum(ix1 - ix0, f32(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area + areas - inter)
    iou = np.where((area <= 0) | (areas <= 0), f32(0), iou)
    return iou.astype(f32)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5):
    """TensorFlow 1.3.0 tensorflow/core/kernels/non_max_suppression_op.cc (DoNonMaxSuppressionOp).

    boxes [N,4] f32 (any corner order), scores [N]; returns int32 indices, at most
    min(max_output_size, N). A candidate is dropped iff IoU > iou_threshold with a selected box.
    Equal scores: ascending index (stable sort; TF's std::sort order is unspecified)."""
    f32 = np.float32
    boxes = np.asarray(boxes, dtype=f32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=f32).reshape(-1)
    n = boxes.shape[0]
    out_size = min(int(max_output_size), n)
    order = np.argsort(-scores, kind="stable")
    nb = np.stack([np.minimum(boxes[:, 0], boxes[:, 2]), np.minimum(boxes[:, 1], boxes[:, 3]),
                   np.maximum(boxes[:, 0], boxes[:, 2]), np.maximum(boxes[:, 1], boxes[:, 3])], 1)
    areas = ((nb[:, 2] - nb[:, 0]) * (nb[:, 3] - nb[:, 1])).astype(f32)
    thr = f32(iou_threshold)
    sel = np.zeros(out_size, dtype=np.int64)
    k = 0
    for cand in order:
        if k >= out_size:
            break
        if k:
            s = sel[:k]
            if (iou_matrix_row(nb[cand], areas[cand], nb[s], areas[s]) > thr).any():
                continue
        sel[k] = cand
        k += 1
    return sel[:k].astype(np.int32)


# ------------------------------------------------------------------------------------------
# (f3) tracking-association IoU
# ------------------------------------------------------------------------------------------


def rotated_bb_corners(box):
    """wavedata/.../obj_detection/evaluation.py:117-161 (get_rotated_3d_bb) for one box
    [ry, l, h, w, tx, ty, tz]: the four base corners (x[4], z[4])."""
    ry, l, _, w, tx, _, tz = (float(v) for v in box)
    c, s = np.cos(ry), np.sin(ry)
    xc = l / 2 * np.array([1.0, 1.0, -1.0, -1.0])
    zc = w / 2 * np.array([1.0, -1.0, -1.0, 1.0])
    return c * xc + s * zc + tx, -s * xc + c * zc + tz


def _clip_polygon(poly, a, b):
    """Sutherland-Hodgman step: the part of convex polygon `poly` [(x, z)] on the left of (or on)
    the directed line a -> b."""
    out = []
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        sq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def rectangle_intersection_area(box, other):
    """EXACT area of the intersection of the two boxes' rotated bases. The reference
    (evaluation.py:164-261, get_rectangular_metrics) rasterises both rectangles at 0.01 m with
    PIL's polygon fill and counts common pixels ("minor precision loss due to discretization"),
    capped at 100 m^2; this is the quantity it approximates."""
    xa, za = rotated_bb_corners(box)
    xb, zb = rotated_bb_corners(other)
    poly = list(zip(xa, za))
    clip = list(zip(xb, zb))
    # orientation of the clip rectangle (the corner order of get_rotated_3d_bb is clockwise in x-z)
    area2 = sum(clip[i][0] * clip[(i + 1) % 4][1] - clip[(i + 1) % 4][0] * clip[i][1] for i in range(4))
    if area2 < 0:
        clip = clip[::-1]
    for i in range(4):
        poly = _clip_polygon(poly, clip[i], clip[(i + 1) % 4])
        if not poly:
            return 0.0
    a = 0.0
    for i in range(len(poly)):
        a += poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
    return min(100.0, abs(a) / 2.0)


def three_d_iou(box, boxes):
    """wavedata/.../obj_detection/evaluation.py:44-92 (three_d_iou) with the base intersection
    computed exactly instead of by rasterisation. box [ry, l, h, w, tx, ty, tz]; boxes [n, 7]."""
    box = np.asarray(box, dtype=np.float64)
    boxes = np.asarray(boxes, dtype=np.float64)
    single = boxes.ndim == 1
    if single:
        boxes = boxes[None]
    box_diag = np.sqrt(box[1] ** 2 + box[2] ** 2 + box[3] ** 2) / 2
    boxes_diag = np.sqrt(boxes[:, 1] ** 2 + boxes[:, 2] ** 2 + boxes[:, 3] ** 2) / 2
    dist = np.sqrt(((boxes[:, 4:7] - box[4:7]) ** 2).sum(axis=1))
    iou = np.zeros(len(boxes))
    for i in np.flatnonzero(box_diag + boxes_diag >= dist):
        o = boxes[i]
        # height_metrics (evaluation.py:95-128): y is down, ty is the BOTTOM of a box
        h_int = max(0.0, min(box[5], o[5]) - max(box[5] - box[2], o[5] - o[2]))
        inter = h_int * rectangle_intersection_area(box, o)
        union = box[1] * box[2] * box[3] + o[1] * o[2] * o[3] - inter
        iou[i] = inter / union
    return iou[0] if len(iou) == 1 else iou


def track_iou(detections, sigma_l, sigma_h, sigma_iou, t_min, score_fn):
    """avod/experiments/video_detection.py:235-277 (track_iou), loop for loop, with the pair score
    (cal_transformed_ious there) passed in: score_fn(last detection of a track, detection).
    detections: per frame a list of dicts with at least 'scores' and 'frame_id'."""
    tracks_active, tracks_finished = [], []
    for detections_frame in detections:
        if detections_frame == []:
            continue
        dets = [det for det in detections_frame if det['scores'] >= sigma_l]
        updated_tracks = []
        for track in tracks_active:
            if len(dets) > 0:
                ious = [score_fn(track['trajectory'][-1], x) for x in dets]
                best = int(np.argmax(ious))
                if ious[best] > sigma_iou:
                    track['trajectory'].append(dets[best])
                    track['max_score'] = max(track['max_score'], dets[best]['scores'])
                    updated_tracks.append(track)
                    del dets[best]
            if len(updated_tracks) == 0 or track is not updated_tracks[-1]:
                if track['max_score'] >= sigma_h and len(track['trajectory']) >= t_min:
                    tracks_finished.append(track)
        new_tracks = [{'trajectory': [det], 'max_score': det['scores'], 'start_frame': det['frame_id']}
                      for det in dets]
        tracks_active = updated_tracks + new_tracks
    tracks_finished += [t for t in tracks_active if t['max_score'] >= sigma_h and len(t['trajectory']) >= t_min]
    return tracks_finished
