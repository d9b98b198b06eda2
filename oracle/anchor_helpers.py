"""Host-side (NumPy) mirror of the reference's anchor helpers — TEST INFRASTRUCTURE (oracle/).

These restate the small per-frame array maths that sits between the stages of the front end in the
reference; tests and the CPU baseline use them to produce what the reference's host code would feed
the stages, and tests/test_oracle_vs_reference.py pins them to the live reference. The product
computes the same things on the device (dodt_b200/csrc/anchors.cu) and never imports this module.

  tile_anchors_3d         avod/core/anchor_generators/grid_anchor_3d_generator.py:39-108
  box_3d_to_anchor        avod/core/box_3d_encoder.py:85-132
  project_to_bev          avod/core/anchor_projector.py:13-69
  project_to_image_space  avod/core/anchor_projector.py:72-156 (+ wavedata calib_utils.py:394-410)
  offset_to_anchor        avod/core/anchor_encoder.py:99-150
  offset_to_anchor_tf32 / project_to_bev_tf32   the tf.Tensor branches of the two, float32 op by op
  reorder_projected_boxes avod/core/anchor_projector.py:254-273
"""
import numpy as np

from dodt_b200.constants import CAR_ANCHOR_SIZES, KITTI_P2  # noqa: F401  (re-exported)


def tile_anchors_3d(area_extents, anchor_3d_sizes, anchor_stride, ground_plane):
    """Grid of box_3d anchors N x [x, y, z, l, w, h, ry]: z rows (far to near) x x columns x sizes
    x two rotations (0, pi/2), placed on the ground plane."""
    sizes = np.asarray(anchor_3d_sizes)
    rotations = np.asarray([0, np.pi / 2.0])
    xs = np.array(np.arange(area_extents[0][0] + anchor_stride[0] / 2.0, area_extents[0][1],
                            step=anchor_stride[0]), dtype=np.float32)
    zs = np.array(np.arange(area_extents[2][1] - anchor_stride[1] / 2.0, area_extents[2][0],
                            step=-anchor_stride[1]), dtype=np.float32)
    grid = np.stack(np.meshgrid(xs, zs, np.arange(len(sizes)), np.arange(len(rotations))),
                    axis=4).reshape(-1, 4)
    a, b, c, d = ground_plane
    x, z = grid[:, 0], grid[:, 1]
    y = -(a * x + c * z + d) / b
    out = np.zeros((len(grid), 7))
    out[:, 0:3] = np.stack((x, y, z), axis=1)
    out[:, 3:6] = sizes[np.asarray(grid[:, 2], np.int32)]
    out[:, 6] = rotations[np.asarray(grid[:, 3], np.int32)]
    return out


def box_3d_to_anchor(boxes_3d):
    """[x, y, z, l, w, h, ry] -> [x, y, z, dim_x, dim_y, dim_z] by projecting l, w on the axes."""
    b = np.asarray(boxes_3d).reshape(-1, 7)
    out = np.zeros((len(b), 6))
    out[:, [0, 1, 2]] = b[:, [0, 1, 2]]
    l, w, h, ry = b[:, [3]], b[:, [4]], b[:, [5]], b[:, [6]]
    cos_ry, sin_ry = np.abs(np.cos(ry)), np.abs(np.sin(ry))
    out[:, [3]] = l * cos_ry + w * sin_ry
    out[:, [4]] = h
    out[:, [5]] = w * cos_ry + l * sin_ry
    return out


def project_to_bev(anchors, bev_extents):
    """-> (corners [x1, z1, x2, z2] in metres from the top-left of the BEV map, same normalised)."""
    a = np.asarray(anchors)
    x, z, hx, hz = a[:, 0], a[:, 2], a[:, 3] / 2.0, a[:, 5] / 2.0
    x_min, x_max = bev_extents[0][0], bev_extents[0][1]
    z_min, z_max = bev_extents[1][0], bev_extents[1][1]
    corners = np.stack([x - hx, z_max - (z + hz), x + hx, z_max - (z - hz)], axis=1)
    corners = corners - [x_min, z_min, x_min, z_min]
    rng = [x_max - x_min, z_max - z_min, x_max - x_min, z_max - z_min]
    return corners, corners / rng


def project_to_image_space(anchors, stereo_calib_p2, image_shape):
    """-> (corners [x1, y1, x2, y2] px, normalised corners), float32, from the 8 cuboid corners."""
    a = np.asarray(anchors)
    if a.shape[1] != 6:
        raise ValueError("Invalid shape for anchors {}, should be (N, 6)".format(a.shape[1]))
    x, y, z, dx, dy, dz = (a[:, k] for k in range(6))
    hx, hz = dx / 2., dz / 2.
    xc = np.array([x + hx, x + hx, x - hx, x - hx, x + hx, x + hx, x - hx, x - hx]).T.reshape(1, -1)
    yc = np.array([y, y, y, y, y - dy, y - dy, y - dy, y - dy]).T.reshape(1, -1)
    zc = np.array([z + hz, z - hz, z - hz, z + hz, z + hz, z - hz, z - hz, z + hz]).T.reshape(1, -1)
    pts = np.vstack([xc, yc, zc])
    p2d = np.dot(stereo_calib_p2, np.append(pts, np.ones((1, pts.shape[1])), axis=0))
    u = (p2d[0] / p2d[2]).reshape(-1, 8)
    v = (p2d[1] / p2d[2]).reshape(-1, 8)
    corners = np.vstack([u.min(axis=1), v.min(axis=1), u.max(axis=1), v.max(axis=1)]).T
    norm = corners / [image_shape[1], image_shape[0], image_shape[1], image_shape[0]]
    return np.array(corners, dtype=np.float32), np.array(norm, dtype=np.float32)


def offset_to_anchor(anchors, offsets):
    a, o = np.asarray(anchors), np.asarray(offsets)
    return np.stack((o[:, 0] * a[:, 3] + a[:, 0], o[:, 1] * a[:, 4] + a[:, 1],
                     o[:, 2] * a[:, 5] + a[:, 2], np.exp(np.log(a[:, 3]) + o[:, 3]),
                     np.exp(np.log(a[:, 4]) + o[:, 4]), np.exp(np.log(a[:, 5]) + o[:, 5])), axis=1)


def _f32_exp(x):
    """Correctly rounded float32 exp (evaluated in float64, rounded once)."""
    return np.exp(x.astype(np.float64)).astype(np.float32)


def _f32_log(x):
    return np.log(x.astype(np.float64)).astype(np.float32)


def offset_to_anchor_tf32(anchors, offsets):
    """The tf.Tensor branch of offset_to_anchor (anchor_encoder.py:118-139) as the inference graph
    runs it (dt_rpn_model.py:568-572): anchors and offsets are float32 tensors, every TF op is one
    float32 operation. exp / log are correctly rounded here; TF's GPU kernels are within 2 ulp."""
    a, o = np.asarray(anchors).astype(np.float32), np.asarray(offsets).astype(np.float32)
    pos = [(o[:, k] * a[:, 3 + k]) + a[:, k] for k in range(3)]
    dim = [_f32_exp(_f32_log(a[:, 3 + k]) + o[:, 3 + k]) for k in range(3)]
    out = np.stack(pos + dim, axis=1)
    assert out.dtype == np.float32
    return out


def project_to_bev_tf32(anchors, bev_extents):
    """The tf.Tensor branch of project_to_bev (anchor_projector.py:13-69) on float32 anchors: the
    extents are Python floats, so their differences are formed in float64 and enter the graph as
    float32 constants. -> (corners, normalised corners) [x1, z1, x2, z2], float32."""
    a = np.asarray(anchors)
    assert a.dtype == np.float32
    f = np.float32
    x, z, hx, hz = a[:, 0], a[:, 2], a[:, 3] / f(2.0), a[:, 5] / f(2.0)
    x_min, x_max = bev_extents[0][0], bev_extents[0][1]
    z_min, z_max = bev_extents[1][0], bev_extents[1][1]
    corners = np.stack([x - hx, f(z_max) - (z + hz), x + hx, f(z_max) - (z - hz)], axis=1)
    corners = corners - np.array([x_min, z_min, x_min, z_min], dtype=f)
    rng = np.array([x_max - x_min, z_max - z_min, x_max - x_min, z_max - z_min], dtype=f)
    out = corners / rng
    assert corners.dtype == f and out.dtype == f
    return corners, out


def reorder_projected_boxes(box_corners):
    """[x1, y1, x2, y2] -> [y1, x1, y2, x2], the order tf.image.crop_and_resize wants."""
    b = np.asarray(box_corners)
    return np.stack([b[:, 1], b[:, 0], b[:, 3], b[:, 2]], axis=1)
