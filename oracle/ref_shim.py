"""Import shim for the reference's own Python (S1/S2) — TEST INFRASTRUCTURE ONLY.

Puts the read-only checkout (default /root/reference, override with DODT_REFERENCE_ROOT) and its
vendored wavedata on sys.path and injects a stub `tensorflow` module so that
avod.core.{anchor_filter,anchor_projector,box_3d_encoder}, avod.core.bev_generators.bev_slices and
avod.core.anchor_generators.grid_anchor_3d_generator import under NumPy 2 without TensorFlow.
The checkout does not exist on the GPU box; the S1/S2 closure staged by oracle/build_oracle.py
(oracle/_ref/py) is used there. Everything that uses this module must be skipped when
`available()` is False (the committed fixtures under tests/golden/ take over), and whatever needs
more than S1/S2 when `full_checkout()` is False.
"""
import importlib.machinery
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "py")


def _root():
    """The read-only checkout where it exists; otherwise the copy of the S1/S2 modules that
    oracle/build_oracle.py staged into git-ignored oracle/_ref/py (it travels to the GPU box)."""
    env = os.environ.get("DODT_REFERENCE_ROOT")
    for cand in (env, "/root/reference", _STAGED):
        if cand and os.path.isdir(os.path.join(cand, "avod", "core")) and \
                os.path.isdir(os.path.join(cand, "wavedata", "wavedata")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _root()


def full_checkout():
    """True when the whole reference tree (anchor generators, projector, fixtures ...) is there,
    not only the staged S1/S2 modules."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "avod", "core", "anchor_generators"))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "avod", "core")) and \
        os.path.isdir(os.path.join(REFERENCE_ROOT, "wavedata", "wavedata"))


class _AnyModule(types.ModuleType):
    """A module whose every attribute is another permissive module (tf.contrib.slim ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _AnyModule(self.__name__ + "." + name)
        child.__spec__ = importlib.machinery.ModuleSpec(child.__name__, None)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return None


def install():
    """Idempotent. Returns True if the reference can be imported afterwards."""
    if not available():
        return False
    for p in (os.path.join(REFERENCE_ROOT, "wavedata"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "tensorflow" not in sys.modules:
        tf = _AnyModule("tensorflow")
        tf.__spec__ = importlib.machinery.ModuleSpec("tensorflow", None)
        tf.Tensor = type("Tensor", (), {})   # anchor_projector.py:28 does isinstance(x, tf.Tensor)
        sys.modules["tensorflow"] = tf
    return True


class KittiUtilsStandIn:
    """The one method BevSlices needs from KittiUtils (avod/datasets/kitti/kitti_utils.py:81-109);
    the real class needs protobuf-generated modules that are not in the checkout. It calls the
    reference's own obj_utils.get_point_filter, so the arithmetic is the reference's."""

    def create_slice_filter(self, point_cloud, area_extents, ground_plane, ground_offset_dist,
                            offset_dist):
        import numpy as np
        from wavedata.tools.obj_detection import obj_utils
        offset_filter = obj_utils.get_point_filter(point_cloud, area_extents, ground_plane,
                                                   offset_dist)
        road_filter = obj_utils.get_point_filter(point_cloud, area_extents, ground_plane,
                                                 ground_offset_dist)
        return np.logical_xor(offset_filter, road_filter)


class SlicesConfig:
    """Stand-in for the bev_generator.slices protobuf message (fp32 fields,
    avod/protos/kitti_utils.proto:31-32)."""

    def __init__(self, height_lo, height_hi, num_slices):
        self.height_lo = height_lo
        self.height_hi = height_hi
        self.num_slices = num_slices


def reference_bev_slices(height_lo, height_hi, num_slices):
    """The reference's BevSlices generator (avod/core/bev_generators/bev_slices.py:8-31)."""
    assert install()
    from avod.core.bev_generators.bev_slices import BevSlices
    return BevSlices(SlicesConfig(height_lo, height_hi, num_slices), KittiUtilsStandIn())


def reference_sliced_voxel_grid_2d(point_cloud, ground_plane, area_extents, voxel_size,
                                   height_lo=0.2, height_hi=2.0):
    """kitti_utils.py:212-277 (create_sliced_voxel_grid_2d_v2) with the reference's VoxelGrid2D."""
    assert install()
    import numpy as np
    from wavedata.tools.core.voxel_grid_2d import VoxelGrid2D
    mask = KittiUtilsStandIn().create_slice_filter(point_cloud, area_extents, ground_plane,
                                                   height_lo, height_hi)
    pts = np.asarray(point_cloud).T[mask]
    vg = VoxelGrid2D()
    vg.voxelize_2d(pts, voxel_size, extents=area_extents, ground_plane=ground_plane,
                   create_leaf_layout=True)
    return vg
