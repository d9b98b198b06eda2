"""The C-ABI boundary on a host without a GPU: libdodt_fe.so builds (nvcc cross-compiles sm_100a),
loads, exports every symbol include/dodt_fe.h declares, and its host-only entry points (shape and
workspace arithmetic, error strings) work. No compute entry point is executed here; that the
product path FAILS LOUDLY without a device or without the library is checked instead.
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dodt_fe.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dodt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdodt_fe.so does not export %s" % n
    from dodt_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_header_is_plain_c(tmp_path):
    """The boundary is a C header: it must compile as C99 with no CUDA / torch types."""
    src = tmp_path / "t.c"
    src.write_text('#include "dodt_fe.h"\nint main(void){ dodt_bev_params p; (void)p; return DODT_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layouts_match_header(lib, tmp_path):
    """sizeof/offsetof of the ABI structs as gcc sees the header == the ctypes mirrors."""
    from dodt_b200 import _lib
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dodt_fe.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dodt_bev_params), '
                   'offsetof(dodt_bev_params, occ_lo), offsetof(dodt_bev_params, density_lut), '
                   'sizeof(dodt_gather_spec), offsetof(dodt_gather_spec, width), '
                   'sizeof(dodt_crop_spec), offsetof(dodt_crop_spec, channels), sizeof(dodt_anchor_grid), '
                   'offsetof(dodt_anchor_grid, sizes), offsetof(dodt_anchor_grid, n_sizes)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.BevParams), _lib.BevParams.occ_lo.offset, _lib.BevParams.density_lut.offset,
            ctypes.sizeof(_lib.GatherSpec), _lib.GatherSpec.width.offset,
            ctypes.sizeof(_lib.CropSpec), _lib.CropSpec.channels.offset,
            ctypes.sizeof(_lib.AnchorGrid), _lib.AnchorGrid.sizes.offset, _lib.AnchorGrid.n_sizes.offset]
    assert got == want


def test_host_only_entry_points(lib):
    from dodt_b200 import _lib, ops
    assert lib.dodt_version() == 201
    assert lib.dodt_strerror(0).decode().lower().startswith("ok") or lib.dodt_strerror(0)
    for code in (_lib.DODT_EINVAL, _lib.DODT_ESHAPE, _lib.DODT_ECAPACITY, _lib.DODT_ECUDA, _lib.DODT_EALIGN):
        assert len(lib.dodt_strerror(code)) > 0
    # wavedata voxel_grid_2d.py:125-130: floor(min/v), ceil(max/v - 1)
    assert ops.bev_grid([[-40, 40], [-5, 3], [0, 70]], float(np.float32(0.1))) == (800, 80, 700, -400, -50, 0)
    assert ops.bev_grid([[-40, 40], [-5, 3], [0, 70]], float(np.float32(0.05)))[:3:2] == (1600, 1400)
    assert ops.bev_grid([[-50, 50], [-5, 5], [0, 70]], 0.1)[:3:2] == (1000, 700)
    # avod/core/ops/correlation/correlation_kernel.cc:39-57
    assert ops.correlation_out_shape(700, 800, 1, 5, 1, 2, 5) == (700, 800, 25)
    assert ops.correlation_out_shape(64, 96, 1, 20, 1, 2, 20) == (64, 96, 441)
    assert ops.correlation_out_shape(48, 64, 3, 4, 2, 2, 4) == (23, 31, 25)
    with pytest.raises(ValueError):
        ops.correlation_out_shape(4, 64, 1, 8, 1, 2, 0)
    hwc = (ctypes.c_int32 * 3)()
    assert lib.dodt_correlation_out_shape(64, 64, 2, 4, 1, 2, 4, hwc) == _lib.DODT_EINVAL
    # workspaces grow with their sizes and are non-zero
    assert 0 < ops.bev_workspace_bytes(1000, 5, 800, 700) <= ops.bev_workspace_bytes(500000, 5, 1600, 1400)
    assert 0 < ops.integral_workspace_bytes(800, 700) <= ops.integral_workspace_bytes(1600, 1400)
    assert 0 < ops.nms_workspace_bytes(1024) < ops.nms_workspace_bytes(89600)
    assert lib.dodt_nms_state_offset(89600) < ops.nms_workspace_bytes(89600)
    assert lib.dodt_compact_workspace_bytes(89600) > 0
    assert lib.dodt_launch_count() >= 0
    # S4 backward: no scratch since the displacement flip is gathered per tile inside the kernel
    assert ops.correlation_grad_workspace_bytes(1, 700, 800, 32, 1, 5, 1, 2, 5) == 0
    assert ops.correlation_grad_workspace_bytes(1, 64, 96, 8, 3, 4, 2, 2, 4) == 0


def test_new_entry_points_validate_arguments_on_the_host(lib):
    """Argument errors are reported before any device work (so they show on a host without a GPU)."""
    from dodt_b200 import _lib
    buf = (ctypes.c_float * 8)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    one = (ctypes.c_void_p * 1)(ptr)
    two = (ctypes.c_void_p * 2)(ptr, ptr)
    # fewer than two maps, a NULL table, a negative CTA cap, an even kernel size
    assert lib.dodt_correlation_stream(one, 1, one, 8, 8, 8, 1, 2, 1, 2, 2, 0, None) == _lib.DODT_EINVAL
    assert lib.dodt_correlation_stream(None, 2, one, 8, 8, 8, 1, 2, 1, 2, 2, 0, None) == _lib.DODT_EINVAL
    assert lib.dodt_correlation_stream(two, 2, one, 8, 8, 8, 1, 2, 1, 2, 2, -1, None) == _lib.DODT_EINVAL
    assert lib.dodt_correlation_stream(two, 2, one, 8, 8, 8, 2, 2, 1, 2, 2, 0, None) == _lib.DODT_EINVAL
    assert lib.dodt_correlation_stream(two, 2, one, 4, 64, 8, 1, 8, 1, 2, 0, 0, None) == _lib.DODT_ESHAPE
    # gradients: both outputs NULL, even kernel size, neighbourhood that does not fit
    assert lib.dodt_correlation_grad(ptr, ptr, ptr, 1, 8, 8, 8, 1, 2, 1, 2, 2, None, None, None, 0, None) \
        == _lib.DODT_EINVAL
    assert lib.dodt_correlation_grad(ptr, ptr, ptr, 1, 8, 8, 8, 2, 2, 1, 2, 2, ptr, ptr, None, 0, None) \
        == _lib.DODT_EINVAL
    assert lib.dodt_correlation_grad(ptr, ptr, ptr, 1, 4, 64, 8, 1, 8, 1, 2, 0, ptr, ptr, None, 0, None) \
        == _lib.DODT_ESHAPE


def test_density_lut_matches_reference_formula():
    """min(1, ln(n+1)/ln 16) (avod/core/bev_generators/bev_generator.py:34-35): LUT saturates at 15."""
    from dodt_b200 import ops
    lut = ops.density_lut(np.log(16))
    assert len(lut) == 15 and lut[0] == 0.0
    np.testing.assert_array_equal(lut, np.minimum(1.0, np.log(np.arange(15) + 1) / np.log(16)))
    p = ops.make_bev_params([0, -1, 0, 1.65], [[-40, 40], [-5, 3], [0, 70]], 0.1, -0.2, 2.3, 5)
    assert p.num_slices == 5 and p.filter_mode == 1 and p.density_lut_len == 15
    with pytest.raises(ValueError):
        ops.make_bev_params([0, -1, 0, 1.65], [[-40, 40], [-5, 3]], 0.1, -0.2, 2.3, 5)
    with pytest.raises(ValueError):
        ops.make_bev_params([0, -1, 0, 1.65], [[-40, 40], [-5, 3], [0, 70]], 0.1, -0.2, 2.3, 16)


def test_product_path_has_no_cpu_fallback(lib):
    """CPU tensors are refused outright; the package never imports the oracle."""
    from dodt_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.correlation(torch.zeros(1, 8, 8, 4), torch.zeros(1, 8, 8, 4), 1, 2, 1, 1, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.nms(torch.zeros(4, 4), torch.zeros(4), 2, 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.integral_image_2d(torch.zeros(4, 4, dtype=torch.uint8))
    pkg = os.path.join(ROOT, "dodt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks behaviour on a host without a GPU")
def test_compute_entry_point_reports_ecuda_without_device(lib):
    """One call with NULL-free host-side-valid arguments: without a device the library must answer
    DODT_ECUDA (never compute on the CPU). Nothing is dereferenced before the device check."""
    from dodt_b200 import _lib
    hwc = (ctypes.c_int32 * 3)()
    lib.dodt_correlation_out_shape(8, 8, 1, 2, 1, 1, 2, hwc)
    buf = (ctypes.c_float * 8)()
    rc = lib.dodt_correlation(ctypes.cast(buf, ctypes.c_void_p), ctypes.cast(buf, ctypes.c_void_p),
                              1, 1, 1, 1, 1, 0, 1, 1, 0, ctypes.cast(buf, ctypes.c_void_p), None)
    assert rc == _lib.DODT_ECUDA, rc
    assert len(lib.dodt_last_cuda_error()) > 0


def test_missing_library_fails_loudly(tmp_path):
    """A fresh interpreter whose LIB_PATH does not exist raises on load (no silent fallback)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dodt_b200 import _lib\n"
            "_lib.LIB_PATH = %r\n"
            "try:\n    _lib.load()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n"
            % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert "RAISED True" in out


def test_grid_anchor_shape_matches_reference_arange(lib):
    """Host-only: (nz, nx, sizes, rotations) == the np.arange lengths of
    avod/core/anchor_generators/grid_anchor_3d_generator.py:62-72, including the degenerate areas
    of grid_anchor_3d_generator_test.py:32-70."""
    from oracle import anchor_helpers as A
    from dodt_b200 import ops
    from oracle import synth_ref as synth
    assert ops.grid_anchor_shape(synth.AREA_EXTENTS, synth.ANCHOR_STRIDE, 2) == (140, 160, 2, 2)
    for ext, stride in (([(-1., 1.), (-1., 0.), (0., 1.)], [1, 1]), ([(0., 0.), (-1., 0.), (0., 2.)], [1, 1]),
                        ([(-1., 1.), (-1., 0.), (0., 0.)], [1, 1]), ([(-3.3, 7.1), (-1, 0), (0.2, 9.9)], [0.7, 0.3])):
        boxes = A.tile_anchors_3d(ext, [[1., 1., 1.], [2., 1., 1.]], stride, [0., -1., 0., 0.])
        nz, nx, ns, nr = ops.grid_anchor_shape(ext, stride, 2)
        assert nz * nx * ns * nr == len(boxes), (ext, stride)
