"""Builds the checkers. TEST INFRASTRUCTURE: building the checker is not using it.

1. oracle/c_oracle.c -> oracle/_build/liboracle.so (gcc): the C restatement of S3/S4/S5.
2. oracle/_ref/libcorr_ref.so (nvcc, sm_100a): the reference's OWN correlation kernels,
   /root/reference/avod/core/ops/correlation/{correlation_kernel,pad,correlation_grad_kernel}.cu.cc,
   compiled UNMODIFIED from where they lie against a stub header tree (oracle/ref_stubs/: the
   Eigen::GpuDevice / CUDA_1D_KERNEL_LOOP / GetCudaLaunchConfig they use) and linked with
   oracle/ref_corr_driver.cu, which restates the host-side shape/padding logic of the two TF
   OpKernels (correlation_kernel.cc:26-124, correlation_grad_kernel.cc:28-151). Needs the read-only
   checkout; on the GPU box the prebuilt .so (git-ignored, NOT gpurun-ignored) is used as is.
3. oracle/_ref/py/: the reference's own S1/S2 Python modules staged as they are (a git-ignored copy
   that travels to the GPU box, where /root/reference does not exist) so that bench.py's
   cpu_baseline / --impl reference legs and the live-reference oracle tests can run the reference's
   NumPy there. Nothing under oracle/_ref/ is committed.

The reference's other native code (wavedata integral_images_3d.cpp) belongs to the 3-D filter,
which is off the path.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")

REFERENCE_ROOT = os.environ.get("DODT_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
CORR_REF_LIB = os.path.join(REF_DIR, "libcorr_ref.so")
REF_PY_DIR = os.path.join(REF_DIR, "py")
CORR_SRC_DIR = os.path.join(REFERENCE_ROOT, "avod", "core", "ops", "correlation")
CORR_SOURCES = ["correlation_kernel.cu.cc", "pad.cu.cc", "correlation_grad_kernel.cu.cc"]

# the S1/S2 closure of the reference's Python plus the host anchor helpers between the stages and
# three_d_iou, relative to the checkout
REF_PY_FILES = [
    "avod/__init__.py",
    "avod/core/__init__.py",
    "avod/core/anchor_filter.py",
    "avod/core/anchor_encoder.py",
    "avod/core/anchor_generator.py",
    "avod/core/anchor_generators/__init__.py",
    "avod/core/anchor_generators/grid_anchor_3d_generator.py",
    "avod/core/anchor_projector.py",
    "avod/core/box_3d_encoder.py",
    "avod/core/format_checker.py",
    "avod/core/bev_generators/__init__.py",
    "avod/core/bev_generators/bev_generator.py",
    "avod/core/bev_generators/bev_slices.py",
    "wavedata/wavedata/__init__.py",
    "wavedata/wavedata/tools/__init__.py",
    "wavedata/wavedata/tools/core/__init__.py",
    "wavedata/wavedata/tools/core/calib_utils.py",
    "wavedata/wavedata/tools/core/geometry_utils.py",
    "wavedata/wavedata/tools/core/integral_image.py",
    "wavedata/wavedata/tools/core/integral_image_2d.py",
    "wavedata/wavedata/tools/core/voxel_grid_2d.py",
    "wavedata/wavedata/tools/obj_detection/__init__.py",
    "wavedata/wavedata/tools/obj_detection/obj_utils.py",
    "wavedata/wavedata/tools/obj_detection/evaluation.py",
]


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("%s failed:\n%s\n%s" % (cmd[0], res.stdout, res.stderr))
    return res


def build_c_oracle(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    _run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def reference_present():
    return all(os.path.exists(os.path.join(CORR_SRC_DIR, f)) for f in CORR_SOURCES)


def build_corr_ref(force=False):
    """The reference's correlation kernels as a shared library; None when the checkout is absent
    and nothing was prebuilt."""
    if not reference_present():
        return CORR_REF_LIB if os.path.exists(CORR_REF_LIB) else None
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return CORR_REF_LIB if os.path.exists(CORR_REF_LIB) else None
    os.makedirs(REF_DIR, exist_ok=True)
    driver = os.path.join(HERE, "ref_corr_driver.cu")
    srcs = [os.path.join(CORR_SRC_DIR, f) for f in CORR_SOURCES]
    deps = srcs + [driver, os.path.join(HERE, "ref_stubs", "tensorflow", "core", "util", "cuda_kernel_helper.h"),
                   os.path.join(HERE, "ref_stubs", "third_party", "eigen3", "unsupported", "Eigen", "CXX11", "Tensor")]
    if not force and os.path.exists(CORR_REF_LIB) and \
            all(os.path.getmtime(CORR_REF_LIB) >= os.path.getmtime(d) for d in deps):
        return CORR_REF_LIB
    # the reference sources are .cc files holding CUDA code: -x cu; GOOGLE_CUDA=1 opens their #if
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-x", "cu", "-DGOOGLE_CUDA=1", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-Wno-deprecated-gpu-targets",
           "-I", os.path.join(HERE, "ref_stubs"), "-I", CORR_SRC_DIR,
           "-o", CORR_REF_LIB] + srcs + [driver]
    _run(cmd)
    return CORR_REF_LIB


def stage_reference_python(force=False):
    """Copies the reference's S1/S2 modules into oracle/_ref/py (git-ignored). Returns the directory,
    or None when neither the checkout nor an earlier staging exists."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "avod", "core")) or \
            os.path.realpath(REFERENCE_ROOT) == os.path.realpath(REF_PY_DIR):
        return REF_PY_DIR if os.path.isdir(os.path.join(REF_PY_DIR, "avod")) else None
    for rel in REF_PY_FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(REF_PY_DIR, rel)
        if not os.path.exists(src):
            raise RuntimeError("reference file missing: " + src)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy2(src, dst)
    return REF_PY_DIR


def build(force=False):
    lib = build_c_oracle(force)
    build_corr_ref(force)
    stage_reference_python(force)
    return lib


if __name__ == "__main__":
    print(build(force=True))
    print(CORR_REF_LIB, os.path.exists(CORR_REF_LIB))
    print(REF_PY_DIR, os.path.isdir(REF_PY_DIR))
