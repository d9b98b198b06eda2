"""Build recipe of libdodt_fe.so (hand-written sm_100a kernels + the C ABI in include/dodt_fe.h).

nvcc cross-compiles for sm_100a without a GPU, so this runs on a CPU-only host; the built library
stays in-tree (dodt_b200/libdodt_fe.so, git-ignored) so that it travels to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libdodt_fe.so")

# the diagnostic variant (tools/ only): -DDODT_DIAG turns the DODT_KNOB measurement knobs into
# environment lookups and compiles the A/B / diagnostic kernel instantiations in
DIAG_OBJ_DIR = os.path.join(HERE, "_obj_diag")
DIAG_LIB_PATH = os.path.join(HERE, "libdodt_fe_diag.so")

SOURCES = ["common.cu", "bev_slices.cu", "anchor_filter.cu", "anchor_fused.cu", "crop_resize.cu", "correlation.cu", "correlation_tma.cu",
           "correlation_feed.cu", "correlation_grad.cu", "nms.cu", "frontend.cu", "anchors.cu", "lidar.cu", "iou3d.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",   # host-side constants are rounded op by op like NumPy
    "-Xptxas", "-v",
    "-Wno-deprecated-gpu-targets",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built on this host")
    return exe


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "dodt_fe.h"))
    return hs


def _compile_one(src, verbose, diag=False):
    obj = os.path.join(DIAG_OBJ_DIR if diag else OBJ_DIR, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    digest = _digest([os.path.join(CSRC, src)] + _headers()) + ("-diag" if diag else "")
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DDODT_DIAG"] if diag else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return obj, True, res.stderr


def build(force=False, verbose=False, diag=False):
    """Compile every kernel source for sm_100a and link dodt_b200/libdodt_fe.so (or, diag=True, the
    diagnostic variant libdodt_fe_diag.so that only tools/ load). Returns the library's path."""
    obj_dir, lib_path = (DIAG_OBJ_DIR, DIAG_LIB_PATH) if diag else (OBJ_DIR, LIB_PATH)
    os.makedirs(obj_dir, exist_ok=True)
    if force:
        for f in os.listdir(obj_dir):
            os.remove(os.path.join(obj_dir, f))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, verbose, diag), sources))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if rebuilt or not os.path.exists(lib_path):
        cmd = [_nvcc(), "-shared", "-o", lib_path] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return lib_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv)
    print(path)
