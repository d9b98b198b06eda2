"""Builds the C oracle (oracle/c_oracle.c -> oracle/_build/liboracle.so) with gcc.
TEST INFRASTRUCTURE: building the checker is not using it.

Reference-built helpers (oracle/_ref/): the only compiled code of the reference on this path is the
TensorFlow custom op under avod/core/ops/correlation/, which needs TensorFlow 1.3 headers and a
TF runtime that are not in this image — treated as unbuildable. The reference's
other native code (wavedata integral_images_3d.cpp) belongs to the 3D filter, which is off the path.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
