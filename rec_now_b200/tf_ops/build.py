"""Build recnow_tf_ops.so against the installed TensorFlow (python -m rec_now_b200.tf_ops.build).

Cannot run in the development image: TensorFlow is not installed there.  Needs librecnow_b200.so
(python -m rec_now_b200.build) and g++.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(HERE, "recnow_tf_ops.so")


def build() -> str:
    try:
        import tensorflow as tf
    except ImportError as e:                       # say so loudly; there is nothing to fall back to
        raise RuntimeError("TensorFlow is not installed: the TF op shim cannot be built here") from e
    lib = os.path.join(PKG, "librecnow_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("build librecnow_b200.so first: python -m rec_now_b200.build")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-shared", "-fPIC", "-O2", os.path.join(HERE, "recnow_tf_ops.cc"), "-o", OUT,
           "-I", os.path.join(os.path.dirname(PKG), "include"), "-I", os.path.join(cuda, "include"),
           *tf.sysconfig.get_compile_flags(), *tf.sysconfig.get_link_flags(),
           "-L", PKG, "-lrecnow_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
    sys.exit(0)
