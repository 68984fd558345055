"""Build librii_b200.so in-tree (sm_100a only).  `python -m rii_b200.build`"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rii_b200.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "kernels.cuh"), os.path.join(HERE, "csrc", "topk.cuh"), os.path.join(HERE, "csrc", "scan_dual.cuh"), os.path.join(HERE, "csrc", "scan_stream.cuh"),
        os.path.join(HERE, "..", "include", "rii_b200.h")]
OUT = os.path.join(HERE, "librii_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17"]


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
