"""Build librii_b200.so in-tree (sm_100a only).  `python -m rii_b200.build [--force] [-v]`

Every translation unit under csrc/ is compiled on its own (in parallel) and linked into one shared library."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "librii_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(HERE, "..", "include", "rii_b200.h")])


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(target) < os.path.getmtime(d) for d in deps)


def up_to_date():
    return not _stale(OUT, sources() + headers())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    hdr = headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdr)]

    def cc(src):
        cmd = [NVCC, "-c"] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", _obj(src), src]
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(cc, todo))
    subprocess.check_call([NVCC, "-shared"] + FLAGS + ["-o", OUT] + [_obj(s) for s in sources()])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
