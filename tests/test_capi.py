"""CPU: the C-ABI shared library loads and exports exactly what include/rii_b200.h declares; the product path
has no CPU fallback (fails loudly without a device) and never references the oracle."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rii_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rii_[a-z_A-Z0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rii_b200 import _capi, build
    build.build()
    lib = _capi.lib()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), "librii_b200.so does not export " + name
    assert sorted(_capi.SYMBOLS) == decl, "ctypes table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH]).decode()
    exported = set(re.findall(r" T (rii_\w+)", out))
    assert exported == set(decl)


def test_library_is_sm100a_only():
    from rii_b200 import _capi
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", _capi.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    import numpy as np
    from rii_b200 import main, _capi
    with pytest.raises(_capi.RiiError):
        main.RiiCpp(np.zeros((4, 16, 2), np.float32), False)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rii_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_strict_dtypes_and_errors_without_device():
    from rii_b200 import main
    import numpy as np
    with pytest.raises(TypeError):
        main._strict(np.zeros(3, np.float64), np.float32, 1, "query")
    with pytest.raises(TypeError):
        main._strict(np.zeros((3, 2), np.float32)[:, 0], np.float32, 1, "query")  # non-contiguous
