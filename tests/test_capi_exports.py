"""The C-ABI library loads and exports every symbol include/xxmotif_gpu.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "xxmotif_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xxm_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    import xxmotif_b200 as xb
    lib = ctypes.CDLL(xb.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    from xxmotif_b200 import hostprep, localization, refine, sharded
    assert sorted(xb.EXPORTS + hostprep.EXTRA_EXPORTS + sharded.EXTRA_EXPORTS + localization.EXTRA_EXPORTS + refine.EXTRA_EXPORTS) == names


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xxmotif_b200 as xb
    with pytest.raises(xb.XxmError) as e:
        xb.Context(0)
    assert "no CUDA device" in str(e.value)


def test_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/xxmotif_gpu.h compiles as C99 and as C++17 on its own (no torch, no CUDA types)"""
    import subprocess
    hdr = os.path.join(ROOT, "include", "xxmotif_gpu.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)  # comments may mention how a caller distributes the NCCL id
    assert "torch" not in code and "cuda" not in code.lower() and "nccl" not in code.lower()
