"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/sphgpu.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sphgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sphgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from phantom_b200 import api
    L = api.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(api.EXPORTS) == set(names)


def test_struct_layouts_match_header():
    import ctypes as C
    from phantom_b200.params import SphParams, SphScalars
    # sizes computed from include/sphgpu.h: 20 int32 + (6+2+8+5+4+5+1+3+8) doubles ; 7 doubles + 12 int64
    assert C.sizeof(SphParams) == 20 * 4 + 42 * 8
    assert C.sizeof(SphScalars) == 7 * 8 + 12 * 8


def test_no_cpu_fallback_without_device():
    import ctypes as C
    from phantom_b200 import api, default_params
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.SphGpuError):
        api.SphGpu(default_params())


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "phantom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("oracle/ ", ""), f
