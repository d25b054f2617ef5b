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


def _header_struct_fields(name):
    """(type, field, count) of every member of `struct name` in include/sphgpu.h, in order"""
    src = open(os.path.join(ROOT, "include", "sphgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?([A-Za-z_0-9]+)\s+(.*)$", decl, flags=re.S)
        ctype, rest = m.group(2), m.group(3)
        for item in rest.split(","):
            item = item.strip()
            ptr = item.startswith("*")
            item = item.lstrip("* ")
            mm = re.match(r"([A-Za-z_0-9]+)(?:\[([A-Za-z_0-9]+)\])?$", item)
            cnt = mm.group(2)
            cnt = 1 if cnt is None else (8 if cnt == "SPHGPU_MAXTYPES" else int(cnt))
            out.append((ctype + ("*" if ptr else ""), mm.group(1), cnt))
    return out


def test_field_offsets_match_a_c_probe(tmp_path):
    """ctypes mirror (phantom_b200/params.py, api.py) against offsetof() printed by a C program compiled from include/sphgpu.h"""
    import ctypes as C
    import subprocess
    from phantom_b200 import params as P, api
    structs = {"sphgpu_params": P.SphParams, "sphgpu_scalars": P.SphScalars, "sphgpu_step_out": P.SphStepOut,
               "sphgpu_energies": P.SphEnergies, "sphgpu_host_arrays": api.HostArrays}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sphgpu.h"', 'int main(void) {']
    for sname in structs:
        for _, f, _ in _header_struct_fields(sname):
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (sname, f, sname, f))
        lines.append('  printf("%s __size__ %%zu\\n", sizeof(%s));' % (sname, sname))
    lines.append("  return 0; }")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        s_, f, off = ln.split()
        got[(s_, f)] = int(off)
    nchecked = 0
    for sname, cls in structs.items():
        assert C.sizeof(cls) == got[(sname, "__size__")], sname
        names = [f[0] for f in cls._fields_]
        for _, f, _ in _header_struct_fields(sname):
            assert f in names, (sname, f)
            assert getattr(cls, f).offset == got[(sname, f)], (sname, f)
            nchecked += 1
        assert len(names) == len(_header_struct_fields(sname)), sname
    assert nchecked > 100


def test_fortran_shim_in_integration_md_matches_header():
    """the bind(C) derived types of INTEGRATION.md list the fields of include/sphgpu.h in the same order with the same kinds and counts"""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    kinds = {"int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)", "double": "real(c_double)"}
    for sname in ("sphgpu_params", "sphgpu_scalars"):
        body = re.search(r"type, bind\(C\) :: %s\b.*?\n(.*?)\n\s*end type" % sname, md, flags=re.S).group(1)
        shim = []
        for ln in body.splitlines():
            kind, rest = ln.split("::")
            for item in re.findall(r"([A-Za-z_0-9]+)(?:\(([0-9:]+)\))?", rest):
                cnt = 1
                if item[1]:
                    lo_hi = item[1].split(":")
                    cnt = int(lo_hi[0]) if len(lo_hi) == 1 else int(lo_hi[1]) - int(lo_hi[0]) + 1
                shim.append((kind.strip(), item[0].lower(), cnt))
        hdr = [(kinds[t], f.lower(), c) for t, f, c in _header_struct_fields(sname)]
        assert shim == hdr, (sname, [a for a, b in zip(shim, hdr) if a != b][:3])
    # every hidden input of the header is filled by gpu_fill_params
    fill = md[md.index("subroutine gpu_fill_params"):md.index("end module gpuderivs")]
    for _, f, _ in _header_struct_fields("sphgpu_params"):
        assert re.search(r"p%%%s\b" % f, fill), f


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
