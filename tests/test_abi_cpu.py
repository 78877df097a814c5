"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/csm_abi.h declares; struct layouts match; argument
validation returns status codes (never aborts).  No compute calls (no GPU here).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def csm():
    from cartographer_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "csm_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(csm):
    lib = csm.lib()
    names = _declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), "libcsm_b200.so does not export %s" % n


def test_struct_layouts(csm):
    assert C.sizeof(csm.CsmStats) == 64
    assert C.sizeof(csm.CsmJob2D) == 48
    assert C.sizeof(csm.CsmResult2D) == 48
    assert csm.JOB2D_DTYPE.itemsize == 48 and csm.RESULT2D_DTYPE.itemsize == 48
    # the refinement's records against the C compiler's view of include/csm_abi.h
    import subprocess
    import tempfile
    from cartographer_b200 import scan_matching as sm
    names = ["csm_ceres_options2d", "csm_ceres_job2d", "csm_ceres_result2d",
             "csm_ceres_options3d", "csm_ceres_job3d", "csm_ceres_result3d"]
    src = '#include <stdio.h>\n#include "include/csm_abi.h"\nint main(){printf("' + \
        " ".join(["%zu"] * len(names)) + '\\n",' + ",".join("sizeof(%s)" % n for n in names) + ");}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "sz.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", ROOT, os.path.join(d, "sz.c"), "-o", os.path.join(d, "sz")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "sz")]).split()]
    py = [sm.CsmCeresOptions2D, sm.CsmCeresJob2D, sm.CsmCeresResult2D, sm.CsmCeresOptions3D,
          sm.CsmCeresJob3D, sm.CsmCeresResult3D]
    assert sizes == [C.sizeof(t) for t in py]
    # ... and every field offset
    fields = [(n, t, f[0]) for n, t in zip(names, py) for f in t._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "include/csm_abi.h"\nint main(){' + \
        "".join('printf("%%zu\\n", offsetof(%s, %s));' % (n, f) for n, _, f in fields) + "}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "off.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", ROOT, os.path.join(d, "off.c"), "-o", os.path.join(d, "off")])
        offsets = [int(v) for v in subprocess.check_output([os.path.join(d, "off")]).split()]
    assert offsets == [getattr(t, f).offset for _, t, f in fields]
    # arrays of handles / pointers marshal element-wise
    job = sm.CsmCeresJob3D()
    cloud = np.zeros((4, 3), np.float32)
    job.grid[1] = C.c_void_p(0x1234)
    job.xyz[1] = cloud.ctypes.data_as(C.POINTER(C.c_float))
    raw = bytes(job)
    assert int.from_bytes(raw[8:16], "little") == 0x1234
    assert int.from_bytes(raw[24:32], "little") == cloud.ctypes.data


def test_invalid_arguments_return_status(csm):
    lib = csm.lib()
    out = C.c_void_p()
    cells = np.zeros((4, 4), np.uint16)
    # depth 0 is a CHECK failure in the reference (fast...2d.cc:174); here a status code
    st = lib.csm_stack2d_create(cells.ctypes.data_as(C.POINTER(C.c_uint16)), 4, 4,
                                C.c_double(0.05), C.c_double(0.1), C.c_double(0.1),
                                C.c_float(0.1), C.c_float(0.9), 0, 0, C.byref(out))
    assert st == 1
    assert b"branch_and_bound_depth" in lib.csm_last_error_string()
    st = lib.csm_stack2d_create(None, 4, 4, C.c_double(0.05), C.c_double(0.1), C.c_double(0.1),
                                C.c_float(0.1), C.c_float(0.9), 3, 0, C.byref(out))
    assert st == 1


def test_no_cpu_fallback_without_gpu(csm):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = csm.lib()
    out = C.c_void_p()
    cells = np.zeros((4, 4), np.uint16)
    st = lib.csm_stack2d_create(cells.ctypes.data_as(C.POINTER(C.c_uint16)), 4, 4,
                                C.c_double(0.05), C.c_double(0.1), C.c_double(0.1),
                                C.c_float(0.1), C.c_float(0.9), 3, 0, C.byref(out))
    assert st == 2, "expected CSM_E_CUDA"
    from cartographer_b200 import scan_matching as sm
    from benchmarks import synthetic
    grid = synthetic.GridSpec(cells, 0.05, 0.1, 0.1)
    with pytest.raises(csm.CsmError):
        sm.FastCorrelativeScanMatcher2D(grid, sm.FastCorrelativeScanMatcherOptions2D(1.0, 0.5, 3))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under cartographer_b200/ may touch it."""
    pkg = os.path.join(ROOT, "cartographer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
