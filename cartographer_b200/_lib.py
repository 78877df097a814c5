"""ctypes binding of libcsm_b200.so (the C ABI declared in include/csm_abi.h).

There is no CPU fallback: if the shared library is missing this module raises at
import time of the first call, and every entry point returns CSM_E_CUDA without a
usable device.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# CSM_B200_LIB points at an alternative build of the same library (A/B timing of kernel
# variants); the default is the in-tree product library.
SO_PATH = os.environ.get("CSM_B200_LIB") or os.path.join(_PKG, "libcsm_b200.so")
CSRC = os.path.join(_PKG, "csrc")


class CsmStats(C.Structure):
    _fields_ = [("candidates_scored", C.c_int64), ("lowest_resolution_candidates", C.c_int64),
                ("nodes_expanded", C.c_int64), ("leaves_tied", C.c_int64),
                ("num_scans", C.c_int32), ("best_scan_index", C.c_int32),
                ("best_x_offset", C.c_int32), ("best_y_offset", C.c_int32),
                ("host_tie_resolves", C.c_int32), ("host_syncs", C.c_int32),
                ("device_ms", C.c_float), ("collective_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class CsmJob2D(C.Structure):
    _fields_ = [("stack_index", C.c_int32), ("cloud_index", C.c_int32),
                ("full_submap", C.c_int32), ("reserved", C.c_int32),
                ("initial_pose", C.c_double * 3), ("min_score", C.c_float),
                ("reserved_f", C.c_float)]


class CsmResult2D(C.Structure):
    _fields_ = [("found", C.c_int32), ("score", C.c_float), ("pose_estimate", C.c_double * 3),
                ("best_scan_index", C.c_int32), ("best_x_offset", C.c_int32),
                ("best_y_offset", C.c_int32), ("leaves_tied", C.c_int32)]


JOB2D_DTYPE = np.dtype([("stack_index", "<i4"), ("cloud_index", "<i4"), ("full_submap", "<i4"),
                        ("reserved", "<i4"), ("initial_pose", "<f8", (3,)),
                        ("min_score", "<f4"), ("reserved_f", "<f4")], align=True)
RESULT2D_DTYPE = np.dtype([("found", "<i4"), ("score", "<f4"), ("pose_estimate", "<f8", (3,)),
                           ("best_scan_index", "<i4"), ("best_x_offset", "<i4"),
                           ("best_y_offset", "<i4"), ("leaves_tied", "<i4")], align=True)
assert JOB2D_DTYPE.itemsize == C.sizeof(CsmJob2D)
assert RESULT2D_DTYPE.itemsize == C.sizeof(CsmResult2D)


class CsmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libcsm_b200 status %d: %s" % (status, message))
        self.status = status


def build(verbose=False):
    """Compile libcsm_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libcsm_b200.so failed")
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "%s is missing — build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % SO_PATH)
        _lib = C.CDLL(SO_PATH)
        _lib.csm_last_error_string.restype = C.c_char_p
        _lib.csm_kernel_launch_count.restype = C.c_int64
    return _lib


def check(status):
    if status != 0:
        raise CsmError(status, lib().csm_last_error_string().decode("utf-8", "replace"))


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))
