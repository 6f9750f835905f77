"""
TEST INFRASTRUCTURE ONLY -- builds and wraps oracle/walk_oracle.c (plain C restatement of the
reference walk rule, see that file's header for the reference file:line it follows).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')
LIB_PATH = os.path.join(BUILD_DIR, 'libwalk_oracle.so')
SRC_PATH = os.path.join(HERE, 'walk_oracle.c')
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-o', LIB_PATH, SRC_PATH, '-lm'])
    return LIB_PATH


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_walks.restype = ctypes.c_int
        _lib.oracle_walks.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def c_walks(rowptr, col, w, w_is_int, starts, length, p, q, node2vec, rule, uniforms, col_sorted=None):
    lib = _load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    if w is not None and len(w) == 0:
        w = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64)
    if col_sorted is not None:
        col_sorted = np.ascontiguousarray(col_sorted, dtype=np.int32)
    out = np.empty((len(starts), length), dtype=np.int32)
    assert uniforms.size == len(starts) * max(length - 1, 0)
    rc = lib.oracle_walks(
        rowptr.ctypes.data, col.ctypes.data, w.ctypes.data if w is not None else None, int(bool(w_is_int)),
        col_sorted.ctypes.data if col_sorted is not None else None, len(rowptr) - 1,
        starts.ctypes.data, len(starts), int(length), float(p), float(q), int(bool(node2vec)), int(rule),
        uniforms.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f'oracle_walks failed with status {rc}')
    return out
