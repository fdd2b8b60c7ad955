"""ctypes loader for the plain-C oracle (oracle/liboracle_flat.so) - test helper."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"


def load():
    so = ORACLE_DIR / "liboracle_flat.so"
    if not so.exists():
        subprocess.check_call(["make", "-C", str(ORACLE_DIR)])
    lib = ctypes.CDLL(str(so))
    lib.oracle_knn_flat.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                    ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int]
    lib.oracle_knn_flat.restype = None
    lib.oracle_normalize_l2.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long]
    lib.oracle_normalize_l2.restype = None
    return lib


def knn_flat(xq, xb, k, metric, nthreads=1):
    lib = load()
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    D = np.empty((xq.shape[0], k), dtype=np.float32)
    I = np.empty((xq.shape[0], k), dtype=np.int64)
    lib.oracle_knn_flat(xq.ctypes.data, xb.ctypes.data, xq.shape[0], xb.shape[0], xq.shape[1], k,
                        metric, D.ctypes.data, I.ctypes.data, nthreads)
    return D, I


def normalize_l2(x):
    lib = load()
    assert x.dtype == np.float32 and x.flags.c_contiguous
    lib.oracle_normalize_l2(x.ctypes.data, x.shape[0], x.shape[1])
