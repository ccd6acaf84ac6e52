"""ctypes wrapper of the oracle's C restatement (oracle/c/bann_cpu.c). TEST INFRASTRUCTURE:
used by tests (pinned against the NumPy oracle) and by bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libbann_cpu.so")
_fp = C.POINTER(C.c_float)
_ACT = {"tanh": 0, "relu": 1, "leaky_relu": 2, "silu": 3, "identity": 4}


def load():
    if not os.path.exists(_PATH):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(_PATH)
    lib.bann_cpu_num_threads.restype = C.c_int
    lib.bann_cpu_set_threads.restype = C.c_int
    lib.bann_cpu_set_threads.argtypes = [C.c_int]
    lib.bann_cpu_decode_std.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32, _fp, _fp, _fp]
    lib.bann_cpu_rss.restype = C.c_float
    lib.bann_cpu_rss.argtypes = [_fp, _fp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_int, _fp, _fp]
    lib.bann_cpu_backprop.restype = C.c_float
    lib.bann_cpu_backprop.argtypes = [_fp, _fp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_int, _fp, _fp]
    lib.bann_cpu_leapfrog.restype = C.c_float
    lib.bann_cpu_leapfrog.argtypes = [_fp, _fp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.c_int,
                                      C.c_int, _fp, _fp, _fp, _fp, C.c_float, C.c_uint32, _fp]
    return lib


def _p(a):
    return a.ctypes.data_as(_fp)


class CPort:
    def __init__(self, threads=None):
        """threads: OpenMP team size to pin (None: what the environment gives, e.g. OMP_NUM_THREADS)."""
        self.lib = load()
        self.threads = self.lib.bann_cpu_set_threads(int(threads)) if threads else self.lib.bann_cpu_num_threads()

    def decode_std(self, payload, n, cols, means, stds):
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        X = np.empty(n * len(cols), dtype=np.float32)
        self.lib.bann_cpu_decode_std(payload.ctypes.data_as(C.c_void_p), n, cols.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     len(cols), _p(np.ascontiguousarray(means, dtype=np.float32)),
                                     _p(np.ascontiguousarray(stds, dtype=np.float32)), _p(X))
        return X  # column-major [n, m]

    def backprop(self, X, y, n, m, widths, act, theta):
        w = np.ascontiguousarray(widths, dtype=np.uint32)
        d = np.empty(theta.size, dtype=np.float32)
        rss = self.lib.bann_cpu_backprop(_p(X), _p(y), n, m, w.ctypes.data_as(C.POINTER(C.c_uint32)), len(widths),
                                         _ACT[act], _p(theta), _p(d))
        return float(rss), d

    def leapfrog(self, X, y, n, m, widths, act, lasso, stdn, theta, mom, eps, lam, lam_e, L):
        w = np.ascontiguousarray(widths, dtype=np.uint32)
        scratch = np.empty(theta.size, dtype=np.float32)
        return float(self.lib.bann_cpu_leapfrog(_p(X), _p(y), n, m, w.ctypes.data_as(C.POINTER(C.c_uint32)), len(widths),
                                                _ACT[act], int(lasso), int(stdn), _p(theta), _p(mom), _p(eps), _p(lam),
                                                float(lam_e), L, _p(scratch)))
