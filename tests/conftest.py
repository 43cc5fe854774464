import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/liboracle.so) -- the checker for every parity test."""
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(path)
    lib.or_hsum_f32_seq.restype = ctypes.c_double
    lib.or_hsum_f32_f64.restype = ctypes.c_double
    lib.or_hmin_f32.restype = ctypes.c_float
    lib.or_hmax_f32.restype = ctypes.c_float
    lib.or_hsum_u32.restype = ctypes.c_uint32
    lib.or_sin.restype = lib.or_cos.restype = lib.or_exp.restype = lib.or_log.restype = ctypes.c_float
    return lib


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference CPU path (oracle/_ref/libenoki_ref.so); None when absent."""
    path = os.path.join(ROOT, "oracle", "_ref", "libenoki_ref.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.ref_info.restype = ctypes.c_char_p
    lib.ref_hsum_f32.restype = lib.ref_hprod_f32.restype = ctypes.c_float
    lib.ref_hmin_f32.restype = lib.ref_hmax_f32.restype = ctypes.c_float
    lib.ref_hsum_u32.restype = ctypes.c_uint32
    return lib


@pytest.fixture(scope="session")
def P():
    return _P


def ulp_diff(a, b):
    """ULP distance as in tests/test.h:155-171 of the reference (NaN == NaN)."""
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    ai = a.view(np.int32).astype(np.int64); bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7fffffff), ai); bi = np.where(bi < 0, -(bi & 0x7fffffff), bi)
    d = np.abs(ai - bi)
    d[np.isnan(a) & np.isnan(b)] = 0
    return d


@pytest.fixture(scope="session")
def ulp():
    return ulp_diff


@pytest.fixture(scope="session")
def ek():
    import enoki_b200
    return enoki_b200


@pytest.fixture()
def gpu(ek):
    if ek.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return ek
