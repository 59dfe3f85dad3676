"""TEST INFRASTRUCTURE: ctypes access to oracle/liboracle.so (the CPU
restatement of the reference loop) and loaders for the committed goldens.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use this."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
ORACLE_DIR = os.path.join(REPO, "oracle")
sys.path.insert(0, REPO)

DESC_DTYPE = np.dtype(
    [("prn", "<i4"), ("ms0", "<i4"), ("navbits", "<u8"), ("code_phase0", "<f8"), ("code_step", "<f8"),
     ("carr_step", "<f8"), ("carr_phase0", "<f8"), ("gain", "<f8"), ("flags", "<u4"), ("reserved", "<u4")]
)

_lib = None


def oracle():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
        _lib.oracle_synth.restype = C.c_int
        _lib.oracle_synth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_code_nco.argtypes = [C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        _lib.oracle_carr_nco.argtypes = [C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_double)]
        _lib.oracle_get_tables.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_get_ca.argtypes = [C.c_int, C.c_void_p]
    return _lib


def oracle_synth(desc, samples_per_epoch, carrier_mode=0, carr_state=None):
    """desc [E][C] -> (iq int16 [E, N, 2], carrier trace [E][C]); carr_state updated in place if given."""
    d = np.ascontiguousarray(desc, dtype=DESC_DTYPE)
    E, Cn = d.shape
    iq = np.zeros((E, samples_per_epoch, 2), np.int16)
    trace = np.zeros((E, Cn), np.float64)
    st = np.zeros(Cn, np.float64) if carr_state is None else carr_state
    rc = oracle().oracle_synth(d.ctypes.data, E, Cn, samples_per_epoch, carrier_mode, st.ctypes.data, iq.ctypes.data,
                               trace.ctypes.data)
    assert rc == 0
    return iq, trace


def oracle_code_nco(phase, step, n):
    p, w = C.c_double(0), C.c_int64(0)
    oracle().oracle_code_nco(phase, step, n, C.byref(p), C.byref(w))
    return p.value, w.value


def oracle_carr_nco(phase, step, n):
    p = C.c_double(0)
    oracle().oracle_carr_nco(phase, step, n, C.byref(p))
    return p.value


def oracle_tables():
    s, c = np.zeros(512, np.int32), np.zeros(512, np.int32)
    oracle().oracle_get_tables(s.ctypes.data, c.ctypes.data)
    return s, c


def oracle_ca(prn):
    ch = np.zeros(1023, np.uint8)
    oracle().oracle_get_ca(prn, ch.ctypes.data)
    return ch


def load_golden_desc(name):
    return np.load(os.path.join(GOLDEN, name + "_desc.npy"))


def load_golden_meta(name):
    with open(os.path.join(GOLDEN, name + "_meta.json")) as f:
        return json.load(f)


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
