"""ctypes binding of the C-ABI in include/gpsiq.h (libgpsiq.so, built in-tree).

There is no fallback: if the shared library is missing or cannot be loaded this
module raises at import time, and ``Synthesizer`` raises if no CUDA device can
be opened.  Nothing here touches ``oracle/``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpsiq.so")

OK = 0
ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY = -1, -2, -3, -4
CARRIER_FLOAT, CARRIER_INT32 = 0, 1
FLAG_RESET_CARRIER = 1
KERNEL_AUTO, KERNEL_LANE_PER_CHANNEL, KERNEL_LINE = 0, 1, 3   # (2 was the retired segment-list kernel)
MAX_LOOKAHEAD = 3   # batches gpsiq_submit* may run ahead of the one being fetched (scan sets - 1)
LINE_DBG_FORCE_CHUNK, LINE_DBG_FORCE_TILE, LINE_DBG_PERTURB = 1, 2, 4
MAX_CHAN = 32
NCO_CODE, NCO_CARRIER = 0, 1
OPT_CHAIN_KEEPS_ESTIMATE, OPT_RENDER_AFTER_NEXT_CHAIN, OPT_LINE_GRID_CAP, OPT_FREE_RUNNING_ESTIMATE = 1, 2, 3, 4

# numpy view of gpsiq_chan_desc (64 bytes, include/gpsiq.h)
DESC_DTYPE = np.dtype(
    [
        ("prn", "<i4"),
        ("ms0", "<i4"),
        ("navbits", "<u8"),
        ("code_phase0", "<f8"),
        ("code_step", "<f8"),
        ("carr_step", "<f8"),
        ("carr_phase0", "<f8"),
        ("gain", "<f8"),
        ("flags", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert DESC_DTYPE.itemsize == 64


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("max_chan", C.c_int32),
        ("samples_per_epoch", C.c_int32),
        ("carrier_mode", C.c_int32),
        ("max_epochs", C.c_int32),
        ("tile_samples", C.c_int32),
        ("kernel", C.c_int32),
        ("reserved", C.c_int32 * 9),
    ]


# every symbol include/gpsiq.h declares: name -> (restype, argtypes)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_d = C.c_double
SYMBOLS = {
    "gpsiq_create": (_i, [C.POINTER(_vp), C.POINTER(Config)]),
    "gpsiq_destroy": (None, [_vp]),
    "gpsiq_synth": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_synth_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gpsiq_get_carrier": (_i, [_vp, _vp]),
    "gpsiq_set_carrier": (_i, [_vp, _vp]),
    "gpsiq_get_carrier_trace": (_i, [_vp, _vp, _i]),
    "gpsiq_device_iq": (_vp, [_vp]),
    "gpsiq_checksum_device": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_make_desc": (_i, [_vp, _i, _i, _d, _d, _d, _d, _d, _vp, _i, _i, _i, _d, _i]),
    "gpsiq_nco_advance": (_i, [_i, C.POINTER(_d), _d, _i64, C.POINTER(_i64)]),
    "gpsiq_carrier_chain_host": (_i, [_vp, _i, _i, _i, _d, _d, _vp, C.POINTER(_d), C.POINTER(_i)]),
    "gpsiq_carrier_slice_host": (_i, [_vp, _i, _i, _i, _d, _d, _vp, C.POINTER(_d), C.POINTER(_i), C.POINTER(_i), _vp, _vp, _vp]),
    "gpsiq_carrier_study_host": (_i, [_vp, _i, _i, _i, _d, _d, _d, _i, _vp, C.POINTER(_i), C.POINTER(_i)]),
    "gpsiq_host_alloc": (_vp, [C.c_size_t]),
    "gpsiq_host_free": (None, [_vp]),
    "gpsiq_launch_count": (_i64, [_vp]),
    "gpsiq_carrier_fallbacks": (_i, [_vp, C.POINTER(_i64)]),
    "gpsiq_slice_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "gpsiq_device_status": (_i, [_vp]),
    "gpsiq_timing_begin": (_i, [_vp]),
    "gpsiq_timing_collect": (_i, [_vp, C.POINTER(_i), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gpsiq_timing_sample_kernel": (_i, [_vp, C.POINTER(_i), C.POINTER(C.c_float), C.POINTER(_i)]),
    "gpsiq_submit": (_i, [_vp, _vp, _i]),
    "gpsiq_fetch": (_i, [_vp, _vp]),
    "gpsiq_submit_device": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_fetch_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_timing_sample_kernel_isolated": (_i, [_vp, _i, C.POINTER(C.c_float), C.POINTER(_i)]),
    "gpsiq_scan_device": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_prepare_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gpsiq_speculate_device": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_chain_device": (_i, [_vp, _vp, _i, _vp]),
    "gpsiq_multi_create": (_i, [C.POINTER(_vp), _vp, _i]),
    "gpsiq_multi_destroy": (None, [_vp]),
    "gpsiq_multi_submit": (_i, [_vp, _vp, _i]),
    "gpsiq_multi_fetch_begin": (_i, [_vp, _vp]),
    "gpsiq_multi_fetch_end": (_i, [_vp]),
    "gpsiq_multi_fetch": (_i, [_vp, _vp]),
    "gpsiq_multi_devices": (_i, [_vp]),
    "gpsiq_multi_launch_count": (_i64, [_vp]),
    "gpsiq_multi_last_error": (C.c_char_p, [_vp]),
    "gpsiq_estimate_fold_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_carrier_fold_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_estimate_anchor_device": (_i, [_vp, _vp]),
    "gpsiq_estimate_from_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_render_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gpsiq_set_option": (_i, [_vp, _i, _i]),
    "gpsiq_estimate_to_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_estimate_correct_device": (_i, [_vp, _vp, _vp, _d, _vp]),
    "gpsiq_carrier_to_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_carrier_from_device": (_i, [_vp, _vp, _vp]),
    "gpsiq_mailbox_create": (_i, [_vp, _vp]),
    "gpsiq_mailbox_open": (_i, [_vp, _vp, _i]),
    "gpsiq_mailbox_send": (_i, [_vp, C.c_uint64, _vp]),
    "gpsiq_mailbox_recv": (_i, [_vp, C.c_uint64, _vp]),
    "gpsiq_chain_handoff_device": (_i, [_vp, _vp, _i, C.c_uint64, C.c_uint64, _vp, _vp]),
    "gpsiq_trace_dump": (_i, [_vp, _i]),
    "gpsiq_line_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "gpsiq_minmod_host": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "gpsiq_line_probe_host": (_i, [_i, _d, _d, _i, C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i)]),
    "gpsiq_line_verify_host": (_i, [_vp, _i, _i, _i, _vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                    C.POINTER(_i64)]),
    "gpsiq_strerror": (C.c_char_p, [_i]),
    "gpsiq_last_error": (C.c_char_p, [_vp]),
    "gpsiq_version": (C.c_char_p, []),
    "gpsiq_get_tables": (None, [_vp, _vp]),
    "gpsiq_get_ca_code": (_i, [_i, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class GpsiqError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = lib.gpsiq_strerror(status).decode()
        super().__init__("%s (%d)%s" % (msg, status, ": " + detail if detail else ""))


def check(status, ctx=None):
    if status != OK:
        raise GpsiqError(status, lib.gpsiq_last_error(ctx).decode())


def version():
    return lib.gpsiq_version().decode()


def tables():
    s = np.zeros(512, np.int32)
    c = np.zeros(512, np.int32)
    lib.gpsiq_get_tables(s.ctypes.data, c.ctypes.data)
    return s, c


def ca_code(prn):
    chips = np.zeros(1023, np.uint8)
    check(lib.gpsiq_get_ca_code(int(prn), chips.ctypes.data))
    return chips


def nco_advance(mode, phase, step, count):
    """Host run of the scan the kernels use; returns (phase, wraps)."""
    x = _d(phase)
    w = _i64(0)
    check(lib.gpsiq_nco_advance(int(mode), C.byref(x), float(step), int(count), C.byref(w)))
    return x.value, w.value


def minmod(b, a, m, n, stop=0):
    """min over x in [0, n) of (b + a*x) mod m, as the safety check of the line kernel computes it (host run)."""
    return int(lib.gpsiq_minmod_host(int(b), int(a), int(m), int(n), int(stop)))


def line_probe(mode, x0, step, n):
    """Literal recurrence vs the line kernel's straight line -> (max deviation, index mismatches, hazard flagged)."""
    dev, mm, hz = _i64(0), _i(0), _i(0)
    check(lib.gpsiq_line_probe_host(int(mode), float(x0), float(step), int(n), C.byref(dev), C.byref(mm), C.byref(hz)))
    return dev.value, mm.value, bool(hz.value)


def line_verify(desc, samples_per_epoch):
    """Host run of the line kernel's index arithmetic + tile check over desc [E][C] -> (tiles, flagged, bad, lag_flagged)."""
    d = np.ascontiguousarray(desc)
    t, f, b, l = _i64(0), _i64(0), _i64(0), _i64(0)
    check(lib.gpsiq_line_verify_host(d.ctypes.data, d.shape[0], d.shape[1], int(samples_per_epoch), None, C.byref(t),
                                     C.byref(f), C.byref(b), C.byref(l)))
    return t.value, f.value, b.value, l.value


def carrier_chain_host(steps, N, T, x0, est_err=0.0):
    """Host run of the speculate/translate/verify carrier scan -> (ck [E][ntiles], x_end, n_fallback)."""
    st = np.ascontiguousarray(steps, dtype=np.float64)
    ntiles = (N + T - 1) // T
    ck = np.zeros((st.size, ntiles), np.float64)
    xe, fb = _d(0), _i(0)
    check(lib.gpsiq_carrier_chain_host(st.ctypes.data, st.size, N, T, float(x0), float(est_err), ck.ctypes.data,
                                       C.byref(xe), C.byref(fb)))
    return ck, xe.value, fb.value


def carrier_slice_host(steps, N, T, x0, est_err=0.0, ties=None, flags=None, phase0=None):
    """Host run of the carrier scan through the slice level (one exact head scan per batch, groups chained from their
    translated starts) -> (ck [E][ntiles], x_end, n_fallback, how): how 1 translated, 0 serial, -2 internal error.
    ties: optional int32 array of 2 -> [group trajectories holding a tie event, translations changed by one]."""
    st = np.ascontiguousarray(steps, dtype=np.float64)
    ntiles = (N + T - 1) // T
    ck = np.zeros((st.size, ntiles), np.float64)
    xe, fb, how = _d(0), _i(0), _i(0)
    fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.int32)
    p0 = None if phase0 is None else np.ascontiguousarray(phase0, dtype=np.float64)
    check(lib.gpsiq_carrier_slice_host(st.ctypes.data, st.size, N, T, float(x0), float(est_err), ck.ctypes.data,
                                       C.byref(xe), C.byref(fb), C.byref(how), None if ties is None else ties.ctypes.data,
                                       None if fl is None else fl.ctypes.data, None if p0 is None else p0.ctypes.data))
    return ck, xe.value, fb.value, how.value


def make_desc(carrier_mode, prn, f_carr, f_code, delt, carr_phase, code_phase, dwrd60, iword, ibit, icode, gain,
              carr_phase_is_new):
    """gpsiq_make_desc for one slot -> numpy record (DESC_DTYPE)."""
    out = np.zeros(1, DESC_DTYPE)
    d = np.ascontiguousarray(dwrd60, dtype=np.uint64)
    assert d.size == 60
    check(
        lib.gpsiq_make_desc(
            out.ctypes.data, int(carrier_mode), int(prn), float(f_carr), float(f_code), float(delt),
            float(carr_phase), float(code_phase), d.ctypes.data, int(iword), int(ibit), int(icode), float(gain),
            int(bool(carr_phase_is_new)),
        )
    )
    return out[0]
