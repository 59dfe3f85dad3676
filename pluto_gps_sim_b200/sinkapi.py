"""ctypes binding of the output transport (include/gpssink.h, part of libgpshost.so).

A `Sink` takes the interleaved int16 I/Q stream in the reference's push units (300000 pairs per 0.1 s epoch,
plutogpssim.c:2146-2158) and discards it, appends it to a file, or hands it to an ADALM-Pluto through libiio
(dlopen()ed by the library when a radio sink is opened; `SinkError` with ERR_BACKEND if it is not installed)."""
import ctypes as C
import os

import numpy as np

from .hostapi import LIB_PATH

ERR_ARG, ERR_IO, ERR_BACKEND, ERR_DEVICE, ERR_PUSH = -1, -2, -3, -4, -5
PUSH_PAIRS = 300000


class RadioConfig(C.Structure):
    _fields_ = [
        ("uri", C.c_char_p), ("hostname", C.c_char_p), ("gain_db", C.c_double), ("bw_hz", C.c_longlong),
        ("fs_hz", C.c_longlong), ("lo_hz", C.c_longlong), ("rfport", C.c_char_p), ("kernel_buffers", C.c_int32),
        ("pairs_per_push", C.c_int32), ("iio_lib", C.c_char_p), ("ad9361_lib", C.c_char_p), ("reserved", C.c_int32 * 8),
    ]


SYMBOLS = {
    "gpssink_radio_defaults": (None, [C.POINTER(RadioConfig)]),
    "gpssink_radio_option": (C.c_int, [C.POINTER(RadioConfig), C.c_int, C.c_char_p]),
    "gpssink_open_null": (C.c_int, [C.POINTER(C.c_void_p)]),
    "gpssink_open_file": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p]),
    "gpssink_open_radio": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(RadioConfig)]),
    "gpssink_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gpssink_submit": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gpssink_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "gpssink_abort": (C.c_int, [C.c_void_p]),
    "gpssink_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gpssink_close": (C.c_int, [C.c_void_p]),
    "gpssink_last_error": (C.c_char_p, []),
}

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class SinkError(RuntimeError):
    def __init__(self, status):
        self.status = status
        super().__init__("gpssink error %d: %s" % (status, lib.gpssink_last_error().decode()))


def radio_config(options=(), iio_lib=None):
    """Reference defaults (plutogpssim.c:2270-2276) with the reference's radio options applied in order:
    options = [("A", "-35.5"), ("B", "3.0"), ("U", "usb:1.2.5"), ("N", "host"), ("s", "2600000")]."""
    cfg = RadioConfig()
    lib.gpssink_radio_defaults(C.byref(cfg))
    cfg._keep = []
    for letter, arg in options:
        b = str(arg).encode()
        cfg._keep.append(b)
        rc = lib.gpssink_radio_option(C.byref(cfg), ord(letter), b)
        if rc != 0:
            raise SinkError(rc)
    if iio_lib is not None:
        b = os.fsencode(iio_lib)
        cfg._keep.append(b)
        cfg.iio_lib = b
    return cfg


class Sink:
    """Sink() = discard; Sink(path=...) = file ("-" = stdout); Sink(radio=radio_config(...)) = ADALM-Pluto via libiio."""

    def __init__(self, path=None, radio=None):
        self._h = C.c_void_p()
        self._keep = {}
        if radio is not None:
            rc = lib.gpssink_open_radio(C.byref(self._h), C.byref(radio))
        elif path is not None:
            rc = lib.gpssink_open_file(C.byref(self._h), os.fsencode(path))
        else:
            rc = lib.gpssink_open_null(C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise SinkError(rc)

    @staticmethod
    def _check(iq):
        a = np.asarray(iq)
        if a.dtype != np.int16 or not a.flags.c_contiguous or a.size % 2:
            raise ValueError("I/Q must be a C-contiguous int16 array of interleaved pairs")
        return a

    def push(self, iq):
        a = self._check(iq)
        rc = lib.gpssink_push(self._h, a.ctypes.data, a.size // 2)
        if rc != 0:
            raise SinkError(rc)

    def submit(self, iq):
        """Queue a batch for the writer thread; the array is kept alive until wait(ticket)."""
        a = self._check(iq)
        t = lib.gpssink_submit(self._h, a.ctypes.data, a.size // 2)
        if t < 0:
            raise SinkError(int(t))
        self._keep[t] = a
        return t

    def wait(self, ticket):
        rc = lib.gpssink_wait(self._h, ticket)
        self._keep.pop(ticket, None)
        if rc != 0:
            raise SinkError(rc)

    def abort(self):
        """Stop after the push unit in progress; queued and later batches are discarded."""
        lib.gpssink_abort(self._h)

    @property
    def stats(self):
        p, n = C.c_int64(0), C.c_int64(0)
        lib.gpssink_stats(self._h, C.byref(p), C.byref(n))
        return p.value, n.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            h, self._h = self._h, C.c_void_p()
            rc = lib.gpssink_close(h)
            self._keep.clear()
            if rc != 0:
                raise SinkError(rc)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
