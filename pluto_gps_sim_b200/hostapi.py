"""ctypes binding of the host orchestrator (include/gpshost.h, libgpshost.so: plain C++, no CUDA).

`Scenario` yields the per-epoch channel descriptors the sample kernels consume, computed from a RINEX-2
navigation file exactly as the reference computes its per-epoch channel state (plutogpssim.c:2476-2806
minus the sample loop)."""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpshost.so")
SIM_PATH = os.path.join(_HERE, "gpsiq_sim")

POS_LLH, POS_XYZ, POS_MOTION = 0, 1, 2
ERR_ARG, ERR_NAVFILE, ERR_NOEPH, ERR_MOTION, ERR_TIME = -1, -2, -3, -4, -5


class HostConfig(C.Structure):
    _fields_ = [
        ("nav_path", C.c_char_p), ("pos_mode", C.c_int32), ("pos", C.c_double * 3), ("motion_path", C.c_char_p),
        ("have_start", C.c_int32), ("start", C.c_int32 * 5), ("start_sec", C.c_double), ("time_overwrite", C.c_int32),
        ("iono_disable", C.c_int32), ("sample_rate", C.c_int64), ("max_chan", C.c_int32), ("carrier_mode", C.c_int32),
        ("rinex3", C.c_int32), ("threads", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


SYMBOLS = {
    "gpshost_open": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(HostConfig)]),
    "gpshost_close": (None, [C.c_void_p]),
    "gpshost_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpshost_skip": (C.c_int, [C.c_void_p, C.c_int]),
    "gpshost_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "gpshost_describe_iono": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "gpshost_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "gpshost_last_error": (C.c_char_p, []),
    "gpshost_parity": (C.c_uint32, [C.c_uint32, C.c_int]),
    "gpshost_date2gps": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "gpshost_llh2xyz": (None, [C.c_void_p, C.c_void_p]),
    "gpshost_xyz2llh": (None, [C.c_void_p, C.c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class HostError(RuntimeError):
    def __init__(self, status):
        self.status = status
        super().__init__("gpshost error %d: %s" % (status, lib.gpshost_last_error().decode()))


class Scenario:
    """nav: RINEX-2 navigation file (plain or gz); rinex3=True: a RINEX-3 one (the reference's -3).  Exactly one of llh (deg, deg, m), xyz (ECEF m), motion (csv path).
    start: (y, m, d, hh, mm, sec) or None (first TOC of the file, the reference's default)."""

    def __init__(self, nav, llh=None, xyz=None, motion=None, start=None, time_overwrite=False, iono=True,
                 sample_rate=3000000, max_chan=12, carrier_mode=capi.CARRIER_FLOAT, rinex3=False, threads=0):
        cfg = HostConfig()
        self._keep = [os.fsencode(nav), os.fsencode(motion) if motion else None]
        cfg.nav_path = self._keep[0]
        if motion:
            cfg.pos_mode, cfg.motion_path = POS_MOTION, self._keep[1]
        elif xyz is not None:
            cfg.pos_mode = POS_XYZ
            cfg.pos[:] = list(xyz)
        else:
            cfg.pos_mode = POS_LLH
            cfg.pos[:] = list(llh if llh is not None else (35.681298, 139.766247, 10.0))
        if start is not None:
            cfg.have_start = 1
            cfg.start[:] = [int(v) for v in start[:5]]
            cfg.start_sec = float(start[5])
        cfg.time_overwrite = int(bool(time_overwrite))
        cfg.iono_disable = int(not iono)
        cfg.sample_rate = int(sample_rate)
        cfg.max_chan = int(max_chan)
        cfg.carrier_mode = int(carrier_mode)
        cfg.rinex3 = int(bool(rinex3))
        cfg.threads = int(threads)       # 0 = default (min(8, cores)), 1 = serial; same descriptors either way
        self.max_chan = int(max_chan)
        self._h = C.c_void_p()
        rc = lib.gpshost_open(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            raise HostError(rc)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.gpshost_close(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def next(self, n_epochs):
        """-> descriptors [n_epochs][max_chan] (capi.DESC_DTYPE) of the next n_epochs 0.1 s epochs."""
        d = np.zeros((n_epochs, self.max_chan), capi.DESC_DTYPE)
        rc = lib.gpshost_next(self._h, d.ctypes.data, n_epochs)
        if rc != 0:
            raise HostError(rc)
        return d

    def skip(self, n_epochs):
        """Advance n_epochs without producing descriptors (one evaluated epoch per 30 s refresh interval)."""
        rc = lib.gpshost_skip(self._h, int(n_epochs))
        if rc != 0:
            raise HostError(rc)

    def describe(self):
        buf = C.create_string_buffer(8192)
        lib.gpshost_describe(self._h, buf, len(buf))
        return buf.value.decode()

    def describe_iono(self):
        """The reference's -v block: ionosphere / UTC header parameters ('' for an incomplete header)."""
        buf = C.create_string_buffer(1024)
        lib.gpshost_describe_iono(self._h, buf, len(buf))
        return buf.value.decode()

    @property
    def time(self):
        w, s = C.c_int(0), C.c_double(0)
        lib.gpshost_time(self._h, C.byref(w), C.byref(s))
        return w.value, s.value


def parity(source, nib=0):
    return int(lib.gpshost_parity(int(source) & 0xFFFFFFFF, int(nib)))


def date2gps(y, m, d, hh, mm, sec):
    w, s = C.c_int(0), C.c_double(0)
    lib.gpshost_date2gps(y, m, d, hh, mm, float(sec), C.byref(w), C.byref(s))
    return w.value, s.value


def llh2xyz(lat_deg, lon_deg, h):
    a = np.array([lat_deg / 57.2957795131, lon_deg / 57.2957795131, h], np.float64)   # the reference's R2D
    o = np.zeros(3)
    lib.gpshost_llh2xyz(a.ctypes.data, o.ctypes.data)
    return o


def xyz2llh(xyz):
    a = np.ascontiguousarray(xyz, np.float64)
    o = np.zeros(3)
    lib.gpshost_xyz2llh(a.ctypes.data, o.ctypes.data)
    return o


class SliceFeeder:
    """Host side of multi-GPU time slicing (timeslice.py): the descriptors of rank `rank`'s
    slices of ONE scenario.  Step k of a `world`-rank job covers epochs [k*world*E, (k+1)*world*E); this rank owns
    [(k*world + rank)*E, +E).  The other ranks' epochs are skipped (gpshost_skip: one evaluated epoch per 30 s refresh
    interval), so every rank pays for its own slices only and the slices of all ranks, put back in stream order, are
    byte for byte the descriptors of a single run."""

    def __init__(self, rank, world, epochs_per_slice, **scenario_kw):
        if not (0 <= rank < world) or epochs_per_slice < 1:
            raise ValueError("rank/world/epochs_per_slice")
        self.rank, self.world, self.E = int(rank), int(world), int(epochs_per_slice)
        self.scenario = Scenario(**scenario_kw)
        self.scenario.skip(self.rank * self.E)

    def next_slice(self):
        d = self.scenario.next(self.E)
        self.scenario.skip((self.world - 1) * self.E)
        return d

    def close(self):
        self.scenario.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
