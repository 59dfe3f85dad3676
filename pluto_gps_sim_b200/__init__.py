"""B200-native GPS L1 C/A baseband I/Q synthesis (the hot path of
Mictronics/pluto-gps-sim, plutogpssim.c:2689-2756) behind a C-ABI.

  capi   ctypes binding of include/gpsiq.h (libgpsiq.so; no fallback)
  synth  Synthesizer: descriptors in, int16 I/Q out, carrier phase carried
"""
from . import capi  # noqa: F401  (raises if libgpsiq.so is missing)
from .synth import NUM_SAMPLES, Synthesizer, checksum_host  # noqa: F401

__all__ = ["capi", "Synthesizer", "checksum_host", "NUM_SAMPLES"]
