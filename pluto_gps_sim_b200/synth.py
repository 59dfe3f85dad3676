"""Host-side handle on the CUDA synthesizer (one context per GPU).

The reference has no operator interface for its hot loop (it is inline in
``main``, plutogpssim.c:2689-2756); this class is the Python face of the seam
defined in include/gpsiq.h: per-epoch channel descriptors in, interleaved int16
I/Q out, carrier phase carried between calls exactly like ``chan[i].carr_phase``.
All arithmetic happens in libgpsiq.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import capi

NUM_SAMPLES = 300000  # plutogpssim.c:43-44: TX_SAMPLE_FREQ/10, independent of -s


class Synthesizer:
    def __init__(self, max_chan=12, samples_per_epoch=NUM_SAMPLES, max_epochs=100, carrier_mode=capi.CARRIER_FLOAT,
                 device=0, tile_samples=0, kernel=capi.KERNEL_AUTO, serial_carrier_scan=False, line_debug=0):
        cfg = capi.Config()
        cfg.device = device
        cfg.max_chan = max_chan
        cfg.samples_per_epoch = samples_per_epoch
        cfg.carrier_mode = carrier_mode
        cfg.max_epochs = max_epochs
        cfg.tile_samples = tile_samples
        cfg.kernel = kernel
        cfg.reserved[0] = 1 if serial_carrier_scan else 0
        cfg.reserved[1] = int(line_debug)   # test hooks of the line kernel (capi.LINE_DBG_*)
        self._ctx = C.c_void_p()
        capi.check(capi.lib.gpsiq_create(C.byref(self._ctx), C.byref(cfg)))
        self.max_chan = max_chan
        self.samples_per_epoch = samples_per_epoch
        self.max_epochs = max_epochs
        self.carrier_mode = carrier_mode
        self.device = device

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            capi.lib.gpsiq_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- synthesis --------------------------------------------------------
    def _desc_array(self, desc):
        d = np.ascontiguousarray(desc, dtype=capi.DESC_DTYPE)
        if d.ndim == 1:
            d = d.reshape(-1, self.max_chan)
        if d.ndim != 2 or d.shape[1] != self.max_chan:
            raise ValueError("descriptors must be [n_epochs][max_chan=%d]" % self.max_chan)
        return d

    def synth(self, desc, out=None, keep_on_device=False):
        """Host buffers in, host buffers out (H2D + kernels + D2H, blocking).

        desc: [n_epochs][max_chan] records of capi.DESC_DTYPE.
        Returns int16 array [n_epochs, samples_per_epoch, 2] (I, Q)."""
        d = self._desc_array(desc)
        n = d.shape[0]
        if keep_on_device:
            capi.check(capi.lib.gpsiq_synth(self._ctx, d.ctypes.data, n, None), self._ctx)
            return None
        if out is None:
            out = np.empty((n, self.samples_per_epoch, 2), np.int16)
        assert out.dtype == np.int16 and out.size >= n * self.samples_per_epoch * 2 and out.flags.c_contiguous
        capi.check(capi.lib.gpsiq_synth(self._ctx, d.ctypes.data, n, out.ctypes.data), self._ctx)
        return out

    def synth_ptr(self, desc_ptr, n_epochs, iq_ptr):
        """Raw host pointers (e.g. pinned memory) -- the call bench.py times end to end."""
        capi.check(capi.lib.gpsiq_synth(self._ctx, desc_ptr, n_epochs, iq_ptr), self._ctx)

    def synth_device(self, desc_dev_ptr, n_epochs, iq_dev_ptr, stream_ptr=None):
        """All-device, asynchronous on `stream_ptr` (a cudaStream_t as int)."""
        capi.check(capi.lib.gpsiq_synth_device(self._ctx, desc_dev_ptr, n_epochs, iq_dev_ptr, stream_ptr), self._ctx)

    def submit(self, desc):
        """Host descriptors ([n_epochs][max_chan] array or raw pointer via submit_ptr); returns immediately."""
        d = self._desc_array(desc)
        capi.check(capi.lib.gpsiq_submit(self._ctx, d.ctypes.data, d.shape[0]), self._ctx)
        return d.shape[0]

    def submit_ptr(self, desc_ptr, n_epochs):
        capi.check(capi.lib.gpsiq_submit(self._ctx, desc_ptr, n_epochs), self._ctx)

    def fetch_ptr(self, iq_ptr):
        """Blocks until the oldest submitted batch is complete in host memory at iq_ptr."""
        capi.check(capi.lib.gpsiq_fetch(self._ctx, iq_ptr), self._ctx)

    def fetch(self, n_epochs, out=None):
        if out is None:
            out = np.empty((n_epochs, self.samples_per_epoch, 2), np.int16)
        capi.check(capi.lib.gpsiq_fetch(self._ctx, out.ctypes.data), self._ctx)
        return out

    def submit_device(self, desc_dev_ptr, n_epochs, after_stream_ptr=None):
        """Scan a batch ahead of time (asynchronous, on the context's own stream)."""
        capi.check(capi.lib.gpsiq_submit_device(self._ctx, desc_dev_ptr, n_epochs, after_stream_ptr), self._ctx)

    def fetch_device(self, iq_dev_ptr, stream_ptr=None):
        """Render the oldest submitted batch on `stream_ptr` (asynchronous)."""
        capi.check(capi.lib.gpsiq_fetch_device(self._ctx, iq_dev_ptr, stream_ptr), self._ctx)

    def device_iq_ptr(self):
        return capi.lib.gpsiq_device_iq(self._ctx)

    def checksum_device(self, iq_dev_ptr, n_epochs):
        sums = np.zeros(n_epochs, np.uint64)
        capi.check(capi.lib.gpsiq_checksum_device(self._ctx, iq_dev_ptr, n_epochs, sums.ctypes.data), self._ctx)
        return sums

    # -- carrier state ----------------------------------------------------
    @property
    def carrier(self):
        p = np.zeros(self.max_chan, np.float64)
        capi.check(capi.lib.gpsiq_get_carrier(self._ctx, p.ctypes.data), self._ctx)
        return p

    @carrier.setter
    def carrier(self, value):
        p = np.ascontiguousarray(value, dtype=np.float64)
        assert p.size == self.max_chan
        capi.check(capi.lib.gpsiq_set_carrier(self._ctx, p.ctypes.data), self._ctx)

    def carrier_trace(self, n_epochs):
        t = np.zeros((n_epochs, self.max_chan), np.float64)
        capi.check(capi.lib.gpsiq_get_carrier_trace(self._ctx, t.ctypes.data, n_epochs), self._ctx)
        return t

    # -- bookkeeping ------------------------------------------------------
    @property
    def launch_count(self):
        return int(capi.lib.gpsiq_launch_count(self._ctx))

    @property
    def carrier_fallbacks(self):
        n = C.c_int64(0)
        capi.check(capi.lib.gpsiq_carrier_fallbacks(self._ctx, C.byref(n)), self._ctx)
        return n.value

    @property
    def slice_stats(self):
        """-> ((slot, batch) carrier chains passed by the slice-level translation, chained serially)."""
        a, b = C.c_int64(0), C.c_int64(0)
        capi.check(capi.lib.gpsiq_slice_stats(self._ctx, C.byref(a), C.byref(b)), self._ctx)
        return a.value, b.value

    def check_device(self):
        """Wait for the device; raise if a kernel flagged an error since the last check."""
        capi.check(capi.lib.gpsiq_device_status(self._ctx), self._ctx)

    @property
    def line_stats(self):
        """-> (tile-slot pairs re-checked with the literal recurrence, samples patched, chunks flagged)."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        capi.check(capi.lib.gpsiq_line_stats(self._ctx, C.byref(a), C.byref(b), C.byref(c)), self._ctx)
        return a.value, b.value, c.value

    def timing_begin(self):
        capi.check(capi.lib.gpsiq_timing_begin(self._ctx), self._ctx)

    def timing_collect(self):
        """-> (recorded calls, scan-phase ms, synthesis-kernel ms), CUDA events on the launching stream."""
        n, a, b = C.c_int(0), C.c_float(0), C.c_float(0)
        capi.check(capi.lib.gpsiq_timing_collect(self._ctx, C.byref(n), C.byref(a), C.byref(b)), self._ctx)
        return n.value, a.value, b.value

    def timing_sample_kernel(self):
        """-> (launches, summed ms, epochs per launch) of the dominant kernel (k_synth_line) alone."""
        n, a, e = C.c_int(0), C.c_float(0), C.c_int(0)
        capi.check(capi.lib.gpsiq_timing_sample_kernel(self._ctx, C.byref(n), C.byref(a), C.byref(e)), self._ctx)
        return n.value, a.value, e.value

    def timing_sample_kernel_isolated(self, reps=20):
        """-> (mean ms, epochs per launch) of the dominant kernel re-launched alone on an idle device."""
        a, e = C.c_float(0), C.c_int(0)
        capi.check(capi.lib.gpsiq_timing_sample_kernel_isolated(self._ctx, reps, C.byref(a), C.byref(e)), self._ctx)
        return a.value, e.value

    # -- time-slice phases (multi-GPU) --------------------------------------
    def scan_device(self, desc_dev_ptr, n_epochs, stream_ptr=None):
        capi.check(capi.lib.gpsiq_scan_device(self._ctx, desc_dev_ptr, n_epochs, stream_ptr), self._ctx)

    def prepare_device(self, desc_dev_ptr, n_epochs, advance_dev_ptr=None, stream_ptr=None):
        capi.check(capi.lib.gpsiq_prepare_device(self._ctx, desc_dev_ptr, n_epochs, advance_dev_ptr, stream_ptr), self._ctx)

    def speculate_device(self, desc_dev_ptr, n_epochs, stream_ptr=None):
        capi.check(capi.lib.gpsiq_speculate_device(self._ctx, desc_dev_ptr, n_epochs, stream_ptr), self._ctx)

    def chain_device(self, desc_dev_ptr, n_epochs, stream_ptr=None):
        capi.check(capi.lib.gpsiq_chain_device(self._ctx, desc_dev_ptr, n_epochs, stream_ptr), self._ctx)

    def estimate_fold_device(self, advance_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_estimate_fold_device(self._ctx, advance_dev_ptr, stream_ptr), self._ctx)

    def carrier_fold_device(self, advance_dev_ptr, stream_ptr=None):
        """Integer carrier only: exact carrier state <- fold(state, advance of a slice synthesized elsewhere)."""
        capi.check(capi.lib.gpsiq_carrier_fold_device(self._ctx, advance_dev_ptr, stream_ptr), self._ctx)

    def estimate_from_device(self, src_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_estimate_from_device(self._ctx, src_dev_ptr, stream_ptr), self._ctx)

    def estimate_anchor_device(self, stream_ptr=None):
        capi.check(capi.lib.gpsiq_estimate_anchor_device(self._ctx, stream_ptr), self._ctx)

    def render_device(self, desc_dev_ptr, n_epochs, iq_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_render_device(self._ctx, desc_dev_ptr, n_epochs, iq_dev_ptr, stream_ptr), self._ctx)

    def set_option(self, option, value):
        capi.check(capi.lib.gpsiq_set_option(self._ctx, int(option), int(value)), self._ctx)

    def estimate_to_device(self, dst_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_estimate_to_device(self._ctx, dst_dev_ptr, stream_ptr), self._ctx)

    def estimate_correct_device(self, exact_old_ptr, est_old_ptr, gain, stream_ptr=None):
        capi.check(capi.lib.gpsiq_estimate_correct_device(self._ctx, exact_old_ptr, est_old_ptr, float(gain), stream_ptr),
                   self._ctx)

    def carrier_to_device(self, dst_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_carrier_to_device(self._ctx, dst_dev_ptr, stream_ptr), self._ctx)

    def carrier_from_device(self, src_dev_ptr, stream_ptr=None):
        capi.check(capi.lib.gpsiq_carrier_from_device(self._ctx, src_dev_ptr, stream_ptr), self._ctx)

    # SM-free carrier hand-off between the GPUs of one node (include/gpsiq.h: gpsiq_mailbox_*)
    def mailbox_create(self):
        """-> the 64-byte CUDA IPC handle of this context's mailbox."""
        import ctypes
        h = ctypes.create_string_buffer(64)
        capi.check(capi.lib.gpsiq_mailbox_create(self._ctx, h), self._ctx)
        return h.raw

    def mailbox_open(self, handle, peer_device):
        import ctypes
        capi.check(capi.lib.gpsiq_mailbox_open(self._ctx, ctypes.create_string_buffer(bytes(handle), 64), int(peer_device)),
                   self._ctx)

    def mailbox_send(self, seq, stream_ptr=None):
        capi.check(capi.lib.gpsiq_mailbox_send(self._ctx, int(seq), stream_ptr), self._ctx)

    def chain_handoff_device(self, desc_dev_ptr, n_epochs, recv_seq, send_seq, start_copy_ptr=None, stream_ptr=None):
        """gpsiq_chain_handoff_device: [wait for message recv_seq] -> chain -> [send message send_seq], one kernel."""
        capi.check(capi.lib.gpsiq_chain_handoff_device(self._ctx, desc_dev_ptr, n_epochs, int(recv_seq), int(send_seq),
                                                       start_copy_ptr, stream_ptr), self._ctx)

    def mailbox_recv(self, seq, stream_ptr=None):
        capi.check(capi.lib.gpsiq_mailbox_recv(self._ctx, int(seq), stream_ptr), self._ctx)


def checksum_host(iq):
    """numpy mirror of gpsiq_checksum_device for one epoch (int16 [N,2] or [2N])."""
    w = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1).view(np.uint32).astype(np.uint64)
    idx = np.arange(w.size, dtype=np.uint64)
    x = (idx << np.uint64(32)) | w
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        return np.uint64(np.sum(x, dtype=np.uint64))
