"""Multi-GPU partition of one sample stream by contiguous time slices.

The reference is one sequential stream; everything its per-sample loop consumes
is recomputed from time at every 0.1 s epoch EXCEPT ``chan[i].carr_phase``,
which chains across all epochs (plutogpssim.c:2741-2746).
So the stream shards along the epoch axis with exactly one exchange value per
channel: rank r owns slice r of every step (epochs [r*E, (r+1)*E) of the step's
N*E epochs) and needs the carrier phases at the end of the previous slice.

Per step, rank r:
    1. receives max_chan doubles from rank r-1 (rank 0: from rank N-1's slice
       of the previous step -- the stream is continuous), except at stream start;
    2. runs the scan phase (exact NCO state at every tile boundary) -- this is
       what advances the carrier phase;
    3. sends its end phases to rank r+1 *before* starting the per-sample work;
    4. renders its slice (the bulk of the work) while later ranks scan.

The only collective traffic is max_chan doubles per slice boundary (point to
point, latency bound); there is deliberately no bulk collective (NVLink
bandwidth is irrelevant to this path).  In FLOAT carrier mode the scan phases of
the N ranks are therefore serialised (a relay) and the renders overlap; in INT32
mode the phase advance of a slice is a closed form, so the relay costs nothing.

``engine`` abstracts the device: ``Synthesizer``-backed on GPUs (GpuSliceEngine),
an oracle-backed stand-in in the gloo/CPU tests of this logic.
"""
import torch
import torch.distributed as dist


class GpuSliceEngine:
    """Synthesizer + torch CUDA tensors; everything is enqueued on the current stream."""

    def __init__(self, synth):
        self.s = synth
        self.phase = torch.zeros(synth.max_chan, dtype=torch.float64, device="cuda")

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def load_carrier(self):            # self.phase -> engine state
        self.s.carrier_from_device(self.phase.data_ptr(), self._stream())

    def store_carrier(self):           # engine state -> self.phase
        self.s.carrier_to_device(self.phase.data_ptr(), self._stream())

    def scan(self, desc_dev, n_epochs):
        self.s.scan_device(desc_dev.data_ptr(), n_epochs, self._stream())

    def render(self, desc_dev, n_epochs, out_dev):
        self.s.render_device(desc_dev.data_ptr(), n_epochs, out_dev.data_ptr(), self._stream())


class TimeSliceRunner:
    def __init__(self, engine, rank=None, world=None):
        self.engine = engine
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.step_index = 0

    def step(self, desc, n_epochs, out):
        """Synthesize this rank's slice of the next step."""
        eng, r, n = self.engine, self.rank, self.world
        if n > 1:
            first = self.step_index == 0 and r == 0          # stream start: phases come from RESET descriptors
            if not first:
                dist.recv(eng.phase, src=(r - 1) % n)
                eng.load_carrier()
        eng.scan(desc, n_epochs)
        if n > 1:
            eng.store_carrier()
            dist.send(eng.phase, dst=(r + 1) % n)
        eng.render(desc, n_epochs, out)
        self.step_index += 1

    def finish(self):
        """Drain the last hand-off (rank 0 receives the phases after the final slice)."""
        if self.world > 1 and self.rank == 0 and self.step_index > 0:
            dist.recv(self.engine.phase, src=self.world - 1)
            self.engine.load_carrier()
