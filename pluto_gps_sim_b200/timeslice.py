"""Multi-GPU partition of one sample stream by contiguous time slices.

The reference is one sequential stream; everything its per-sample loop consumes
is recomputed from time at every 0.1 s epoch EXCEPT ``chan[i].carr_phase``,
which chains across all epochs (plutogpssim.c:2741-2746).
So the stream shards along the epoch axis with exactly one exchange value per
channel: rank r owns slice r of every step (epochs [r*E, (r+1)*E) of the step's
N*E epochs) and needs the carrier phases at the end of the previous slice.

Per step, rank r:
    1. prepare: LUTs/tables and the slice's CLOSED-FORM phase advance per channel;
       one tiny all_gather shares the advances, and every rank folds the advances
       of the slices between its previous slice and this one into its start-phase
       ESTIMATE;
    2. speculate: code-NCO scan + speculative carrier scans from that estimate --
       no exact phase needed, so all ranks do this at the same time;
    3. receives max_chan exact phases from rank r-1 (rank 0: from rank N-1's slice
       of the previous step -- the stream is continuous), except at stream start;
    4. chain: the exact phase is chained through the slice (O(one carrier cycle)
       per epoch: the only serial work of the whole scheme);
    5. sends its end phases to rank r+1 *before* starting the per-sample work;
    6. renders its slice (the bulk of the work) while later ranks chain.

The collective traffic is 2*max_chan doubles per rank per step in the all_gather
plus max_chan doubles per slice boundary point to point (latency bound);
there is deliberately no bulk collective (NVLink bandwidth is irrelevant to this
path).  Estimates only affect speed: a poor one makes epochs fall back to the
serial scan inside `chain`, it never changes a sample.

``engine`` abstracts the device: ``Synthesizer``-backed on GPUs (GpuSliceEngine),
an oracle-backed stand-in in the gloo/CPU tests of this logic.
"""
import contextlib
import os

import torch
import torch.distributed as dist


class GpuSliceEngine:
    """Synthesizer + torch CUDA tensors; everything is enqueued on the current stream."""

    def __init__(self, synth):
        self.s = synth
        self.phase = torch.zeros(synth.max_chan, dtype=torch.float64, device="cuda")
        self.adv = torch.zeros(2 * synth.max_chan, dtype=torch.float64, device="cuda")
        # scan phases and the hand-offs run on their own stream, ahead of the render stream: with the
        # context's two scan sets the next slice is prepared / speculated / chained while this one renders
        self.scan_stream = torch.cuda.Stream(priority=-1)   # small latency-bound kernels: ahead of the sample kernel
        # (GPSIQ_OPT_RENDER_AFTER_NEXT_CHAIN was measured and is NOT enabled: with two scan sets the next prepare
        # waits for this render, so delaying the render behind the ring serialises speculation, ring and rendering:
        # 6.6 ms instead of 4.3 ms per step at 2 GPUs)
        # estimate feedback: the closed-form advance of the slices a rank does not own carries a small systematic
        # error (predicted vs actual rounding drift, ~1e-14 cycles per epoch); `bias` integrates, per slot, the
        # difference between the start-phase estimate a slice was speculated from and the exact phase received
        # for it later, and is subtracted from the next estimate.  Speed only: never part of a result.
        z = lambda: torch.zeros(synth.max_chan, dtype=torch.float64, device="cuda")
        self.bias, self.est_start, self.zero = z(), z(), z()
        # pipelined runner: the NEXT slice's prepare + all_gather run on their own stream beside this slice's
        # speculation and hand-off (nothing in them depends on the ring)
        self.prep_stream = torch.cuda.Stream(priority=-1)
        self.prep_event = torch.cuda.Event()
        # free-running runner: the hop (one fused kernel per slice) has a stream of its own, so that the speculation of
        # the next slice does not queue behind a hop that is still waiting for the previous GPU
        self.hop_stream = torch.cuda.Stream(priority=-1)

    def set_free_running(self, on=True):
        """GPSIQ_OPT_FREE_RUNNING_ESTIMATE: the start-phase estimate follows the slice-level speculation instead of
        being re-anchored by the exact chain (the library corrects its accumulated error open loop)."""
        from . import capi
        self.s.set_option(capi.OPT_FREE_RUNNING_ESTIMATE, 1 if on else 0)

    def hop_context(self):
        return torch.cuda.stream(self.hop_stream)

    def prep_context(self, first=False):
        """Stream context for the next slice's prepare + all_gather.  Nothing in them depends on what the scan stream
        is doing (the library orders the reuse of a scan set behind its last render by itself); only the very first
        call is ordered behind the scan stream, which at that point has waited for the caller's stream (descriptors
        created there right before the first step)."""
        if first:
            self.prep_stream.wait_stream(self.scan_stream)
        return torch.cuda.stream(self.prep_stream)

    def prep_done(self):
        self.prep_event.record(self.prep_stream)

    def wait_prep(self):
        self.scan_stream.wait_event(self.prep_event)

    def scan_context(self, order_after_caller=False):
        # The scan stream must not be ordered after the caller's stream in steady state (that would
        # serialise it behind the previous slice's render); descriptors uploaded on the caller's
        # stream right before a step need order_after_caller=True (the runner does it for step 0).
        if order_after_caller:
            self.scan_stream.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self.scan_stream)

    def apply_bias(self):              # estimate <- estimate - bias, then remember what the slice starts from
        self.s.estimate_correct_device(self.zero.data_ptr(), self.bias.data_ptr(), 1.0, self._stream())
        self.s.estimate_to_device(self.est_start.data_ptr(), self._stream())

    def update_bias(self):             # self.phase = the exact start phase just received for that slice
        diff = self.est_start - self.phase
        diff -= torch.round(diff)
        # only small, systematic differences are integrated (a re-seeded slot or the start of the stream is not one)
        self.bias += torch.where(diff.abs() < 1e-8, diff, torch.zeros_like(diff))

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def load_carrier(self):            # self.phase -> engine state
        self.s.carrier_from_device(self.phase.data_ptr(), self._stream())

    def store_carrier(self):           # engine state -> self.phase
        self.s.carrier_to_device(self.phase.data_ptr(), self._stream())

    def prepare(self, desc_dev, n_epochs):      # -> self.adv
        self.s.prepare_device(desc_dev.data_ptr(), n_epochs, self.adv.data_ptr(), self._stream())

    def estimate_fold(self, adv):
        self.s.estimate_fold_device(adv.data_ptr(), self._stream())

    def carrier_fold(self, adv):       # integer carrier: exact state <- fold(state, advance of a foreign slice)
        self.s.carrier_fold_device(adv.data_ptr(), self._stream())

    def estimate_anchor(self):         # estimate <- exact carrier state
        self.s.estimate_anchor_device(self._stream())

    def estimate_from(self, phases):   # estimate <- max_chan phases in a device tensor (e.g. another rank's exact start)
        self.s.estimate_from_device(phases.data_ptr(), self._stream())

    def gather_payload(self):
        """What a rank contributes to the per-step all_gather of the lockstep runner: its slice's closed-form advance
        and flags (2 * max_chan) + the exact phases it last received (max_chan; rank 0's are the exact start of the
        step, which every other rank anchors its estimate on)."""
        return torch.cat([self.adv, self.phase])

    def speculate(self, desc_dev, n_epochs):
        self.s.speculate_device(desc_dev.data_ptr(), n_epochs, self._stream())

    def chain(self, desc_dev, n_epochs):
        self.s.chain_device(desc_dev.data_ptr(), n_epochs, self._stream())

    def render(self, desc_dev, n_epochs, out):
        """out: a CUDA tensor (asynchronous, on the current stream) or a HOST address (int; pinned memory from
        gpsiq_host_alloc): then the slice goes through the host-buffer call gpsiq_fetch -- rendered in sub-batches
        whose device-to-host copies overlap the rendering, blocking until the host buffer is complete."""
        if isinstance(out, int):
            self.s.fetch_ptr(out)
        else:
            self.s.render_device(desc_dev.data_ptr(), n_epochs, out.data_ptr(), self._stream())

    def upload(self, desc_host, index):
        """Pinned host descriptors (uint8 tensor) -> device staging buffer (three, in turn), on the current stream.
        A staging buffer is reused three slices later; the host-buffer render of the slice that used it has returned
        by then (gpsiq_fetch blocks)."""
        if not hasattr(self, "_stage"):   # three: slices k-1 (render pending), k and -- pipelined runner -- k+1 are alive at once
            self._stage = [torch.empty(self.s.max_epochs * self.s.max_chan * 64, dtype=torch.uint8, device="cuda") for _ in range(3)]
        d = self._stage[index % 3][: desc_host.numel()]
        d.copy_(desc_host, non_blocking=True)
        return d

    # ---- SM-free hand-off (gpsiq_mailbox_*): the sender's copy engine writes the carrier state into the next
    # rank's mailbox over NVLink peer memory and a stream memory operation publishes a sequence number; the
    # receiver's stream waits on it.  No kernel on either GPU, so a hop does not queue for an SM behind the
    # sample kernel the way an NCCL send/recv pair does.
    def mailbox_setup(self, rank, world, device_of_rank=None):
        """Collective (torch.distributed, any backend): every rank exports its mailbox and opens the next rank's.
        -> True if every rank succeeded; False (on all ranks alike) if any could not (no peer access, no CUDA IPC,
        no stream memory operations): the caller then keeps the NCCL hand-off."""
        mine, err = None, None
        try:
            mine = self.s.mailbox_create()
        except Exception as e:                                   # reported below, decided collectively
            err = e
        handles = [None] * world
        dist.all_gather_object(handles, (mine, torch.cuda.current_device()))
        ok = all(h[0] is not None for h in handles)
        if ok:
            nxt = (rank + 1) % world
            handle, dev = handles[nxt]
            try:
                self.s.mailbox_open(handle, dev if device_of_rank is None else device_of_rank(nxt))
            except Exception as e:
                err, ok = e, False
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        torch.cuda.synchronize()
        if not all(oks):
            if err is not None:
                print("mailbox hand-off unavailable on rank %d: %s" % (rank, err), flush=True)
            return False
        return True

    def chain_handoff(self, desc_dev, n_epochs, recv_seq, send_seq):
        """The hop as ONE kernel (gpsiq_chain_handoff_device): wait for message recv_seq (0: none) in the own mailbox,
        chain, end phases as message send_seq (0: none) into the next rank's.  self.phase <- the exact start phases."""
        self.s.chain_handoff_device(desc_dev.data_ptr(), n_epochs, recv_seq, send_seq, self.phase.data_ptr(), self._stream())

    def handoff_send(self, seq):       # engine state -> next rank's mailbox
        self.s.mailbox_send(seq, self._stream())

    def handoff_recv(self, seq):       # own mailbox -> engine state (and self.phase, for the estimate feedback)
        self.s.mailbox_recv(seq, self._stream())
        self.s.carrier_to_device(self.phase.data_ptr(), self._stream())


class TimeSliceRunner:
    def __init__(self, engine, rank=None, world=None, deferred_render=False, handoff="nccl", pipelined=False):
        """handoff: "nccl" (dist.send / dist.recv of the phases) or "mailbox" (the engine's SM-free peer-memory
        hand-off, GpuSliceEngine.mailbox_setup must have run; same node only).
        deferred_render: step k enqueues the scan phases of slice k and then the rendering of slice k-1
        (the last slice is rendered by finish()).  The device then sees the next slice's chunk speculation
        before this slice's sample kernel, which is the order gpsiq_submit/gpsiq_fetch produce on one GPU:
        the speculation runs first at full occupancy and the rest of the chain beside the sample kernel.
        An `out` buffer passed to step k is complete only after step k+1 (or finish) has been enqueued.
        pipelined (float carrier, a hand-off that does not block the host: "mailbox"): the ranks are not kept in
        lockstep.  Rank 0 does not wait for the previous step's ring before it speculates -- like every other rank it
        estimates its start phase from the exact end of its own previous slice and the closed-form advances of the
        N - 1 slices in between, and receives the exact phases only in front of its chain -- and when step() is given
        the NEXT slice's descriptors, that slice is prepared and its advances all-gathered on a side stream beside
        this slice's speculation, so that no rank's speculation waits for the slowest rank of the ring.  The ring of
        step k then only orders [receive, one head scan + translation, send] per rank, and the step time is the
        larger of one rank's own scan work and N hops instead of their sum."""
        self.engine = engine
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.step_index = 0
        self.prev_adv = None
        self.deferred = deferred_render
        self.pending = None
        if handoff not in ("nccl", "mailbox", "prefix"):
            raise ValueError("handoff must be 'nccl', 'mailbox' or 'prefix'")
        self.mailbox = handoff == "mailbox" and self.world > 1
        # "prefix": the integer carrier NCO (GPSIQ_CARRIER_INT32; plutogpssim.c:2748) advances in closed form, so
        # there is no ring at all: one all_gather of the slices' exact advances per step, and every rank folds the
        # advances of the slices before its own into its carrier state (and the ones after it afterwards, so that
        # all ranks hold the phase at the start of the next step).  SURVEY 8e: "truly parallel".
        self.prefix = handoff == "prefix"
        self.pipelined = bool(pipelined) and not self.prefix and self.world > 1
        # free-running (GPSIQ_TS_FREE=1; pipelined runners over the fused mailbox hop; NOT the default): a rank's
        # speculation of slice k+1 does not wait for its hop of slice k.  The estimate it starts from is the end of
        # slice k's own slice-level speculative trajectory + the closed-form advances of the foreign slices in between,
        # corrected open loop by the library; the hop runs on its own stream whenever the previous GPU's phases arrive.
        # Measured at 2 GPUs (profiles/r02_y_*): parity-clean, but 2.25 ms per step against 2.07 -- the corrections
        # (residual rate, accumulated estimate error) arrive two slices later, more slices are speculated from poor
        # estimates at the start of a run, and each of those chains serially ON the ring.
        self.free_running = (self.pipelined and self.mailbox and hasattr(engine, "set_free_running")
                             and os.environ.get("GPSIQ_TS_FREE", "0") == "1"
                             and os.environ.get("GPSIQ_HANDOFF_FUSED", "1") != "0")
        if self.free_running:
            engine.set_free_running(True)
        # experiments: GPSIQ_TS_LOOKAHEAD=0 ignores next_desc, GPSIQ_TS_SIDE=0 keeps the look-ahead on the scan stream
        self.use_lookahead = os.environ.get("GPSIQ_TS_LOOKAHEAD", "1") != "0"
        self.use_side = os.environ.get("GPSIQ_TS_SIDE", "1") != "0"
        # GPSIQ_HANDOFF_FUSED=0: the hop as stream operations (wait value, copies, chain kernel, copy, write value)
        self.fused = self.mailbox and os.environ.get("GPSIQ_HANDOFF_FUSED", "1") != "0"
        self.anchor0 = os.environ.get("GPSIQ_TS_ANCHOR0", "1") != "0"   # lockstep runner: estimates anchored on rank 0's exact start
        self.next_adv = None           # advances of the NEXT step's slices, all-gathered one step ahead (pipelined)
        self.next_desc = None          # ... and that slice's device descriptors (host descriptors are uploaded once)

    def step(self, desc, n_epochs, out, next_desc=None, next_epochs=None):
        """Synthesize this rank's slice of the next step.

        Every rank issues its communication in the same global order -- hand-off into
        rank 0 (closing the previous step's ring), all_gather of this step, then the
        hand-offs 0->1->...->N-1 -- so in-order streams (NCCL) and blocking calls
        (gloo) cannot deadlock.  (pipelined: all_gather of the NEXT step, hand-off into this rank, hand-off out of
        it; the hand-off must not block the host.)

        next_desc / next_epochs (pipelined runners): the descriptors of this rank's slice of the FOLLOWING step, if
        the caller has them -- that step() call must then be given the same slice."""
        eng, r, n = self.engine, self.rank, self.world
        ctx = eng.scan_context(self.step_index == 0) if hasattr(eng, "scan_context") else contextlib.nullcontext()
        with ctx:
            if self.pipelined and self.next_adv is not None:
                desc = self.next_desc                           # prepared (and uploaded) one step ahead
            elif isinstance(desc, torch.Tensor) and desc.device.type == "cpu" and hasattr(eng, "upload"):
                desc = eng.upload(desc, self.step_index)        # host descriptors: H2D on the scan stream
            if self.pipelined:
                self._scan_phases_pipelined(desc, n_epochs, next_desc, next_epochs if next_epochs else n_epochs)
            else:
                self._scan_phases(desc, n_epochs)
        if self.deferred:
            if self.pending is not None:
                eng.render(*self.pending)
            self.pending = (desc, n_epochs, out)
        else:
            eng.render(desc, n_epochs, out)
        self.step_index += 1

    def _scan_phases(self, desc, n_epochs):
        eng, r, n = self.engine, self.rank, self.world
        if self.prefix:
            eng.prepare(desc, n_epochs)
            adv_all = [eng.adv]
            if n > 1:
                adv_all = [torch.empty_like(eng.adv) for _ in range(n)]
                dist.all_gather(adv_all, eng.adv)
            for a in adv_all[:r]:
                eng.carrier_fold(a)
            eng.speculate(desc, n_epochs)                       # (integer carrier: the code-NCO scan only)
            eng.chain(desc, n_epochs)
            for a in adv_all[r + 1:]:
                eng.carrier_fold(a)
            return
        have_exact = False
        if n > 1 and r == 0 and self.step_index > 0:            # close the previous step's ring first
            if self.mailbox:
                eng.handoff_recv(self.step_index)               # message number = sender's step index + 1
            else:
                dist.recv(eng.phase, src=n - 1)
                eng.load_carrier()
            eng.estimate_anchor()                               # rank 0 speculates from the exact phase
            have_exact = True
        eng.prepare(desc, n_epochs)
        if n > 1:
            # Rank 0 has just received the exact phases at the start of this step: it shares them with the advances, and
            # rank r anchors its estimate THERE, r closed-form advances away -- instead of on the exact end of its own
            # previous slice, N - 1 advances away (fewer and fresher terms: fewer slices speculated from a poor
            # estimate, each of which chains serially on the ring).  Step 0 has no such phases: the old way.
            anchor0 = self.anchor0 and hasattr(eng, "gather_payload") and hasattr(eng, "estimate_from")
            payload = eng.gather_payload() if anchor0 else eng.adv
            got = [torch.empty_like(payload) for _ in range(n)]
            dist.all_gather(got, payload)
            na = eng.adv.numel()
            adv_all = [g[:na] for g in got]
            if not have_exact:
                if anchor0 and self.step_index > 0:
                    eng.estimate_from(got[0][na:])              # rank 0's exact start of this step
                    skipped = adv_all[:r]
                else:                                           # slices owned by other ranks since my last one
                    skipped = ([] if self.prev_adv is None else self.prev_adv[r + 1:]) + adv_all[:r]
                for a in skipped:
                    eng.estimate_fold(a)
                if hasattr(eng, "apply_bias"):
                    eng.apply_bias()
            self.prev_adv = adv_all
        eng.speculate(desc, n_epochs)
        if n > 1 and self.mailbox and self.fused and hasattr(eng, "chain_handoff"):   # the hop as one kernel
            eng.chain_handoff(desc, n_epochs, self.step_index + 1 if r > 0 else 0, self.step_index + 1)
            if r > 0 and hasattr(eng, "update_bias"):
                eng.update_bias()
            return
        if n > 1 and r > 0:
            if self.mailbox:
                eng.handoff_recv(self.step_index + 1)
            else:
                dist.recv(eng.phase, src=r - 1)
                eng.load_carrier()
                if hasattr(eng, "update_bias"):
                    eng.update_bias()
        eng.chain(desc, n_epochs)
        if n > 1:
            if self.mailbox:
                eng.handoff_send(self.step_index + 1)
                if r > 0 and hasattr(eng, "update_bias"):       # estimate feedback: off the ring's critical path
                    eng.update_bias()
            else:
                eng.store_carrier()
                dist.send(eng.phase, dst=(r + 1) % n)

    def _prepare_and_gather(self, desc, n_epochs, index):
        """prepare + all_gather of the slices' closed-form advances, into one of four preallocated buffer sets (they
        are written on one stream and read on another: nothing here is left to the caching allocator)."""
        eng, n = self.engine, self.world
        eng.prepare(desc, n_epochs)
        if not hasattr(self, "_adv_ring"):
            self._adv_ring = [[torch.empty_like(eng.adv) for _ in range(n)] for _ in range(4)]
        adv_all = self._adv_ring[index % 4]
        dist.all_gather(adv_all, eng.adv)
        return adv_all

    def _scan_phases_pipelined(self, desc, n_epochs, next_desc, next_epochs):
        eng, r, n, k = self.engine, self.rank, self.world, self.step_index
        if self.next_adv is not None:                           # prepared and gathered during the previous step
            adv_all, self.next_adv, self.next_desc = self.next_adv, None, None
            if hasattr(eng, "wait_prep") and self.use_side:
                eng.wait_prep()
        else:
            adv_all = self._prepare_and_gather(desc, n_epochs, k)
        # start-phase estimate: the exact end of this rank's previous slice (the chain left the estimate there)
        # advanced by the closed-form advances of the N - 1 slices owned by the other ranks in between
        skipped = ([] if self.prev_adv is None else self.prev_adv[r + 1:]) + adv_all[:r]
        for a in skipped:
            eng.estimate_fold(a)
        if hasattr(eng, "apply_bias") and not self.free_running:
            eng.apply_bias()
        self.prev_adv = adv_all
        eng.speculate(desc, n_epochs)
        # the next slice's prepare + all_gather: beside this slice's speculation and hand-off
        if next_desc is not None and self.use_lookahead:
            side = eng.prep_context(k == 0) if (hasattr(eng, "prep_context") and self.use_side) else contextlib.nullcontext()
            with side:
                if isinstance(next_desc, torch.Tensor) and next_desc.device.type == "cpu" and hasattr(eng, "upload"):
                    next_desc = eng.upload(next_desc, k + 1)
                self.next_adv = self._prepare_and_gather(next_desc, next_epochs, k + 1)
                self.next_desc = next_desc
                if hasattr(eng, "prep_done") and self.use_side:
                    eng.prep_done()
        # the ring: exact phases in, one head scan + translation (or the serial chain), exact phases out
        received = not (k == 0 and r == 0)                      # (the stream starts at rank 0 from re-seeded phases)
        recv_seq = (k + 1 if r > 0 else k) if received else 0   # message number = sender's step index + 1
        if self.free_running:                                   # ... on its own stream, behind this slice's speculation
            with eng.hop_context():
                eng.chain_handoff(desc, n_epochs, recv_seq, k + 1)
            return
        if self.fused and hasattr(eng, "chain_handoff"):        # the whole hop is one kernel
            eng.chain_handoff(desc, n_epochs, recv_seq, k + 1)
        else:
            if received:
                eng.handoff_recv(recv_seq)
            eng.chain(desc, n_epochs)
            eng.handoff_send(k + 1)
        if received and hasattr(eng, "update_bias"):            # estimate feedback: off the ring's critical path
            eng.update_bias()

    def finish(self):
        """Render the slice still pending (deferred_render) and drain the last hand-off (rank 0 receives
        the phases after the final slice)."""
        if self.pending is not None:
            self.engine.render(*self.pending)
            self.pending = None
        if self.world > 1 and self.rank == 0 and self.step_index > 0 and not self.prefix:
            eng = self.engine
            ctx = (eng.hop_context() if getattr(self, "free_running", False) else
                   eng.scan_context() if hasattr(eng, "scan_context") else contextlib.nullcontext())
            with ctx:
                if self.mailbox:
                    eng.handoff_recv(self.step_index)
                else:
                    dist.recv(eng.phase, src=self.world - 1)
                    eng.load_carrier()
