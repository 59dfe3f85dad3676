"""Host logic of the multi-GPU time slicing (pluto_gps_sim_b200/timeslice.py) on
CPU: world_size-2 and -3 gloo processes, the device replaced by an oracle-backed engine;
both hand-off modes (the mailbox one emulated over gloo, sequence numbers checked).
The concatenated slices must equal one sequential run of the whole stream --
i.e. the carrier phases are handed from slice to slice (and from the last rank
back to rank 0 for the next step) exactly like chan[i].carr_phase."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

N = 2500          # samples per epoch (short: the oracle is a plain per-sample loop)
E = 2             # epochs per slice
STEPS = 3


class OracleEngine:
    """Same interface as timeslice.GpuSliceEngine, arithmetic by the oracle."""

    def __init__(self, max_chan, rank=0, world=1, n=N, carrier_mode=0):
        self.rank, self.world, self.n, self.mode = rank, world, n, carrier_mode
        self.phase = torch.zeros(max_chan, dtype=torch.float64)
        self.adv = torch.zeros(2 * max_chan, dtype=torch.float64)
        self.state = np.zeros(max_chan)
        self.pending = None
        self.folds = 0

    def prepare(self, desc, n_epochs):        # closed-form advance of the slice (estimates only)
        d = desc[:n_epochs]
        if self.mode == 1:                     # integer carrier: the EXACT advance (what k_int_carrier mode 1 computes)
            Cn = d.shape[1]
            for c in range(Cn):
                u, seeded = 0, False
                for e in range(n_epochs):
                    if d[e, c]["prn"] <= 0:
                        continue
                    if d[e, c]["flags"] & 1:
                        u, seeded = int(d[e, c]["carr_phase0"]), True
                    u = (u + int(d[e, c]["carr_step"]) * self.n) % 2 ** 32
                self.adv[c], self.adv[Cn + c] = float(u), float(seeded)
            return
        self.adv[: d.shape[1]] = torch.from_numpy(np.mod((d["carr_step"] * self.n).sum(axis=0), 1.0))
        self.adv[d.shape[1]:] = torch.from_numpy(((d["flags"] & 1) != 0).any(axis=0).astype(np.float64))

    def estimate_fold(self, adv):
        self.folds += 1

    def estimate_anchor(self):
        pass

    def carrier_fold(self, adv):
        Cn = self.state.size
        a, f = adv[:Cn].numpy(), adv[Cn:].numpy()
        self.state[:] = np.where(f != 0, a, np.mod(self.state + a, 2.0 ** 32))

    def speculate(self, desc, n_epochs):
        pass

    def load_carrier(self):
        self.state[:] = self.phase.numpy()

    def store_carrier(self):
        self.phase.copy_(torch.from_numpy(self.state.copy()))

    # the mailbox hand-off (gpsiq_mailbox_send / _recv), emulated: the message carries its sequence number,
    # which must be exactly what the receiver waits for (on the GPU a wrong number would hang or read a stale slot)
    def handoff_send(self, seq):
        dist.send(torch.cat([torch.tensor([float(seq)], dtype=torch.float64), torch.from_numpy(self.state.copy())]),
                  dst=(self.rank + 1) % self.world)

    def handoff_recv(self, seq):
        buf = torch.zeros(1 + self.state.size, dtype=torch.float64)
        dist.recv(buf, src=(self.rank - 1) % self.world)
        assert int(buf[0]) == seq, (self.rank, int(buf[0]), seq)
        self.state[:] = buf[1:].numpy()

    def chain(self, desc, n_epochs):          # advances the carrier state (and keeps the samples for render)
        self.pending, _ = ol.oracle_synth(desc[:n_epochs], self.n, carrier_mode=self.mode, carr_state=self.state)

    def render(self, desc, n_epochs, out):
        out[...] = self.pending


def _stream_desc(world, name="static12"):
    d = ol.load_golden_desc(name)
    d = np.concatenate([d, d])[: world * STEPS * E].copy()
    d["flags"] = 0
    d[0]["flags"] = 1
    return d


def _worker(rank, WORLD, port, outdir, handoff):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from pluto_gps_sim_b200.timeslice import TimeSliceRunner

    desc = _stream_desc(WORLD)
    eng = OracleEngine(desc.shape[1], rank, WORLD)
    runner = TimeSliceRunner(eng, rank, WORLD, handoff=handoff)
    outs = []
    for s in range(STEPS):
        first = (s * WORLD + rank) * E
        out = np.zeros((E, N, 2), np.int16)
        runner.step(desc[first:first + E], E, out)
        outs.append(out)
    runner.finish()
    # rank 0 re-anchors on the exact phase instead of folding; other ranks fold every foreign slice once
    assert eng.folds == (0 if rank == 0 else STEPS * (WORLD - 1) - (WORLD - 1 - rank))
    np.save(os.path.join(outdir, "rank%d.npy" % rank), np.stack(outs))
    if rank == 0:
        np.save(os.path.join(outdir, "final_phase.npy"), eng.state)
    dist.destroy_process_group()


@pytest.mark.parametrize("WORLD,handoff", [(2, "nccl"), (2, "mailbox"), (3, "nccl"), (3, "mailbox")])
def test_time_slices_equal_sequential_stream(tmp_path, WORLD, handoff):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path), handoff), nprocs=WORLD, join=True)
    desc = _stream_desc(WORLD)
    st = np.zeros(desc.shape[1])
    want, _ = ol.oracle_synth(desc, N, carr_state=st)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(WORLD)])
    assert np.array_equal(got, want)
    # rank 0 ends up holding the phases after the very last slice (ready for the next step)
    assert np.array_equal(np.load(tmp_path / "final_phase.npy"), st)


# ---- lockstep runner, estimates anchored on rank 0's exact start ----------------------------------------------------
class AnchorOracleEngine(OracleEngine):
    """Adds what GpuSliceEngine offers for the rank-0 anchor: the gather payload (advance, flags, last exact phases
    received) and estimate_from; keeps a shadow 'estimate' so that the test can check what it would be anchored on."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.anchors = []

    def gather_payload(self):
        return torch.cat([self.adv, self.phase])

    def estimate_from(self, phases):
        assert phases.numel() == self.state.size and phases.is_contiguous()
        self.anchors.append(phases.clone())

    def handoff_recv(self, seq):
        super().handoff_recv(seq)
        self.phase.copy_(torch.from_numpy(self.state.copy()))   # (GpuSliceEngine: carrier_to_device(self.phase))


def _anchor_worker(rank, WORLD, port, outdir, handoff):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from pluto_gps_sim_b200.timeslice import TimeSliceRunner

    desc = _stream_desc(WORLD)
    eng = AnchorOracleEngine(desc.shape[1], rank, WORLD)
    runner = TimeSliceRunner(eng, rank, WORLD, handoff=handoff)
    outs, starts = [], []
    for s in range(STEPS):
        first = (s * WORLD + rank) * E
        out = np.zeros((E, N, 2), np.int16)
        runner.step(desc[first:first + E], E, out)
        outs.append(out)
    runner.finish()
    # rank r > 0: step 0 folds the r slices before it (no exact start yet); every later step anchors on rank 0's exact
    # start and folds r advances -- not the N - 1 of the old scheme
    assert eng.folds == (0 if rank == 0 else rank * STEPS), (rank, eng.folds)
    assert len(eng.anchors) == (0 if rank == 0 else STEPS - 1)
    np.save(os.path.join(outdir, "rank%d.npy" % rank), np.stack(outs))
    np.save(os.path.join(outdir, "anchors%d.npy" % rank), np.stack([a.numpy() for a in eng.anchors]) if eng.anchors else np.zeros((0, desc.shape[1])))
    dist.destroy_process_group()


@pytest.mark.parametrize("WORLD,handoff", [(2, "mailbox"), (3, "mailbox"), (3, "nccl")])
def test_lockstep_estimates_anchor_on_rank0_exact_start(tmp_path, WORLD, handoff):
    """Lockstep runner: rank 0 shares the exact phases it received at the start of the step in the per-step all_gather;
    the other ranks anchor their start-phase estimates there.  The stream is unchanged (estimates never touch a
    sample), and what the ranks anchor on IS the exact phase at the start of each step."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_anchor_worker, args=(WORLD, port, str(tmp_path), handoff), nprocs=WORLD, join=True)
    desc = _stream_desc(WORLD)
    st = np.zeros(desc.shape[1])
    want, _ = ol.oracle_synth(desc, N, carr_state=st)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(WORLD)])
    assert np.array_equal(got, want)
    # the exact phase at the start of step s = the sequential run's state after s * WORLD * E epochs
    st2 = np.zeros(desc.shape[1])
    for s in range(1, STEPS):
        ol.oracle_synth(desc[(s - 1) * WORLD * E: s * WORLD * E], N, carr_state=st2)
        for r in range(1, WORLD):
            assert np.array_equal(np.load(tmp_path / ("anchors%d.npy" % r))[s - 1], st2), (s, r)


# ---- pipelined runner: no lockstep between the ranks ---------------------------------------------------------------
class AsyncOracleEngine(OracleEngine):
    """The mailbox hand-off never blocks the sender (a copy engine write + a flag): isend here."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.inflight = []
        self.prepared = []

    def prepare(self, desc, n_epochs):
        super().prepare(desc, n_epochs)
        self.prepared.append(desc[:n_epochs].copy())

    def handoff_send(self, seq):
        buf = torch.cat([torch.tensor([float(seq)], dtype=torch.float64), torch.from_numpy(self.state.copy())])
        self.inflight.append((dist.isend(buf, dst=(self.rank + 1) % self.world), buf))

    def drain(self):
        for w, _ in self.inflight:
            w.wait()

    # deferred rendering: slices are rendered one step after they were chained, in order
    def chain(self, desc, n_epochs):
        super().chain(desc, n_epochs)
        self.queue = getattr(self, "queue", []) + [self.pending]

    def render(self, desc, n_epochs, out):
        out[...] = self.queue.pop(0)


def _pipelined_worker(rank, WORLD, port, outdir, lookahead):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from pluto_gps_sim_b200.timeslice import TimeSliceRunner

    desc = _stream_desc(WORLD)
    eng = AsyncOracleEngine(desc.shape[1], rank, WORLD)
    runner = TimeSliceRunner(eng, rank, WORLD, handoff="mailbox", pipelined=True, deferred_render=True)
    outs = []
    for s in range(STEPS):
        first = (s * WORLD + rank) * E
        nxt = ((s + 1) * WORLD + rank) * E
        out = np.zeros((E, N, 2), np.int16)
        give_next = lookahead == "always" or (lookahead == "sometimes" and s == 0)
        runner.step(desc[first:first + E], E, out, next_desc=desc[nxt:nxt + E] if (give_next and s + 1 < STEPS) else None)
        outs.append(out)
    runner.finish()
    eng.drain()
    # every rank -- rank 0 included -- estimates its start phase by folding the N - 1 foreign slices in between
    assert eng.folds == rank + (STEPS - 1) * (WORLD - 1), (rank, eng.folds)
    # every slice was prepared exactly once, in stream order, whether ahead of its step or inside it
    assert len(eng.prepared) == STEPS
    for s in range(STEPS):
        first = (s * WORLD + rank) * E
        assert np.array_equal(eng.prepared[s], desc[first:first + E])
    np.save(os.path.join(outdir, "rank%d.npy" % rank), np.stack(outs))
    if rank == 0:
        np.save(os.path.join(outdir, "final_phase.npy"), eng.state)
    dist.destroy_process_group()


@pytest.mark.parametrize("WORLD,lookahead", [(2, "always"), (3, "always"), (3, "sometimes"), (2, "never")])
def test_pipelined_time_slices_equal_sequential_stream(tmp_path, WORLD, lookahead):
    """TimeSliceRunner(pipelined=True): rank 0 speculates from an estimate like everyone else and receives the exact
    phases only in front of its chain; the next slice is prepared and its advances all-gathered one step ahead (when
    the caller hands its descriptors over); rendering is deferred by one step.  Same stream as one sequential run."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_pipelined_worker, args=(WORLD, port, str(tmp_path), lookahead), nprocs=WORLD, join=True)
    desc = _stream_desc(WORLD)
    st = np.zeros(desc.shape[1])
    want, _ = ol.oracle_synth(desc, N, carr_state=st)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(WORLD)])
    assert np.array_equal(got, want)
    assert np.array_equal(np.load(tmp_path / "final_phase.npy"), st)


# ---- integer carrier: closed-form prefix instead of a ring ------------------------------------------------------
def _prefix_worker(rank, WORLD, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from pluto_gps_sim_b200.timeslice import TimeSliceRunner

    desc = _stream_desc(WORLD, "circle12int")
    desc[WORLD * E + 1]["flags"][3] |= 1      # a slot re-seeded in the middle of a slice: its advance is absolute
    eng = OracleEngine(desc.shape[1], rank, WORLD, carrier_mode=1)
    runner = TimeSliceRunner(eng, rank, WORLD, handoff="prefix")
    outs = []
    for s in range(STEPS):
        first = (s * WORLD + rank) * E
        out = np.zeros((E, N, 2), np.int16)
        runner.step(desc[first:first + E], E, out)
        outs.append(out)
    runner.finish()
    np.save(os.path.join(outdir, "rank%d.npy" % rank), np.stack(outs))
    np.save(os.path.join(outdir, "final_phase%d.npy" % rank), eng.state)
    dist.destroy_process_group()


@pytest.mark.parametrize("WORLD", [2, 3])
def test_integer_carrier_slices_need_no_ring(tmp_path, WORLD):
    """GPSIQ_CARRIER_INT32 (plutogpssim.c:2748): the per-slice advance is exact modulo 2^32, so the hand-off is an
    all_gather + local prefix; the slices still concatenate to the sequential stream, and EVERY rank ends up with
    the phases after the last slice."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_prefix_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    desc = _stream_desc(WORLD, "circle12int")
    desc[WORLD * E + 1]["flags"][3] |= 1
    st = np.zeros(desc.shape[1])
    want, _ = ol.oracle_synth(desc, N, carrier_mode=1, carr_state=st)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(WORLD)])
    assert np.array_equal(got, want)
    for r in range(WORLD):
        assert np.array_equal(np.load(tmp_path / ("final_phase%d.npy" % r)), st)


# ---- the whole host side, per rank, from the navigation file ---------------------------------------------------
def _nav_worker(rank, WORLD, port, outdir, slice_epochs, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from pluto_gps_sim_b200 import hostapi
    from pluto_gps_sim_b200.timeslice import TimeSliceRunner

    nav = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
    feeder = hostapi.SliceFeeder(rank, WORLD, slice_epochs, nav=nav, llh=(30.286502, 120.032669, 100), sample_rate=2600000)
    eng = OracleEngine(12, rank, WORLD, n=300000)
    runner = TimeSliceRunner(eng, rank, WORLD, handoff="mailbox")
    outs = []
    for _ in range(steps):
        out = np.zeros((slice_epochs, 300000, 2), np.int16)
        runner.step(feeder.next_slice(), slice_epochs, out)      # this rank's own epochs only; the others' are skipped
        outs.append(out)
    runner.finish()
    np.save(os.path.join(outdir, "rank%d.npy" % rank), np.stack(outs))
    dist.destroy_process_group()


@pytest.mark.parametrize("WORLD,slice_epochs,steps", [(2, 5, 1), (2, 1, 5), (5, 2, 1)])
def test_ranks_fed_from_the_navigation_file_reproduce_the_reference_stream(tmp_path, WORLD, slice_epochs, steps):
    """Navigation file in, every rank computing ONLY its own slices' descriptors (hostapi.SliceFeeder), carrier
    phases handed from slice to slice: the reassembled 10 epochs are the reference's own bytes (golden SHA-256 of
    BASELINE config[0]).  The device is the oracle here; tests/test_gpu_timeslice.py runs the same runner on GPUs."""
    import hashlib
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nav_worker, args=(WORLD, port, str(tmp_path), slice_epochs, steps), nprocs=WORLD, join=True)
    parts = [np.load(tmp_path / ("rank%d.npy" % r)) for r in range(WORLD)]
    got = np.concatenate([parts[r][s] for s in range(steps) for r in range(WORLD)])
    assert got.shape[0] == 10
    assert hashlib.sha256(got.tobytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"]
