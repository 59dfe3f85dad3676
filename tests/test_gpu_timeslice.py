"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): the user-motion golden
stream time-sliced over 2 ranks with the NCCL carrier-phase hand-off must give the
reference's per-epoch checksums, slice by slice."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

pytestmark = pytest.mark.gpu

E = 16
STEPS = 9


CIRCLE = os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv")


def _worker(rank, world, port, outdir, deferred=False, handoff="nccl", feed="golden", golden="circle12", pipelined=False,
            E=E, STEPS=STEPS):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pluto_gps_sim_b200 import Synthesizer
    from pluto_gps_sim_b200.timeslice import GpuSliceEngine, TimeSliceRunner

    desc = ol.load_golden_desc(golden)
    carrier_mode = 1 if handoff == "prefix" else 0
    feeder = None
    if feed == "navfile":   # every rank computes ONLY its own slices' descriptors from the navigation file (gpshost_skip)
        from pluto_gps_sim_b200 import hostapi
        feeder = hostapi.SliceFeeder(rank, world, E, nav=os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz"), motion=CIRCLE,
                                     sample_rate=2600000)
    synth = Synthesizer(max_chan=12, max_epochs=E, device=rank, carrier_mode=carrier_mode)
    engine = GpuSliceEngine(synth)
    if handoff == "mailbox":
        assert engine.mailbox_setup(rank, world)
    runner = TimeSliceRunner(engine, rank, world, deferred_render=deferred, handoff=handoff, pipelined=pipelined)
    outs = [torch.empty(E * 300000 * 2, dtype=torch.int16, device="cuda") for _ in range(2)]
    sums = []
    keep = []                                                  # (descriptor tensors stay alive until their slice is rendered)
    for s in range(STEPS):
        first = (s * world + rank) * E
        mine = feeder.next_slice() if feeder else desc[first:first + E].copy()
        d = torch.from_numpy(mine.view(np.uint8).reshape(-1)).cuda()
        nxt = None
        if pipelined and not feeder and s + 1 < STEPS:         # the next slice's descriptors: prepared one step ahead
            f2 = ((s + 1) * world + rank) * E
            nxt = torch.from_numpy(desc[f2:f2 + E].copy().view(np.uint8).reshape(-1)).cuda()
            keep.append(nxt)
        keep.append(d)
        runner.step(d, E, outs[s & 1], next_desc=nxt)
        torch.cuda.synchronize()
        if not deferred:
            sums.append(synth.checksum_device(outs[s & 1].data_ptr(), E))
        elif s > 0:                              # slice s-1 was rendered by this step
            sums.append(synth.checksum_device(outs[(s - 1) & 1].data_ptr(), E))
    runner.finish()
    torch.cuda.synchronize()
    if deferred:
        sums.append(synth.checksum_device(outs[(STEPS - 1) & 1].data_ptr(), E))
    np.save(os.path.join(outdir, "sums%d.npy" % rank), np.stack(sums))
    synth.check_device()
    np.save(os.path.join(outdir, "fb%d.npy" % rank), np.array([synth.carrier_fallbacks]))
    np.save(os.path.join(outdir, "slice%d.npy" % rank), np.array(synth.slice_stats))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("deferred,handoff", [(False, "nccl"), (True, "nccl"), (False, "mailbox"), (True, "mailbox")])
def test_two_gpu_time_slices_match_reference(tmp_path, deferred, handoff):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), deferred, handoff), nprocs=world, join=True)
    meta = ol.load_golden_meta("circle12")
    parts = [np.load(tmp_path / ("sums%d.npy" % r)) for r in range(world)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(world)])
    assert [int(x) for x in got] == meta["epoch_checksums"][: world * STEPS * E]
    # the estimates handed around the ring are good enough that the serial fallback stays rare
    fb = sum(int(np.load(tmp_path / ("fb%d.npy" % r))[0]) for r in range(world))
    assert fb < world * STEPS * E * 12 // 10, fb


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("slice_epochs,steps,fused", [(16, 9, "1"), (70, 2, "1"), (16, 9, "0")])
def test_two_gpu_pipelined_time_slices_match_reference(tmp_path, monkeypatch, slice_epochs, steps, fused):
    """TimeSliceRunner(pipelined=True) over the mailbox hand-off: no lockstep between the ranks (rank 0 speculates
    from an estimate, the next slice is prepared and its advances all-gathered one step ahead on a side stream), the
    exact chain of a slice is one head scan + a translation (slices of one and of two groups of 64 epochs).  The hop
    is ONE kernel (gpsiq_chain_handoff_device: acquire-wait on the mailbox flag, chain, peer stores + release of the
    next rank's flag) or -- GPSIQ_HANDOFF_FUSED=0 -- stream operations around the chain kernel.  The reference's
    per-epoch checksums, and the slice-level translation is what actually ran."""
    monkeypatch.setenv("GPSIQ_HANDOFF_FUSED", fused)
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), True, "mailbox", "golden", "circle12", True, slice_epochs, steps),
             nprocs=world, join=True)
    meta = ol.load_golden_meta("circle12")
    parts = [np.load(tmp_path / ("sums%d.npy" % r)) for r in range(world)]
    got = np.concatenate([parts[r][s] for s in range(steps) for r in range(world)])
    assert [int(x) for x in got] == meta["epoch_checksums"][: world * steps * slice_epochs]
    translated = sum(int(np.load(tmp_path / ("slice%d.npy" % r))[0]) for r in range(world))
    serial = sum(int(np.load(tmp_path / ("slice%d.npy" % r))[1]) for r in range(world))
    assert translated >= serial, (translated, serial)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("deferred", [False, True])
def test_two_gpu_integer_carrier_slices_without_a_ring(tmp_path, deferred):
    """GPSIQ_CARRIER_INT32: closed-form prefix hand-off (all_gather of the slices' exact advances), against the
    stream of the reference built without FLOAT_CARR_PHASE, across the 30 s re-allocation."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), deferred, "prefix", "golden", "circle12int"), nprocs=world, join=True)
    meta = ol.load_golden_meta("circle12int")
    parts = [np.load(tmp_path / ("sums%d.npy" % r)) for r in range(world)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(world)])
    assert [int(x) for x in got] == meta["epoch_checksums"][: world * STEPS * E]


@pytest.mark.skipif(torch.cuda.device_count() < 2 or not os.path.exists(CIRCLE), reason="needs 2 GPUs and oracle/_ref/circle.csv")
def test_two_gpu_time_slices_fed_from_the_navigation_file(tmp_path):
    """As above, but nothing is precomputed: each rank runs the host orchestrator for its own slices only
    (hostapi.SliceFeeder), renders them, and the stream is the reference's (per-epoch checksums of its 310-epoch run)."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), False, "mailbox", "navfile"), nprocs=world, join=True)
    meta = ol.load_golden_meta("circle12")
    parts = [np.load(tmp_path / ("sums%d.npy" % r)) for r in range(world)]
    got = np.concatenate([parts[r][s] for s in range(STEPS) for r in range(world)])
    assert [int(x) for x in got] == meta["epoch_checksums"][: world * STEPS * E]
