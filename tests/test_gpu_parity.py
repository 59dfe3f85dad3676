"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called
through the C-ABI, against the oracle and the committed reference goldens."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from pluto_gps_sim_b200 import Synthesizer, capi, checksum_host

pytestmark = pytest.mark.gpu

KERNELS = [capi.KERNEL_LANE_PER_CHANNEL, capi.KERNEL_LINE]
FAST_KERNELS = [capi.KERNEL_LINE]


def first_diff(a, b):
    bad = np.argwhere(a != b)
    return None if len(bad) == 0 else tuple(bad[0])


@pytest.mark.parametrize("kernel", KERNELS)
def test_config1_static12_bit_exact_vs_reference_golden(kernel):
    """BASELINE config[1] == config[0]: 1 s, 12 channels, 2.6 MS/s."""
    meta = ol.load_golden_meta("static12")
    desc = ol.load_golden_desc("static12")
    with Synthesizer(max_chan=12, max_epochs=10, kernel=kernel) as s:
        iq = s.synth(desc)
        trace = s.carrier_trace(10)
        assert s.launch_count >= 4
    assert ol.sha256(iq) == meta["iq_sha256"], first_diff(iq, ol.oracle_synth(desc, 300000)[0])
    want = np.array([[float.fromhex(h) for h in row] for row in meta["carr_phase_end_hex"]])
    assert np.array_equal(trace, want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_config3_allsky32_bit_exact_vs_reference_golden(kernel):
    meta = ol.load_golden_meta("allsky32")
    desc = ol.load_golden_desc("allsky32")
    with Synthesizer(max_chan=32, max_epochs=20, kernel=kernel) as s:
        iq = s.synth(desc)
    assert ol.sha256(iq) == meta["iq_sha256"], first_diff(iq, ol.oracle_synth(desc, 300000)[0])


@pytest.mark.parametrize("kernel", KERNELS)
def test_config2_circle_310_epochs_batched_with_carry(kernel):
    """User motion, 310 epochs across the 30 s refresh, in uneven batches: the
    carrier phase must carry between calls like chan[i].carr_phase does."""
    meta = ol.load_golden_meta("circle12")
    desc = ol.load_golden_desc("circle12")
    sums = []
    with Synthesizer(max_chan=12, max_epochs=128, kernel=kernel) as s:
        e = 0
        for n in (1, 7, 128, 100, 74):
            s.synth(desc[e:e + n], keep_on_device=True)
            sums += [int(x) for x in s.checksum_device(s.device_iq_ptr(), n)]
            e += n
    assert e == 310
    assert sums == meta["epoch_checksums"]


@pytest.mark.parametrize("tile", [32, 96, 1024, 4096, 300000])
def test_tile_size_independence(tile):
    desc = ol.load_golden_desc("static12")[:2]
    want, _ = ol.oracle_synth(desc, 50000)
    with Synthesizer(max_chan=12, samples_per_epoch=50000, max_epochs=2, tile_samples=tile,
                     kernel=capi.KERNEL_LANE_PER_CHANNEL) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None


@pytest.mark.parametrize("n", [1, 31, 33, 1000, 4097])
def test_ragged_epoch_lengths(n):
    desc = ol.load_golden_desc("allsky32")[:3]
    want, _ = ol.oracle_synth(desc, n)
    with Synthesizer(max_chan=32, samples_per_epoch=n, max_epochs=3) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None


@pytest.mark.parametrize("kernel", KERNELS)
def test_inactive_slots_and_empty_epoch(kernel):
    desc = ol.load_golden_desc("static12")[:2].copy()
    desc[0, 3]["prn"] = 0
    desc[1, :]["prn"] = 0          # an epoch with no satellites: all-zero I/Q
    want, _ = ol.oracle_synth(desc, 20000)
    with Synthesizer(max_chan=12, samples_per_epoch=20000, max_epochs=2, kernel=kernel) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None
    assert not got[1].any()


@pytest.mark.parametrize("kernel", FAST_KERNELS)
@pytest.mark.parametrize("n", [8, 1020, 1028, 4100, 30004])
def test_fixed_point_kernel_ragged_epochs(n, kernel):
    desc = ol.load_golden_desc("circle12")[300:303]
    want, _ = ol.oracle_synth(desc, n)
    with Synthesizer(max_chan=12, samples_per_epoch=n, max_epochs=3, kernel=kernel) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None


@pytest.mark.parametrize("kernel", FAST_KERNELS)
@pytest.mark.parametrize("nchan", [13, 17, 24, 32])
def test_fixed_point_kernel_channel_groups(nchan, kernel):
    """More than 16 slots: the kernel walks the channels in groups of 16 resident tables."""
    desc = ol.load_golden_desc("allsky32")[:3, :nchan].copy()
    want, _ = ol.oracle_synth(desc, 20000)
    with Synthesizer(max_chan=nchan, samples_per_epoch=20000, max_epochs=3, kernel=kernel) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None


@pytest.mark.parametrize("kernel", FAST_KERNELS)
def test_fixed_point_kernel_routes_out_of_contract_epochs_to_lane_kernel(kernel):
    """Huge gains (packed int16 accumulation could overflow) and extreme Doppler
    (segment lists too long) are rendered by the lane-per-channel kernel; the
    reference's (short) wrap-around is reproduced either way."""
    desc = ol.load_golden_desc("static12")[:4].copy()
    desc[1]["gain"] *= 40.0                    # |sum| far beyond int16: the reference wraps, so must we
    desc[2, 5]["carr_step"] = 0.0123           # ~32 kHz Doppler at 2.6 MS/s
    desc[3, 7]["carr_step"] = -0.0077
    want, wt = ol.oracle_synth(desc, 40000)
    with Synthesizer(max_chan=12, samples_per_epoch=40000, max_epochs=4, kernel=kernel) as s:
        got = s.synth(desc)
        gt = s.carrier_trace(4)
    assert first_diff(got, want) is None
    assert np.array_equal(gt, wt)


@pytest.mark.parametrize("kernel", FAST_KERNELS)
def test_fixed_point_kernel_random_descriptors(kernel):
    """Synthetic descriptors over the whole contract range (both Doppler signs up to
    ~9 kHz, code phases near the wrap, NAV edges in the first tile)."""
    rng = np.random.default_rng(7)
    E, Cn, n = 6, 12, 66000
    desc = np.zeros((E, Cn), capi.DESC_DTYPE)
    for e in range(E):
        for c in range(Cn):
            f = rng.uniform(-9000, 9000) if c % 4 else rng.uniform(-40, 40)
            d = desc[e, c]
            d["prn"] = int(rng.integers(1, 33))
            d["ms0"] = int(rng.integers(0, 1000))
            d["navbits"] = int(rng.integers(0, 2 ** 62))
            d["code_phase0"] = [rng.uniform(0, 1023), 1022.9999, 0.0, 1e-7][int(rng.integers(0, 4))]
            d["code_step"] = (1.023e6 + f / 1540.0) / 2.6e6
            d["carr_step"] = f / 2.6e6
            d["carr_phase0"] = rng.random()
            d["gain"] = rng.uniform(0.05, 1.5)
            d["flags"] = 1 if (e == 0 or rng.random() < 0.1) else 0
    want, wt = ol.oracle_synth(desc, n)
    with Synthesizer(max_chan=Cn, samples_per_epoch=n, max_epochs=E, kernel=kernel) as s:
        got = s.synth(desc)
        gt = s.carrier_trace(E)
    assert first_diff(got, want) is None
    assert np.array_equal(gt, wt)


@pytest.mark.parametrize("dbg", [capi.LINE_DBG_FORCE_CHUNK, capi.LINE_DBG_FORCE_TILE,
                                 capi.LINE_DBG_FORCE_TILE | capi.LINE_DBG_PERTURB])
def test_line_kernel_check_and_patch_paths(dbg):
    """The line kernel's second-level check, its literal-recurrence patch path, and -- with the anchors the
    main kernel uses deliberately shifted -- the patch application: all must still give the reference's bytes."""
    desc = ol.load_golden_desc("circle12")[298:302]       # crosses the 30 s refresh (re-allocation, NAV rebuild)
    n = 70000
    want, wt = ol.oracle_synth(desc, n)
    with Synthesizer(max_chan=12, samples_per_epoch=n, max_epochs=4, kernel=capi.KERNEL_LINE, line_debug=dbg) as s:
        got = s.synth(desc)
        gt = s.carrier_trace(4)
        hz, patches, chunks = s.line_stats
    assert first_diff(got, want) is None
    assert np.array_equal(gt, wt)
    n_active = int((desc["prn"] > 0).sum())
    if dbg & capi.LINE_DBG_FORCE_TILE:
        assert hz == n_active * 69 and chunks == 0         # every tile listed directly, no refinement pass
    else:
        assert chunks == n_active                          # every active (epoch, slot) went through the refinement
    if dbg & capi.LINE_DBG_PERTURB:
        assert patches > 1000                              # the shifted anchors really were wrong
    else:
        assert patches == 0                                # the line alone already matches the recurrence


def test_line_kernel_clears_almost_every_tile():
    """On the reference scenario the check must clear (nearly) all tiles: the patch path is a safety net."""
    desc = ol.load_golden_desc("static12")
    with Synthesizer(max_chan=12, max_epochs=10, kernel=capi.KERNEL_LINE) as s:
        s.synth(desc, keep_on_device=True)
        hz, patches, chunks = s.line_stats
    assert chunks <= 12 and hz <= 8 and patches == 0, (hz, patches, chunks)


def test_line_kernel_nav_edges_in_every_position():
    """NAV bit edges and code-period wraps at every offset inside a tile (the 4-variant chip tables)."""
    rng = np.random.default_rng(21)
    E, Cn, n = 3, 12, 9000
    desc = np.zeros((E, Cn), capi.DESC_DTYPE)
    for e in range(E):
        for c in range(Cn):
            d = desc[e, c]
            f = rng.uniform(-5000, 5000)
            d["prn"] = int(rng.integers(1, 33))
            d["ms0"] = 19 + 20 * int(rng.integers(0, 40))            # icode = 19: the first wrap is a bit edge
            d["navbits"] = int(rng.integers(0, 2 ** 63))
            d["code_step"] = (1.023e6 + f / 1540.0) / 2.6e6
            d["code_phase0"] = 1023.0 - d["code_step"] * (1 + 97 * c + 13 * e) - 1e-9   # wrap after 1 + 97c + 13e samples
            d["carr_step"] = f / 2.6e6
            d["carr_phase0"] = rng.random()
            d["gain"] = rng.uniform(0.2, 1.0)
            d["flags"] = 1
    want, _ = ol.oracle_synth(desc, n)
    with Synthesizer(max_chan=Cn, samples_per_epoch=n, max_epochs=E, kernel=capi.KERNEL_LINE) as s:
        got = s.synth(desc)
    assert first_diff(got, want) is None


def test_integer_carrier_mode_vs_oracle():
    """The reference's compiled-out uint32 carrier (plutogpssim.c:2699, 2748)."""
    dump = np.load(os.path.join(ol.GOLDEN, "static12_dump.npy"))[:3]
    desc = np.zeros(dump.shape, capi.DESC_DTYPE)
    for e in range(dump.shape[0]):
        for c in range(dump.shape[1]):
            r = dump[e, c]
            ph = float(np.uint32(512.0 * 65536.0 * r["carr_phase"]))     # plutogpssim.c:1967
            desc[e, c] = capi.make_desc(capi.CARRIER_INT32, r["prn"], r["f_carr"], r["f_code"], r["delt"], ph,
                                        r["code_phase"], r["dwrd"].astype(np.uint64), r["iword"], r["ibit"],
                                        r["icode"], r["gain"], e == 0)
    want, wt = ol.oracle_synth(desc, 100000, carrier_mode=1)
    with Synthesizer(max_chan=12, samples_per_epoch=100000, max_epochs=3, carrier_mode=capi.CARRIER_INT32) as s:
        got = s.synth(desc)
        gt = s.carrier_trace(3)
    assert first_diff(got, want) is None
    assert np.array_equal(gt, wt)


@pytest.mark.parametrize("kernel", KERNELS)
def test_integer_carrier_static12_bit_exact_vs_reference_golden(kernel):
    """Integer carrier NCO against the reference compiled without FLOAT_CARR_PHASE (tests/golden/static12int_*)."""
    meta = ol.load_golden_meta("static12int")
    desc = ol.load_golden_desc("static12int")
    with Synthesizer(max_chan=12, max_epochs=10, carrier_mode=capi.CARRIER_INT32, kernel=kernel) as s:
        iq = s.synth(desc)
        trace = s.carrier_trace(10)
    assert ol.sha256(iq) == meta["iq_sha256"], first_diff(iq, ol.oracle_synth(desc, 300000, carrier_mode=1)[0])
    want = np.array([[float.fromhex(h) for h in row] for row in meta["carr_phase_end_hex"]])
    assert np.array_equal(trace, want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_integer_carrier_circle_310_epochs_batched_with_carry(kernel):
    """Integer carrier across the 30 s refresh (slots re-seeded), uneven batches, submit/fetch carry."""
    meta = ol.load_golden_meta("circle12int")
    desc = ol.load_golden_desc("circle12int")
    sums = []
    with Synthesizer(max_chan=12, max_epochs=128, carrier_mode=capi.CARRIER_INT32, kernel=kernel) as s:
        e = 0
        for n in (1, 7, 128, 100, 74):
            s.synth(desc[e:e + n], keep_on_device=True)
            sums += [int(x) for x in s.checksum_device(s.device_iq_ptr(), n)]
            e += n
    assert sums == meta["epoch_checksums"]


def test_parallel_and_serial_carrier_scan_agree():
    """The speculate/translate/verify carrier scan vs the plain serial chain on the
    device: same stream, same carrier trace; and the fast path is really taken."""
    meta = ol.load_golden_meta("circle12")
    desc = ol.load_golden_desc("circle12")[:64]
    with Synthesizer(max_chan=12, max_epochs=64) as a, Synthesizer(max_chan=12, max_epochs=64, serial_carrier_scan=True) as b:
        a.synth(desc, keep_on_device=True)
        b.synth(desc, keep_on_device=True)
        sa = a.checksum_device(a.device_iq_ptr(), 64)
        sb = b.checksum_device(b.device_iq_ptr(), 64)
        assert np.array_equal(a.carrier_trace(64), b.carrier_trace(64))
        n_active = int((desc["prn"] > 0).sum())
        assert a.carrier_fallbacks < n_active // 4, a.carrier_fallbacks
        assert b.carrier_fallbacks == 0
    assert [int(x) for x in sa] == meta["epoch_checksums"][:64]
    assert np.array_equal(sa, sb)


def test_set_get_carrier_handoff():
    """Time-slice hand-off: synthesizing epochs [5,10) on a fresh context seeded
    with the carrier phases after epoch 4 equals the second half of one run."""
    meta = ol.load_golden_meta("static12")
    desc = ol.load_golden_desc("static12")
    with Synthesizer(max_chan=12, max_epochs=10) as a:
        a.synth(desc[:5], keep_on_device=True)
        ph = a.carrier
    with Synthesizer(max_chan=12, max_epochs=10) as b:
        b.carrier = ph
        iq = b.synth(desc[5:])
    assert [int(checksum_host(iq[e])) for e in range(5)] == meta["epoch_checksums"][5:]


def test_out_of_contract_descriptor_is_rejected():
    desc = ol.load_golden_desc("static12")[:1].copy()
    desc[0, 2]["code_phase0"] = 2000.0
    with Synthesizer(max_chan=12, samples_per_epoch=1000, max_epochs=1) as s:
        with pytest.raises(capi.GpsiqError) as ei:
            s.synth(desc)
        assert ei.value.status == capi.ERR_ARG
    with Synthesizer(max_chan=12, samples_per_epoch=1000, max_epochs=1) as s:
        with pytest.raises(capi.GpsiqError) as ei:
            s.synth(ol.load_golden_desc("static12")[:2])
        assert ei.value.status == capi.ERR_CAPACITY


def test_epoch_longer_than_the_nav_window_is_refused():
    """The descriptor carries 64 NAV bits (1.28 s): an epoch that would reach past them is out of contract and must be
    refused (GPSIQ_ERR_ARG), not rendered with a wrapped-around window."""
    desc = ol.load_golden_desc("static12")[:1].copy()
    with Synthesizer(max_chan=12, samples_per_epoch=3400000, max_epochs=1) as s:      # 1.31 s of signal at 2.6 MS/s
        with pytest.raises(capi.GpsiqError) as ei:
            s.synth(desc)
        assert ei.value.status == capi.ERR_ARG
    with Synthesizer(max_chan=12, samples_per_epoch=3200000, max_epochs=1) as s:      # 1.23 s: inside the window
        got = s.synth(desc)
    want, _ = ol.oracle_synth(desc, 3200000)
    assert first_diff(got, want) is None


def test_two_batches_of_lookahead_and_the_multi_device_entry_points():
    """Deeper pipelining (three scan sets: two batches scanned ahead of the one being rendered, scans of consecutive
    batches overlapping on their own streams), and the same stream through gpsiq_multi_* on one device."""
    import ctypes as C
    import torch

    meta = ol.load_golden_meta("circle12")
    desc = ol.load_golden_desc("circle12")
    sizes = [40, 64, 64, 1, 64, 64, 13]
    st = torch.cuda.Stream()
    sums = []
    with torch.cuda.stream(st), Synthesizer(max_chan=12, max_epochs=64) as s:
        bufs = [torch.empty(64 * 300000 * 2, dtype=torch.int16, device="cuda") for _ in range(2)]
        descs, e = [], 0
        for n in sizes:
            descs.append(torch.from_numpy(desc[e:e + n].copy().view(np.uint8).reshape(-1)).cuda())
            e += n
        torch.cuda.synchronize()
        sub = 0
        for k in range(len(sizes)):
            while sub < len(sizes) and sub <= k + 2:
                s.submit_device(descs[sub].data_ptr(), sizes[sub], st.cuda_stream)
                sub += 1
            s.fetch_device(bufs[k & 1].data_ptr(), st.cuda_stream)
            st.synchronize()
            sums += [int(x) for x in s.checksum_device(bufs[k & 1].data_ptr(), sizes[k])]
    assert sums == meta["epoch_checksums"]

    # gpsiq_multi_* with one device: host descriptors in, host I/Q out, begin/end pairs
    cfg = capi.Config()
    cfg.device, cfg.max_chan, cfg.samples_per_epoch, cfg.carrier_mode, cfg.max_epochs = 0, 12, 300000, 0, 64
    m = C.c_void_p()
    capi.check(capi.lib.gpsiq_multi_create(C.byref(m), C.byref(cfg), 1))
    try:
        outs = [np.empty((n, 300000, 2), np.int16) for n in sizes]
        e = 0
        parts = []
        for n in sizes:
            parts.append(np.ascontiguousarray(desc[e:e + n]))
            e += n
        sub = 0
        got = []
        for k in range(len(sizes)):
            while sub < len(sizes) and sub <= k + 2:
                assert capi.lib.gpsiq_multi_submit(m, parts[sub].ctypes.data, sizes[sub]) == 0, capi.lib.gpsiq_multi_last_error(m)
                sub += 1
            assert capi.lib.gpsiq_multi_fetch_begin(m, outs[k].ctypes.data) == 0, capi.lib.gpsiq_multi_last_error(m)
            assert capi.lib.gpsiq_multi_fetch_end(m) == 0, capi.lib.gpsiq_multi_last_error(m)
            got += [int(checksum_host(outs[k][i])) for i in range(sizes[k])]
        assert got == meta["epoch_checksums"]
    finally:
        capi.lib.gpsiq_multi_destroy(m)


def test_streaming_submit_fetch_matches_reference():
    """submit/fetch with one batch of lookahead over the 310-epoch user-motion golden (uneven batches)."""
    import torch

    meta = ol.load_golden_meta("circle12")
    desc = ol.load_golden_desc("circle12")
    sizes = [40, 64, 64, 1, 64, 64, 13]
    assert sum(sizes) == 310
    st = torch.cuda.Stream()
    sums = []
    with torch.cuda.stream(st), Synthesizer(max_chan=12, max_epochs=64) as s:
        bufs = [torch.empty(64 * 300000 * 2, dtype=torch.int16, device="cuda") for _ in range(2)]
        descs, e = [], 0
        for n in sizes:
            descs.append(torch.from_numpy(desc[e:e + n].copy().view(np.uint8).reshape(-1)).cuda())
            e += n
        torch.cuda.synchronize()
        s.submit_device(descs[0].data_ptr(), sizes[0], st.cuda_stream)
        for k in range(len(sizes)):
            if k + 1 < len(sizes):
                s.submit_device(descs[k + 1].data_ptr(), sizes[k + 1], st.cuda_stream)
            s.fetch_device(bufs[k & 1].data_ptr(), st.cuda_stream)
            st.synchronize()
            sums += [int(x) for x in s.checksum_device(bufs[k & 1].data_ptr(), sizes[k])]
    assert sums == meta["epoch_checksums"]


def test_host_streaming_submit_fetch_matches_reference():
    """gpsiq_submit / gpsiq_fetch with host buffers, one batch of lookahead, static golden in 3+3+4 epochs."""
    meta = ol.load_golden_meta("static12")
    desc = ol.load_golden_desc("static12")
    parts = [desc[0:3], desc[3:6], desc[6:10]]
    outs = []
    with Synthesizer(max_chan=12, max_epochs=4) as s:
        s.submit(parts[0])
        for k in range(3):
            if k + 1 < 3:
                s.submit(parts[k + 1])
            outs.append(s.fetch(len(parts[k])))
    iq = np.concatenate(outs)
    assert ol.sha256(iq) == meta["iq_sha256"]


def test_device_path_with_torch_buffers():
    import torch

    meta = ol.load_golden_meta("static12")
    desc = ol.load_golden_desc("static12")
    d = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).cuda()
    out = torch.empty(10 * 300000 * 2, dtype=torch.int16, device="cuda")
    with Synthesizer(max_chan=12, max_epochs=10) as s:
        st = torch.cuda.current_stream().cuda_stream
        s.synth_device(d.data_ptr(), 10, out.data_ptr(), st)
        torch.cuda.synchronize()
    assert ol.sha256(out.cpu().numpy()) == meta["iq_sha256"]


# ---- host orchestrator + command-line front end ------------------------------
def test_host_orchestrator_descriptors_drive_the_kernels():
    """navigation file -> libgpshost descriptors -> kernels == the reference's stream (no reference-made input)."""
    from pluto_gps_sim_b200 import hostapi

    meta = ol.load_golden_meta("static12")
    nav = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
    with hostapi.Scenario(nav, llh=(30.286502, 120.032669, 100), sample_rate=2600000) as sc, \
            Synthesizer(max_chan=12, max_epochs=10) as s:
        iq = s.synth(sc.next(10))
    assert ol.sha256(iq) == meta["iq_sha256"]


@pytest.mark.parametrize("scenario,args,epochs", [
    ("static12", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-d", "1.0"], 10),
    ("static12", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-d", "1.0", "-b", "3"], 10),
    ("allsky32", ["-l", "30.286502,120.032669,100", "-s", "10000000", "-d", "2.0", "-n", "32", "-b", "8"], 20),
    ("static12", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-d", "1.0", "-b", "1", "-g", "2"], 10),
])
def test_command_line_front_end_reproduces_the_reference_stream(tmp_path, scenario, args, epochs):
    """gpsiq_sim with the reference's options writes byte for byte what the reference pushes to the SDR."""
    import hashlib
    import subprocess
    from pluto_gps_sim_b200 import hostapi

    import torch
    if "-g" in args and torch.cuda.device_count() < int(args[args.index("-g") + 1]):
        pytest.skip("needs %s GPUs" % args[args.index("-g") + 1])
    meta = ol.load_golden_meta(scenario)
    nav = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz" if scenario == "static12" else "allsky32_synth.14n.gz")
    out = tmp_path / "iq.bin"
    r = subprocess.run([hostapi.SIM_PATH, "-e", nav, "-o", str(out)] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = out.read_bytes()
    assert len(data) == epochs * 300000 * 4
    assert hashlib.sha256(data).hexdigest() == meta["iq_sha256"], r.stderr


def test_front_end_radio_sink_pushes_the_reference_stream_once_each(tmp_path):
    """gpsiq_sim -r on the GPU: the reference's libiio set-up calls (tests/golden/iio_calls.json, recorded from the
    unmodified reference) and then the golden stream, one 300000-pair buffer per push, none lost or repeated.
    libiio is the capture backend oracle/libfakeiio.so, dlopen()ed by the sink."""
    import hashlib
    import json
    import subprocess
    from pluto_gps_sim_b200 import hostapi

    fake = os.path.join(ol.ORACLE_DIR, "libfakeiio.so")
    if not os.path.exists(fake):
        subprocess.run(["make", "-C", ol.ORACLE_DIR, "libfakeiio.so"], check=True, stdout=subprocess.DEVNULL)
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    env = dict(os.environ, GPSSINK_IIO_LIB=fake, FAKE_IIO_LOG=str(log), FAKE_IIO_OUT=str(out), FAKE_IIO_EPOCHS="1000",
               FAKE_IIO_NO_DEFAULT="1")
    nav = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
    r = subprocess.run([hostapi.SIM_PATH, "-e", nav, "-r", "-l", "30.286502,120.032669,100", "-s", "2600000", "-d", "1.0",
                        "-b", "4", "-A", "-35.5", "-B", "3.0", "-U", "usb:1.2.5"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    golden = json.load(open(os.path.join(ol.GOLDEN, "iio_calls.json")))["uri_gain_bandwidth"]
    assert log.read_text().splitlines() == golden["calls"]
    assert hashlib.sha256(out.read_bytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"], r.stderr


@pytest.mark.parametrize("gpus", [1, 2])
def test_config2_full_300s_user_motion_stream_through_the_front_end(gpus):
    """BASELINE config[2] in full: 3000 epochs (300 s, 3.6 GB) of circle.csv user motion, navigation file in,
    bytes out, SHA-256 against the reference's own 300 s run (tools/gen_golden_long.py).  Crosses ten 30 s
    refreshes (NAV frame rebuild, re-allocation) and exercises ~3 of the line kernel's patch-path tiles."""
    import hashlib
    import json
    import subprocess
    import refdump
    from pluto_gps_sim_b200 import hostapi

    import torch
    circle = os.path.join(refdump.REF_DIR, "circle.csv")
    if not os.path.exists(circle):
        pytest.skip("needs the reference's circle.csv (oracle/_ref)")
    if torch.cuda.device_count() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    with open(os.path.join(ol.GOLDEN, "circle12_300s_meta.json")) as f:
        meta = json.load(f)
    nav = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
    p = subprocess.Popen([hostapi.SIM_PATH, "-e", nav, "-u", circle, "-s", "2600000", "-d", "300", "-b", "250", "-o", "-",
                          "-g", str(gpus)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    sha, total = hashlib.sha256(), 0
    while True:
        buf = p.stdout.read(1 << 24)
        if not buf:
            break
        sha.update(buf)
        total += len(buf)
    err = p.stderr.read().decode()
    assert p.wait() == 0, err
    assert total == 3000 * 300000 * 4
    assert sha.hexdigest() == meta["iq_sha256"], err


def _stream_checksums(s, desc, sizes, ahead=2):
    """The golden stream through the pipelined submit / fetch pair -> per-epoch device checksums."""
    import torch

    st = torch.cuda.Stream()
    sums = []
    with torch.cuda.stream(st):
        cap = max(sizes)
        bufs = [torch.empty(cap * 300000 * 2, dtype=torch.int16, device="cuda") for _ in range(2)]
        descs, e = [], 0
        for n in sizes:
            descs.append(torch.from_numpy(desc[e:e + n].copy().view(np.uint8).reshape(-1)).cuda())
            e += n
        torch.cuda.synchronize()
        sub = 0
        for k in range(len(sizes)):
            while sub < len(sizes) and sub <= k + ahead:
                s.submit_device(descs[sub].data_ptr(), sizes[sub], st.cuda_stream)
                sub += 1
            s.fetch_device(bufs[k & 1].data_ptr(), st.cuda_stream)
            st.synchronize()
            sums += [int(x) for x in s.checksum_device(bufs[k & 1].data_ptr(), sizes[k])]
    s.check_device()
    return sums


def test_slice_level_chain_one_head_scan_per_batch(monkeypatch):
    """Level 5 of the carrier scan (csrc/nco_scan.cuh: slice_chain_group / slice_verify; k_carr_slice, k_carr_final,
    k_carr_final_groups): a batch's exact chain is ONE head scan + a translation, its groups are chained in parallel
    afterwards.  The 310-epoch user-motion golden (5 tie-capable steps, a re-allocation at 30 s, negative and positive
    Doppler) in batches of several groups: the reference's per-epoch checksums, most (slot, batch) chains passed by
    translation, no internal inconsistency flagged; and the same stream with the level switched off."""
    meta = ol.load_golden_meta("circle12")
    desc = ol.load_golden_desc("circle12")
    sizes = [100, 128, 82]
    with Synthesizer(max_chan=12, max_epochs=128) as s:
        sums = _stream_checksums(s, desc, sizes)
        translated, serial = s.slice_stats
        trace_on = s.carrier_trace(sizes[-1])
    assert sums == meta["epoch_checksums"]
    # the first batch starts from re-seeded phases (serial by construction); afterwards translation is the rule
    assert translated >= 8 and serial >= 10, (translated, serial)     # (10 of the 12 slots are active)
    monkeypatch.setenv("GPSIQ_SLICE_SPEC", "0")
    with Synthesizer(max_chan=12, max_epochs=128) as s:
        sums_off = _stream_checksums(s, desc, sizes)
        assert s.slice_stats[0] == 0
        trace_off = s.carrier_trace(sizes[-1])
    assert sums_off == meta["epoch_checksums"]
    assert np.array_equal(trace_on, trace_off)


def test_slice_level_chain_with_a_poor_estimate_falls_back_to_the_serial_chain():
    """Exactness never depends on the start-phase estimate: a batch speculated from a deliberately wrong estimate is
    chained serially (slice_verify finds no match), a good one is translated -- same samples, same end phases."""
    import torch

    base = ol.load_golden_desc("static12")
    E = 200                                            # 4 groups of 64 epochs, the last one ragged
    desc = np.concatenate([base] * 20)[:E].copy()
    desc["flags"] = 0
    first = desc.copy()
    first[0]["flags"] = capi.FLAG_RESET_CARRIER
    stride = 25
    want_sums, want_end = None, None
    for wrong in (0.0, 3e-7):
        with Synthesizer(max_chan=12, max_epochs=E) as s:
            d0 = torch.from_numpy(first.view(np.uint8).reshape(-1)).cuda()
            d1 = torch.from_numpy(desc.view(np.uint8).reshape(-1)).cuda()
            out = torch.empty(E * 300000 * 2, dtype=torch.int16, device="cuda")
            s.synth_device(d0.data_ptr(), E, out.data_ptr())            # batch 0: from the allocation phases
            torch.cuda.synchronize()
            t0, s0 = s.slice_stats
            est = torch.from_numpy(s.carrier.copy()).cuda()
            est = torch.remainder(est + wrong, 1.0)
            s.prepare_device(d1.data_ptr(), E)
            s.estimate_from_device(est.data_ptr())
            s.speculate_device(d1.data_ptr(), E)
            s.chain_device(d1.data_ptr(), E)
            s.render_device(d1.data_ptr(), E, out.data_ptr())
            torch.cuda.synchronize()
            s.check_device()
            t1, s1 = s.slice_stats
            sums = [int(x) for x in s.checksum_device(out.data_ptr(), E)]
            end = s.carrier.copy()
            trace = s.carrier_trace(E)
        if wrong == 0.0:
            assert t1 - t0 >= 10 and (t1 - t0) + (s1 - s0) == 12, (t0, s0, t1, s1)   # (nearly) every slot translated
            want_sums, want_end, want_trace = sums, end, trace
            # against the oracle: a stride of epochs, each from the chain's own start phase
            for e in range(stride, E, stride):
                ref, ph = ol.oracle_synth(desc[e:e + 1], 300000, carr_state=trace[e - 1].copy())
                assert int(checksum_host(ref[0])) == sums[e], e
                assert np.array_equal(ph[0], trace[e]), e
        else:
            assert s1 - s0 == 12 and t1 == t0, (t0, s0, t1, s1)       # nothing fits: all serial
            assert sums == want_sums and np.array_equal(end, want_end) and np.array_equal(trace, want_trace)
