"""Control flow of the command-line front end (gpsiq_sim) on the CPU.

The GPU tests (tests/test_gpu_parity.py) run the real thing.  Here a COPY of the gpsiq_sim binary is placed in a
scratch directory next to libgpshost.so (the product's) and tests/mock/mock_gpsiq.c (TEST INFRASTRUCTURE: the seven
libgpsiq entry points gpsiq_sim calls, computed by the parity oracle), so that option parsing, host orchestration,
batching with two buffers in flight and the hand-off to the sink's writer thread are checked end to end against
the reference's streams without a GPU.  What is NOT covered here is the CUDA library -- nothing in this file is a
statement about it, and the product package never contains or loads the mock."""
import hashlib
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

from pluto_gps_sim_b200 import checksum_host, hostapi

N = 300000
FAKE_IIO = os.path.join(ol.ORACLE_DIR, "libfakeiio.so")
NAV12 = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
NAV32 = os.path.join(ol.GOLDEN, "allsky32_synth.14n.gz")
STATIC = ["-l", "30.286502,120.032669,100", "-s", "2600000"]


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    d = tmp_path_factory.mktemp("front_end")
    shutil.copy(hostapi.SIM_PATH, d / "gpsiq_sim")
    shutil.copy(hostapi.LIB_PATH, d / "libgpshost.so")
    subprocess.run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ol.REPO, "include"),
                    "-o", str(d / "libgpsiq.so"), os.path.join(ol.REPO, "tests", "mock", "mock_gpsiq.c"),
                    os.path.join(ol.ORACLE_DIR, "gpsiq_oracle.c"), "-lm"], check=True)
    if not os.path.exists(FAKE_IIO):
        subprocess.run(["make", "-C", ol.ORACLE_DIR, "libfakeiio.so"], check=True, stdout=subprocess.DEVNULL)

    def popen(args, env=None):
        return subprocess.Popen([str(d / "gpsiq_sim")] + args, stderr=subprocess.PIPE, env=dict(os.environ, **(env or {})))

    def run(args, env=None, check=True, binary=False):
        r = subprocess.run([str(d / "gpsiq_sim")] + args, capture_output=True, env=dict(os.environ, **(env or {})))
        r.stderr = r.stderr.decode(errors="replace")
        if not binary:
            r.stdout = r.stdout.decode(errors="replace")
        if check:
            assert r.returncode == 0, r.stderr
        return r
    run.popen = popen
    return run


@pytest.mark.parametrize("batch", ["4", "3", "128"])      # even, ragged last batch, one batch
def test_static12_stream_byte_identical_to_the_reference(sim, tmp_path, batch):
    out = tmp_path / "iq.bin"
    r = sim(["-e", NAV12, "-o", str(out), "-d", "1.0", "-b", batch] + STATIC)
    data = out.read_bytes()
    assert len(data) == 10 * N * 4
    assert hashlib.sha256(data).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"], r.stderr
    assert "10 buffers to the sink" in r.stderr and "Using static location mode." in r.stderr


def test_32_channels_10MSps_first_epochs(sim, tmp_path):
    out = tmp_path / "iq.bin"
    sim(["-e", NAV32, "-o", str(out), "-l", "30.286502,120.032669,100", "-s", "10000000", "-d", "0.3", "-n", "32", "-b", "2"])
    iq = np.fromfile(out, np.int16).reshape(3, N, 2)
    want = ol.load_golden_meta("allsky32")["epoch_checksums"]
    assert [int(checksum_host(iq[e])) for e in range(3)] == want[:3]


def test_user_motion_file_first_epochs(sim, tmp_path):
    circle = os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv")
    if not os.path.exists(circle):
        pytest.skip("oracle/_ref/circle.csv (the reference's motion fixture) not present")
    out = tmp_path / "iq.bin"
    r = sim(["-e", NAV12, "-o", str(out), "-u", circle, "-s", "2600000", "-d", "0.5", "-b", "2"])
    iq = np.fromfile(out, np.int16).reshape(5, N, 2)
    assert [int(checksum_host(iq[e])) for e in range(5)] == ol.load_golden_meta("circle12")["epoch_checksums"][:5]
    assert "Using user motion mode." in r.stderr


def test_user_motion_across_the_30s_refresh_full_golden(sim, tmp_path):
    """31 s of circle.csv: NAV frame rebuild and channel re-allocation at 30 s (plutogpssim.c:2764-2806) happen inside
    the front end's batch loop; SHA-256 of the 372 MB stream against the reference's own run."""
    circle = os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv")
    if not os.path.exists(circle):
        pytest.skip("oracle/_ref/circle.csv (the reference's motion fixture) not present")
    h = hashlib.sha256()
    r = sim(["-e", NAV12, "-o", "-", "-u", circle, "-s", "2600000", "-d", "31.0", "-b", "64"], binary=True)
    h.update(r.stdout)
    assert len(r.stdout) == 310 * N * 4
    assert h.hexdigest() == ol.load_golden_meta("circle12")["iq_sha256"]


def test_stdout_sink_and_null_sink(sim):
    res = sim(["-e", NAV12, "-d", "0.2", "-b", "1"] + STATIC)                 # discard
    assert "2 buffers to the sink" in res.stderr and "600000 samples" in res.stderr
    res = sim(["-e", NAV12, "-d", "0.2", "-o", "-"] + STATIC, binary=True)    # "-" = stdout: bytes only, messages on stderr
    iq = np.frombuffer(res.stdout, np.int16).reshape(2, N, 2)
    assert [int(checksum_host(iq[e])) for e in range(2)] == ol.load_golden_meta("static12")["epoch_checksums"][:2]


def test_radio_sink_through_the_front_end(sim, tmp_path):
    """-r: the stream goes out through libiio in 300000-pair buffers, each exactly once (the reference's own
    hand-off can drop or repeat one), after the reference's radio set-up with its options."""
    import json
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    env = {"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_LOG": str(log), "FAKE_IIO_OUT": str(out), "FAKE_IIO_EPOCHS": "1000",
           "FAKE_IIO_NO_DEFAULT": "1"}
    r = sim(["-e", NAV12, "-r", "-d", "1.0", "-b", "3", "-A", "-35.5", "-B", "3.0", "-U", "usb:1.2.5"] + STATIC, env=env)
    golden = json.load(open(os.path.join(ol.GOLDEN, "iio_calls.json")))["uri_gain_bandwidth"]
    assert log.read_text().splitlines() == golden["calls"]
    assert hashlib.sha256(out.read_bytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"]
    assert "Gain: -35.5dB" in r.stderr and "10 buffers to the sink" in r.stderr


def test_start_up_text_is_the_reference_s(sim, tmp_path):
    """-v -r: mode, ionosphere/UTC block, gain, RINEX date, start time and channel table, line for line what the
    unmodified reference prints (recorded by tools/gen_iio_golden.py)."""
    import json
    want = json.load(open(os.path.join(ol.GOLDEN, "iio_calls.json")))["_banner_v"]["stderr"]
    env = {"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_EPOCHS": "1000"}
    r = sim(["-e", NAV12, "-v", "-r", "-d", "0.1"] + STATIC, env=env)
    got = r.stderr.splitlines()
    assert got[:len(want)] == want, "\n".join(got)


@pytest.mark.skipif(not os.path.exists(os.path.join(ol.ORACLE_DIR, "_ref", "ref_verbatim")), reason="needs the compiled reference (oracle/_ref)")
@pytest.mark.parametrize("extra", [
    ["-u", "@circle", "-s", "2600000", "-t", "2014/12/20,00:30:00"],
    ["-c", "-2758918.64,4772301.12,3197889.44", "-s", "3000000", "-i", "-A", "-12.5"],
    ["-l", "30.286502,120.032669,100", "-s", "2600000", "-T", "x", "-t", "2021/06/07,08:09:10"],
])
def test_start_up_text_equals_a_live_run_of_the_reference(sim, tmp_path, extra):
    """Other option combinations, against the reference binary run here: mode line, -v block, gain, RINEX date, start
    time (incl. -T overwrite) and the channel table (azimuth, elevation, range, ionospheric delay per PRN)."""
    extra = [os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv") if a == "@circle" else a for a in extra]
    ref = subprocess.run([os.path.join(ol.ORACLE_DIR, "_ref", "ref_verbatim"), "-e", NAV12, "-v"] + extra, capture_output=True, text=True,
                         env=dict(os.environ, FAKE_IIO_EPOCHS="1"), cwd=str(tmp_path))
    want = ref.stderr.split("Error pushing buf")[0].splitlines()
    assert len(want) > 12
    got = sim(["-e", NAV12, "-v", "-r", "-d", "0.1"] + extra, env={"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_EPOCHS": "1000"}).stderr.splitlines()
    assert got[:len(want)] == want, "\n".join(got[:len(want)]) + "\n--- reference ---\n" + "\n".join(want)


def test_radio_sink_failures_end_the_run_with_an_error(sim, tmp_path):
    env = {"GPSSINK_IIO_LIB": str(tmp_path / "absent-libiio.so")}
    r = sim(["-e", NAV12, "-r", "-d", "0.2"] + STATIC, env=env, check=False)
    assert r.returncode == 1 and "libiio is not available" in r.stderr
    # the device refuses the third buffer: the run stops, exit status 1, tear-down still happens
    log = tmp_path / "calls.log"
    env = {"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_LOG": str(log), "FAKE_IIO_EPOCHS": "3", "FAKE_IIO_OUT": str(tmp_path / "p.bin")}
    r = sim(["-e", NAV12, "-r", "-d", "1.0", "-b", "2"] + STATIC, env=env, check=False)
    assert r.returncode == 1 and "Error pushing buf -1" in r.stderr
    assert log.read_text().splitlines()[-1] == "context_destroy"
    assert os.path.getsize(tmp_path / "p.bin") == 3 * N * 4


def test_runs_until_interrupted_and_shuts_the_radio_down(sim, tmp_path):
    """-d 0: like the reference, run until SIGINT/SIGTERM; then the batches in flight are delivered, the TX LO is
    powered down and the context destroyed (plutogpssim.c:2014-2022, 2160-2178), exit status 0."""
    import signal
    import time
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    env = {"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_LOG": str(log), "FAKE_IIO_OUT": str(out), "FAKE_IIO_EPOCHS": "1000000"}
    p = sim.popen(["-e", NAV12, "-r", "-d", "0", "-b", "2"] + STATIC, env=env)
    deadline = time.time() + 60
    while time.time() < deadline and (not out.exists() or out.stat().st_size < 4 * N * 4):
        time.sleep(0.05)
    p.send_signal(signal.SIGINT)
    err = p.communicate(timeout=60)[1].decode()
    assert p.returncode == 0, err
    calls = log.read_text().splitlines()
    assert calls[-5:] == ["attr ad9361-phy/altvoltage1 powerdown = true", "buffer_destroy", "disable cf-ad9361-dds-core-lpc/voltage0",
                          "disable cf-ad9361-dds-core-lpc/voltage1", "context_destroy"]
    iq = np.fromfile(out, np.int16)
    assert iq.size % (2 * N) == 0 and iq.size >= 4 * 2 * N
    iq = iq.reshape(-1, N, 2)
    want = ol.load_golden_meta("static12")["epoch_checksums"]
    k = min(len(want), len(iq))
    assert [int(checksum_host(iq[e])) for e in range(k)] == want[:k]
    assert "%d buffers to the sink" % len(iq) in err


def test_option_errors_match_the_reference_s_messages(sim, tmp_path):
    assert "ERROR: GPS ephemeris file is not specified." in sim(["-l", "1,2,3", "-d", "1"], check=False).stderr
    assert "ERROR: Invalid sampling frequency." in sim(["-e", NAV12, "-s", "999999"], check=False).stderr
    assert "ERROR: Invalid date and time." in sim(["-e", NAV12, "-t", "2014-12-20"], check=False).stderr
    r = sim(["-e", str(tmp_path / "nope.14n"), "-l", "1,2,3"], check=False)
    assert r.returncode == 1 and "ERROR" in r.stderr
    r = sim(["-e", NAV12, "-o", str(tmp_path / "no" / "dir" / "x.bin")] + STATIC, check=False)
    assert r.returncode == 1 and "cannot open" in r.stderr
    assert sim(["-e", NAV12, "-f"] + STATIC, check=False).returncode == 1


def test_plain_c_pipeline_example_compiles_and_reproduces_the_stream(tmp_path):
    """tests/native/pipeline_example.c = the C usage INTEGRATION.md documents, as a complete C11 program: the three
    headers are plain C, the three-library pipeline with two buffers in flight gives the reference's bytes."""
    d = tmp_path
    subprocess.run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ol.REPO, "include"),
                    "-o", str(d / "libgpsiq.so"), os.path.join(ol.REPO, "tests", "mock", "mock_gpsiq.c"),
                    os.path.join(ol.ORACLE_DIR, "gpsiq_oracle.c"), "-lm"], check=True)
    shutil.copy(hostapi.LIB_PATH, d / "libgpshost.so")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-O2", "-I", os.path.join(ol.REPO, "include"), "-o", str(d / "pipeline_example"),
                    os.path.join(ol.REPO, "tests", "native", "pipeline_example.c"), "-L" + str(d), "-lgpshost", "-lgpsiq",
                    "-Wl,-rpath,$ORIGIN"], check=True)
    for batch in ("3", "4", "16"):
        out = d / ("iq%s.bin" % batch)
        r = subprocess.run([str(d / "pipeline_example"), NAV12, str(out), "10", batch], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha256(out.read_bytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"]


@pytest.mark.skipif(not os.path.exists(os.path.join(ol.ORACLE_DIR, "_ref", "ref_verbatim")), reason="needs the compiled reference (oracle/_ref)")
@pytest.mark.parametrize("args", [
    ["-e", NAV12, "-l", "30.286502,120.032669,100", "-t", "2015/01/01,00:00:00"],          # start time outside the file
    ["-e", NAV12, "-l", "30.286502,120.032669,100", "-t", "2014/13/20,00:00:00"],          # invalid date
    ["-e", NAV12, "-u", "/nonexistent/motion.csv"],                                         # motion file cannot be opened
    ["-e", NAV12, "-l", "30.286502,120.032669,100", "-s", "500000"],                        # sampling frequency
    ["-l", "30.286502,120.032669,100"],                                                     # no navigation file
])
def test_fatal_errors_are_worded_and_ordered_like_the_reference_s(sim, tmp_path, args):
    """Exit status 1 and the same stderr as the reference binary run here, line for line."""
    ref = subprocess.run([os.path.join(ol.ORACLE_DIR, "_ref", "ref_verbatim")] + args, capture_output=True, text=True,
                         env=dict(os.environ, FAKE_IIO_EPOCHS="1"), cwd=str(tmp_path))
    got = sim(args, check=False)
    assert ref.returncode == 1 and got.returncode == 1
    assert got.stderr.splitlines() == ref.stderr.splitlines()


def test_user_motion_wins_whatever_the_option_order(sim, tmp_path):
    """-u clears the reference's staticLocationMode and -l / -c never set it again (plutogpssim.c:2301-2318, 2403):
    `-u file -l ...` is a user-motion run, byte for byte the same stream as `-u file` alone."""
    circle = os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv")
    if not os.path.exists(circle):
        pytest.skip("oracle/_ref/circle.csv (the reference's motion fixture) not present")
    outs = []
    for extra in ([], ["-l", "10.0,20.0,30"], ["-c", "1.0,2.0,3.0"]):
        out = tmp_path / ("iq%d.bin" % len(outs))
        r = sim(["-e", NAV12, "-o", str(out), "-u", circle] + extra + ["-s", "2600000", "-d", "0.3", "-b", "2"])
        assert "Using user motion mode." in r.stderr and "static location" not in r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] == outs[2]
    iq = np.frombuffer(outs[0], np.int16).reshape(3, N, 2)
    assert [int(checksum_host(iq[e])) for e in range(3)] == ol.load_golden_meta("circle12")["epoch_checksums"][:3]


@pytest.mark.parametrize("gpus,batch", [("2", "3"), ("3", "1"), ("4", "128")])
def test_several_gpus_same_stream(sim, tmp_path, gpus, batch):
    """-g N: the front end's multi-device loop (several batches submitted ahead, several being fetched, buffers
    recycled through the sink) must deliver the same bytes in the same order."""
    out = tmp_path / "iq.bin"
    r = sim(["-e", NAV12, "-o", str(out), "-d", "1.0", "-b", batch, "-g", gpus] + STATIC)
    assert hashlib.sha256(out.read_bytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"], r.stderr
    assert "10 buffers to the sink" in r.stderr


@pytest.mark.parametrize("sig", ["SIGINT", "SIGQUIT"])
def test_stop_request_while_transmitting_is_fast_and_powers_the_lo_down(sim, tmp_path, sig):
    """With -r the batches are small and a stop request makes the sink discard what is queued: the process ends within
    about a second (the reference: one 0.1 s epoch), through the orderly radio shut-down, for every signal the reference
    handles (SIGINT, SIGTERM, SIGQUIT; plutogpssim.c:2282-2284)."""
    import signal
    import time
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    env = {"GPSSINK_IIO_LIB": FAKE_IIO, "FAKE_IIO_LOG": str(log), "FAKE_IIO_OUT": str(out), "FAKE_IIO_EPOCHS": "100000",
           "FAKE_IIO_NO_DEFAULT": "0", "FAKE_IIO_PUSH_SLEEP_MS": "100"}
    p = sim.popen(["-e", NAV12, "-r", "-d", "0", "-b", "128"] + STATIC, env=env)
    time.sleep(3.0)
    t0 = time.time()
    p.send_signal(getattr(signal, sig))
    try:
        _, err = p.communicate(timeout=30)
    except subprocess.TimeoutExpired:
        p.kill()
        raise
    dt = time.time() - t0
    assert p.returncode == 0, err.decode(errors="replace")
    assert dt < 5.0, dt                       # (-b 128 is capped for the radio; nothing queued is transmitted)
    calls = log.read_text().splitlines()
    assert calls[-1].startswith("context_destroy") and any("powerdown" in c for c in calls[-5:])
