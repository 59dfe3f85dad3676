"""Output transport (include/gpssink.h) on the CPU.

The radio sink is pinned against the reference itself: tests/golden/iio_calls.json holds the libiio calls the
UNMODIFIED reference makes (tools/gen_iio_golden.py: ref_verbatim linked with the capture backend
oracle/fake_iio.c); here the product's sink dlopen()s the same backend as a shared library
(oracle/libfakeiio.so) and must make the same calls, then push every buffer exactly once."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

sys.path.insert(0, os.path.join(ol.REPO, "tools"))
import refdump  # noqa: E402
from pluto_gps_sim_b200 import sinkapi  # noqa: E402

FAKE_IIO = os.path.join(ol.ORACLE_DIR, "libfakeiio.so")
N = sinkapi.PUSH_PAIRS


@pytest.fixture(scope="module", autouse=True)
def _fake_iio_built():
    if not os.path.exists(FAKE_IIO):
        subprocess.run(["make", "-C", ol.ORACLE_DIR, "libfakeiio.so"], check=True, stdout=subprocess.DEVNULL)


def _stream(units, seed=1):
    rng = np.random.default_rng(seed)
    return rng.integers(-2800, 2800, size=(units, N, 2), dtype=np.int16)


def test_header_symbols_all_exported_and_bound():
    hdr = open(os.path.join(ol.REPO, "include", "gpssink.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gpssink_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(sinkapi.SYMBOLS), declared ^ set(sinkapi.SYMBOLS)
    raw = C.CDLL(sinkapi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name)


def test_radio_defaults_and_option_clamping_follow_the_reference():
    cfg = sinkapi.radio_config()
    # plutogpssim.c:2270-2276
    assert (cfg.fs_hz, cfg.bw_hz, cfg.lo_hz, cfg.rfport, cfg.gain_db) == (3000000, 6000000, 1575420000, b"A", -20.0)
    assert (cfg.kernel_buffers, cfg.pairs_per_push, cfg.uri, cfg.hostname) == (12, 300000, None, None)
    for arg, want in [("-35.5", -35.5), ("3", 0.0), ("-80.5", -80.0), ("junk", 0.0)]:     # plutogpssim.c:2367-2370
        assert sinkapi.radio_config([("A", arg)]).gain_db == want
    for arg, want in [("3.0", 3000000), ("9", 5000000), ("0.5", 1000000), ("2.4999996", 2500000)]:   # MHZ(), 2372-2375
        assert sinkapi.radio_config([("B", arg)]).bw_hz == want
    cfg = sinkapi.radio_config([("s", "2600000")])
    assert (cfg.fs_hz, cfg.bw_hz) == (2600000, 6000000)          # -s does not rescale the bandwidth (plutogpssim.c:2325)
    with pytest.raises(sinkapi.SinkError) as ei:
        sinkapi.radio_config([("e", "x")])
    assert ei.value.status == sinkapi.ERR_ARG


def test_file_sink_sync_and_async_keep_order(tmp_path):
    data = _stream(5)
    p1, p2 = tmp_path / "sync.bin", tmp_path / "async.bin"
    with sinkapi.Sink(path=str(p1)) as s:
        for e in range(5):
            s.push(data[e])
        assert s.stats == (5 * N, 5)
    with sinkapi.Sink(path=str(p2)) as s:
        t1 = s.submit(data[:2])            # a batch of two push units in one buffer
        t2 = s.submit(data[2:3])
        s.push(data[3])                    # a synchronous push waits for the queued batches: order is kept
        t3 = s.submit(data[4:])
        s.wait(t3), s.wait(t1), s.wait(t2)
        assert s.stats[0] == 5 * N
    assert p1.read_bytes() == data.tobytes() == p2.read_bytes()


def test_ragged_and_empty_pushes(tmp_path):
    p = tmp_path / "r.bin"
    data = _stream(1)
    with sinkapi.Sink(path=str(p)) as s:
        s.push(data[0, :7])
        s.push(np.zeros((0, 2), np.int16))
        s.push(data[0, 7:])
        with pytest.raises(ValueError):
            s.push(np.zeros(4, np.int32))
        with pytest.raises(sinkapi.SinkError) as ei:
            s.wait(99)
        assert ei.value.status == sinkapi.ERR_ARG
    assert p.read_bytes() == data.tobytes()
    with sinkapi.Sink() as s:              # null sink: counts only
        s.push(data[0])
        assert s.stats == (N, 1)


def test_file_sink_errors_are_loud(tmp_path):
    with pytest.raises(sinkapi.SinkError) as ei:
        sinkapi.Sink(path=str(tmp_path / "no" / "such" / "dir" / "x.bin"))
    assert ei.value.status == sinkapi.ERR_IO
    if os.path.exists("/dev/full"):
        s = sinkapi.Sink(path="/dev/full")
        t = s.submit(_stream(1))
        with pytest.raises(sinkapi.SinkError) as ei:
            s.wait(t)
            s.close()                      # (buffered: the failure may only surface at the flush)
        assert ei.value.status == sinkapi.ERR_IO
        with pytest.raises(sinkapi.SinkError):   # sticky
            s.push(_stream(1)[0])
            s.close()


def test_radio_sink_without_libiio_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.delenv("GPSSINK_IIO_LIB", raising=False)
    with pytest.raises(sinkapi.SinkError) as ei:
        sinkapi.Sink(radio=sinkapi.radio_config(iio_lib=str(tmp_path / "libiio-not-here.so")))
    assert ei.value.status == sinkapi.ERR_BACKEND and "libiio" in str(ei.value)
    try:
        C.CDLL("libiio.so.0")
    except OSError:                       # this image has no libiio: the default search must say so, not crash
        with pytest.raises(sinkapi.SinkError) as ei:
            sinkapi.Sink(radio=sinkapi.radio_config())
        assert ei.value.status == sinkapi.ERR_BACKEND
    # a library that is not libiio
    with pytest.raises(sinkapi.SinkError) as ei:
        sinkapi.Sink(radio=sinkapi.radio_config(iio_lib=os.path.join(ol.ORACLE_DIR, "liboracle.so")))
    assert ei.value.status == sinkapi.ERR_BACKEND and "lacks iio_" in str(ei.value)


_GOLDEN_FILE = json.load(open(os.path.join(ol.GOLDEN, "iio_calls.json")))
GOLDEN_CALLS = {k: v for k, v in _GOLDEN_FILE.items() if not k.startswith("_")}


def _run_sink_in_subprocess(tmp_path, options, default_ctx, units, limit=None, mode="submit"):
    """The capture backend keeps its state in environment-configured globals: one process per case."""
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
from pluto_gps_sim_b200 import sinkapi
N = sinkapi.PUSH_PAIRS
rng = np.random.default_rng(1)
data = rng.integers(-2800, 2800, size=(%d, N, 2), dtype=np.int16)
s = sinkapi.Sink(radio=sinkapi.radio_config(%r, iio_lib=%r))
try:
    if %r == "submit":
        ts = [s.submit(data[:2]), s.submit(data[2:])]
        for t in ts: s.wait(t)
    else:
        for e in range(len(data)): s.push(data[e])
    print("stats", *s.stats)
except sinkapi.SinkError as e:
    print("error", e.status, e)
try:
    s.close()
except sinkapi.SinkError as e:
    print("close-error", e.status)
""" % (ol.REPO, units, [tuple(o) for o in options], FAKE_IIO, mode)
    env = dict(os.environ, FAKE_IIO_LOG=str(log), FAKE_IIO_OUT=str(out), FAKE_IIO_EPOCHS=str(limit or 1000000),
               FAKE_IIO_NO_DEFAULT="0" if default_ctx else "1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout, log.read_text().splitlines(), out.read_bytes()


@pytest.mark.parametrize("case", sorted(GOLDEN_CALLS))
def test_radio_sink_makes_the_reference_s_libiio_calls(tmp_path, case):
    g = GOLDEN_CALLS[case]
    stdout, calls, pushed = _run_sink_in_subprocess(tmp_path, g["options"], g["default_context"], units=4)
    assert calls == g["calls"]
    assert "stats %d 4" % (4 * N) in stdout
    assert pushed == _stream(4).tobytes()          # every 300000-pair buffer exactly once, in order


def test_radio_sink_rejects_partial_buffers_and_reports_a_refused_push(tmp_path):
    d1 = tmp_path / "a"
    d1.mkdir()
    stdout, calls, pushed = _run_sink_in_subprocess(d1, [], True, units=4, limit=3, mode="push")
    # the backend refuses from the third kept buffer on (plutogpssim.c:2153-2156: "Error pushing buf")
    assert "error %d" % sinkapi.ERR_PUSH in stdout and "Error pushing buf -1" in stdout
    assert pushed == _stream(4)[:3].tobytes()
    assert calls[-5:] == GOLDEN_CALLS["defaults_local_context"]["calls"][-5:]   # LO off ... context destroyed still run
    # partial buffers
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from pluto_gps_sim_b200 import sinkapi\n"
            "s = sinkapi.Sink(radio=sinkapi.radio_config(iio_lib=%r))\n"
            "try: s.push(np.zeros((1000, 2), np.int16))\n"
            "except sinkapi.SinkError as e: print('error', e.status)\n") % (ol.REPO, FAKE_IIO)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "error %d" % sinkapi.ERR_ARG in r.stdout, r.stderr


@pytest.mark.skipif(not refdump.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_call_log_golden_is_what_the_reference_does_now():
    sys.path.insert(0, os.path.join(ol.REPO, "tools"))
    import gen_iio_golden as gg
    for name, (opts, dflt) in gg.CASES.items():
        assert gg.reference_log(opts, dflt) == GOLDEN_CALLS[name]["calls"], name
    assert gg.reference_banner() == _GOLDEN_FILE["_banner_v"]["stderr"]


def test_abort_discards_what_is_queued_and_still_shuts_the_radio_down(tmp_path):
    """gpssink_abort (what gpsiq_sim does on SIGINT / SIGTERM / SIGQUIT): the writer stops at the next 300000-pair push
    unit, queued batches are dropped (their tickets complete), and close still runs the reference's shut-down sequence
    (TX LO powered down, buffer destroyed, channels disabled, context destroyed: plutogpssim.c:2160-2178)."""
    log, out = tmp_path / "calls.log", tmp_path / "pushed.bin"
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
from pluto_gps_sim_b200 import sinkapi
N = sinkapi.PUSH_PAIRS
data = np.zeros((6, N, 2), np.int16)
s = sinkapi.Sink(radio=sinkapi.radio_config(iio_lib=%r))
s.abort()
t1 = s.submit(data[:3]); t2 = s.submit(data[3:])
s.wait(t1); s.wait(t2)
print("stats", *s.stats)
s.close()
""" % (ol.REPO, FAKE_IIO)
    env = dict(os.environ, FAKE_IIO_LOG=str(log), FAKE_IIO_OUT=str(out), FAKE_IIO_EPOCHS="1000", FAKE_IIO_NO_DEFAULT="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "stats 0 0" in r.stdout                     # nothing submitted after the abort was pushed
    calls = log.read_text().splitlines()
    assert calls[-5:] == GOLDEN_CALLS["defaults_local_context"]["calls"][-5:]   # LO off ... context destroyed still run
