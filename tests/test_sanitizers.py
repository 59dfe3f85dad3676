"""Host C++ under sanitizers (the reference has none and has real races in its hand-off).

Builds the host sources (gpshost.cpp, gpssink.cpp) together with small drivers (tests/native/) under
AddressSanitizer + UBSan and under ThreadSanitizer and runs them: mutated navigation files through both RINEX readers
and the epoch loop, the sink's writer thread against a ring-buffer producer, the multi-threaded descriptor generator.
The guards these runs led to (month index, Kepler iteration cap, antenna-table index, stale fixed-column reads) are in
host/gpshost.cpp; tests/test_host_orchestrator.py pins that valid files are unaffected."""
import gzip
import os
import subprocess

import pytest

import oracle_lib as ol

HOST = os.path.join(ol.REPO, "pluto_gps_sim_b200", "host")
NATIVE = os.path.join(ol.REPO, "tests", "native")
SRCS = [os.path.join(HOST, "gpshost.cpp"), os.path.join(HOST, "gpssink.cpp")]


def build(tmp, name, sanitize):
    exe = str(tmp / name)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-ffp-contract=off", "-fsanitize=" + sanitize,
           "-I", os.path.join(ol.REPO, "include"), "-o", exe, os.path.join(NATIVE, name + ".cpp")] + SRCS + ["-lz", "-ldl", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitizer" in (r.stderr or "").lower() and "cannot find" in r.stderr:
        pytest.skip("sanitizer runtime not installed: " + r.stderr.splitlines()[-1])
    assert r.returncode == 0, r.stderr
    return exe


def run(cmd, env=None, timeout=600):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))
    if "unexpected memory mapping" in r.stderr or "Shadow memory range interleaves" in r.stderr:
        pytest.skip("sanitizer runtime cannot map its shadow memory here (address-space layout): " + r.stderr.splitlines()[0])
    report = [l for l in (r.stdout + r.stderr).splitlines() if "runtime error" in l or "Sanitizer" in l]
    assert r.returncode == 0 and not report, (r.returncode, report[:5], r.stderr[-2000:])
    return r.stdout


@pytest.mark.parametrize("fixture,v3", [("brdc3540_synth.14n.gz", 0), ("brdc3540_synth.14p.gz", 1)])
def test_mutated_navigation_files_under_asan_ubsan(tmp_path, fixture, v3):
    exe = build(tmp_path, "fuzz_nav", "address,undefined")
    plain = tmp_path / "nav.txt"
    plain.write_bytes(gzip.open(os.path.join(ol.GOLDEN, fixture)).read())
    out = run([exe, str(plain), str(v3), "800", str(tmp_path / "mutant.nav")], env={"ASAN_OPTIONS": "detect_leaks=1"})
    assert "fuzz_nav:" in out


def test_sink_writer_thread_under_tsan(tmp_path):
    exe = build(tmp_path, "stress_sink", "thread")
    assert "stress_sink: ok" in run([exe, str(tmp_path)])


def test_threaded_descriptor_generation_under_tsan(tmp_path):
    exe = build(tmp_path, "host_threads", "thread")
    assert "host_threads: ok" in run([exe, os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz"), "8"])
