"""The literal drop-in of INTEGRATION.md, compiled and run.

oracle/_ref/ref_dropin (oracle/Makefile, oracle/dropin/) is the UNMODIFIED reference -- its main(), option parsing,
RINEX reader, orbit/range/NAV code, 30 s refresh, libiio TX thread and handshake -- with only the per-sample loop
(plutogpssim.c:2690-2756) replaced by the two C-ABI calls gpsiq_make_desc + gpsiq_synth, one epoch per call.  What
reaches the (capture) libiio backend must be the reference's stream byte for byte.

  CPU  : the binary runs against tests/mock/mock_gpsiq.c (test infrastructure, oracle-backed) placed first on
         LD_LIBRARY_PATH: proves the patch of INTEGRATION.md compiles against include/gpsiq.h and that descriptors
         made from the reference's live channel state + a carrier phase owned by the context reproduce the stream.
  -m gpu: the same binary with the real libgpsiq.so (its RUNPATH): the reference's host side driving the B200 kernels.

The binary is built where /root/reference exists (this container) and travels to the GPU box with oracle/_ref/;
nothing here reads /root/reference at run time.  (File name: runs last, after the kernels' own parity tests.)"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from pluto_gps_sim_b200 import checksum_host

DROPIN = os.path.join(ol.ORACLE_DIR, "_ref", "ref_dropin")
NAV12 = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
CIRCLE = os.path.join(ol.ORACLE_DIR, "_ref", "circle.csv")
N = 300000

needs_dropin = pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/ref_dropin not built (needs /root/reference at build time)")


def run_dropin(tmp_path, args, epochs, lib_dir=None):
    out = tmp_path / "pushed.bin"
    env = dict(os.environ, FAKE_IIO_OUT=str(out), FAKE_IIO_EPOCHS=str(epochs))
    if lib_dir:
        env["LD_LIBRARY_PATH"] = str(lib_dir) + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([DROPIN, "-e", NAV12] + args, env=env, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert "Error pushing buf" in r.stderr, r.stderr          # ended through the reference's own shutdown path
    return np.fromfile(out, np.int16).reshape(-1, N, 2), r.stderr


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("mock_gpsiq")
    subprocess.run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ol.REPO, "include"),
                    "-o", str(d / "libgpsiq.so"), os.path.join(ol.REPO, "tests", "mock", "mock_gpsiq.c"),
                    os.path.join(ol.ORACLE_DIR, "gpsiq_oracle.c"), "-lm"], check=True)
    return d


@needs_dropin
def test_dropin_patch_reproduces_the_reference_stream_cpu(tmp_path, mock_dir):
    iq, err = run_dropin(tmp_path, ["-l", "30.286502,120.032669,100", "-s", "2600000"], 10, lib_dir=mock_dir)
    assert iq.shape[0] == 10
    assert hashlib.sha256(iq.tobytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"], err


@needs_dropin
def test_dropin_patch_user_motion_across_the_refresh_cpu(tmp_path, mock_dir):
    """310 epochs of circle.csv: the 30 s re-allocation makes `fresh` slots, whose carrier phase must come from the
    reference's allocateChannel and not from the context."""
    iq, err = run_dropin(tmp_path, ["-u", CIRCLE, "-s", "2600000"], 310, lib_dir=mock_dir)
    assert iq.shape[0] == 310
    assert hashlib.sha256(iq.tobytes()).hexdigest() == ol.load_golden_meta("circle12")["iq_sha256"], err


@needs_dropin
def test_dropin_without_a_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([DROPIN, "-e", NAV12, "-l", "30.286502,120.032669,100"], capture_output=True, text=True,
                       env=dict(os.environ, FAKE_IIO_EPOCHS="1"), cwd=str(tmp_path), timeout=120)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@needs_dropin
def test_reference_host_side_drives_the_b200_kernels_gpu(tmp_path):
    """The reference's own main() + TX thread with the CUDA library in place of its sample loop: 3 s static, then
    5 s of user motion, against the reference's goldens (full SHA for 1 s, per-epoch checksums beyond)."""
    iq, err = run_dropin(tmp_path, ["-l", "30.286502,120.032669,100", "-s", "2600000"], 10)
    assert hashlib.sha256(iq.tobytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"], err
    iq, err = run_dropin(tmp_path, ["-u", CIRCLE, "-s", "2600000"], 50)
    assert [int(checksum_host(iq[e])) for e in range(50)] == ol.load_golden_meta("circle12")["epoch_checksums"][:50], err


@pytest.mark.gpu
def test_plain_c_pipeline_example_on_the_gpu(tmp_path):
    """tests/native/pipeline_example.c (the C usage INTEGRATION.md documents) linked against the real libraries:
    navigation file in, the reference's bytes out, at three batch sizes."""
    from pluto_gps_sim_b200 import capi, hostapi
    exe = tmp_path / "pipeline_example"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-O2", "-I", os.path.join(ol.REPO, "include"), "-o", str(exe),
                    os.path.join(ol.REPO, "tests", "native", "pipeline_example.c"), "-L" + libdir, "-lgpshost", "-lgpsiq",
                    "-Wl,-rpath," + libdir], check=True)
    assert os.path.dirname(hostapi.LIB_PATH) == libdir
    for batch in ("3", "4", "16"):
        out = tmp_path / ("iq%s.bin" % batch)
        r = subprocess.run([str(exe), NAV12, str(out), "10", batch], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert hashlib.sha256(out.read_bytes()).hexdigest() == ol.load_golden_meta("static12")["iq_sha256"]
