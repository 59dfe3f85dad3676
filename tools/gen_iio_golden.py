"""TEST INFRASTRUCTURE: record what the UNMODIFIED reference does to the radio.

Runs oracle/_ref/ref_verbatim (the reference translation unit linked with the capture backend oracle/fake_iio.c)
with several radio option sets and stores the backend's call log (set-up and tear-down calls of the reference's TX
thread, plutogpssim.c:2058-2190) in tests/golden/iio_calls.json.  tests/test_sink.py replays the same options
through the product's radio sink (which dlopen()s oracle/libfakeiio.so, the same backend) and expects the same log.

    python tools/gen_iio_golden.py        (needs /root/reference at build time: make -C oracle)
"""
import json
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "oracle", "_ref", "ref_verbatim")
NAV = os.path.join(REPO, "tests", "golden", "brdc3540_synth.14n.gz")
OUT = os.path.join(REPO, "tests", "golden", "iio_calls.json")

# name -> (radio options as (letter, argument) in command-line order, default context available?)
CASES = {
    "defaults_local_context": ([("s", "2600000")], True),
    "no_context_options_falls_back_to_pluto_local": ([("s", "2600000")], False),
    "uri_gain_bandwidth": ([("s", "2600000"), ("A", "-35.5"), ("B", "3.0"), ("U", "usb:1.2.5")], False),
    "hostname_wins_over_uri_and_clamps": ([("N", "pluto2.lan"), ("U", "ip:10.0.0.9"), ("A", "7"), ("B", "9.5")], False),
    "low_clamps_10MSps": ([("s", "10000000"), ("A", "-120"), ("B", "0.2")], True),
}


def reference_banner(extra=("-v",)):
    """What the unmodified reference prints at start-up (mode, -v ionosphere/UTC block, gain, RINEX date, start time,
    channel table: plutogpssim.c:2415-2418, 2487-2495, 2571-2574, 2634-2639), up to its first push."""
    with tempfile.TemporaryDirectory() as wd:
        env = dict(os.environ, FAKE_IIO_EPOCHS="1")
        r = subprocess.run([REF, "-e", NAV, "-l", "30.286502,120.032669,100", "-s", "2600000"] + list(extra), env=env,
                           check=True, capture_output=True, text=True, cwd=wd)
        return r.stderr.split("Error pushing buf")[0].splitlines()


def reference_log(options, default_ctx):
    with tempfile.TemporaryDirectory() as wd:
        log = os.path.join(wd, "calls.log")
        env = dict(os.environ, FAKE_IIO_LOG=log, FAKE_IIO_EPOCHS="1", FAKE_IIO_NO_DEFAULT="0" if default_ctx else "1")
        args = [REF, "-e", NAV, "-l", "30.286502,120.032669,100"]
        for letter, arg in options:
            args += ["-" + letter, arg]
        subprocess.run(args, env=env, check=True, capture_output=True, cwd=wd)
        return open(log).read().splitlines()


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/ref_verbatim missing: make -C oracle (needs the reference tree)")
    golden = {name: {"options": opts, "default_context": dflt, "calls": reference_log(opts, dflt)}
              for name, (opts, dflt) in CASES.items()}
    golden["_banner_v"] = {"options": [["s", "2600000"]], "default_context": True, "calls": [], "stderr": reference_banner()}
    with open(OUT, "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote %s: %s" % (OUT, {k: len(v["calls"]) for k, v in golden.items() if not k.startswith("_")}))


if __name__ == "__main__":
    main()
