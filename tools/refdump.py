"""TEST INFRASTRUCTURE: read the per-(epoch, slot) state dump written by
oracle/_ref/ref_harness* (layout: ref_dump_t in oracle/ref_harness.c) and turn
it into gpsiq descriptors through the product's own host helper
(gpsiq_make_desc), so the conversion under test is the shipped one."""
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "oracle", "_ref")
GOLDEN = os.path.join(REPO, "tests", "golden")

DUMP_DTYPE = np.dtype(
    [
        ("epoch", "<i4"), ("slot", "<i4"), ("prn", "<i4"), ("iword", "<i4"), ("ibit", "<i4"), ("icode", "<i4"),
        ("dataBit", "<i4"), ("codeCA", "<i4"),
        ("f_carr", "<f8"), ("f_code", "<f8"), ("delt", "<f8"), ("carr_phase", "<f8"), ("code_phase", "<f8"),
        ("gain", "<f8"), ("carr_phase_end", "<f8"), ("azel", "<f8", (2,)),
        ("rho_range", "<f8"), ("rho_d", "<f8"), ("rho_iono", "<f8"), ("g0_sec", "<f8"),
        ("g0_week", "<i4"), ("pad", "<i4"), ("dwrd", "<u4", (60,)),
    ]
)
assert DUMP_DTYPE.itemsize == 384

SCENARIOS = {
    # name: (harness, nav fixture, extra args, max_chan)
    "static12": ("ref_harness_O2", "brdc3540_synth.14n.gz", ["-l", "30.286502,120.032669,100", "-s", "2600000"], 12),
    "circle12": ("ref_harness_O2", "brdc3540_synth.14n.gz", ["-u", "@circle", "-s", "2600000"], 12),
    # the reference built WITHOUT FLOAT_CARR_PHASE (plutogpssim.h:12 removed): its integer carrier NCO
    "static12int": ("ref_harness_int_O2", "brdc3540_synth.14n.gz", ["-l", "30.286502,120.032669,100", "-s", "2600000"], 12),
    "circle12int": ("ref_harness_int_O2", "brdc3540_synth.14n.gz", ["-u", "@circle", "-s", "2600000"], 12),
    "allsky32": ("ref_harness32_O2", "allsky32_synth.14n.gz", ["-l", "30.286502,120.032669,100", "-s", "10000000"], 32),
}


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "ref_harness_O2"))


def run_reference(scenario, epochs, workdir, want_iq=True, harness=None):
    """Run the compiled reference; returns (dump records [epochs][max_chan], iq int16 [epochs, N, 2] or None, json text)."""
    h, nav, extra, max_chan = SCENARIOS[scenario]
    h = harness or h
    extra = [os.path.join(REF_DIR, "circle.csv") if a == "@circle" else a for a in extra]
    env = dict(os.environ)
    env["REF_EPOCHS"] = str(epochs)
    desc_path = os.path.join(workdir, scenario + ".desc")
    env["REF_DESC_OUT"] = desc_path
    iq_path = os.path.join(workdir, scenario + ".iq")
    if want_iq:
        env["REF_IQ_OUT"] = iq_path
    r = subprocess.run([os.path.join(REF_DIR, h), "-e", os.path.join(GOLDEN, nav)] + extra, env=env,
                       capture_output=True, text=True, check=True)
    recs = np.fromfile(desc_path, DUMP_DTYPE).reshape(epochs, max_chan)
    iq = np.fromfile(iq_path, np.int16).reshape(epochs, -1, 2) if want_iq else None
    return recs, iq, r.stdout.strip().splitlines()[-1]


def to_descriptors(recs, carrier_mode=0):
    """ref dump -> gpsiq_chan_desc records, via gpsiq_make_desc (shipped host code)."""
    sys.path.insert(0, REPO)
    from pluto_gps_sim_b200 import capi

    E, Cn = recs.shape
    out = np.zeros((E, Cn), capi.DESC_DTYPE)
    prev = np.zeros(Cn, np.int64)
    for e in range(E):
        for c in range(Cn):
            r = recs[e, c]
            if r["prn"] <= 0:
                prev[c] = 0
                continue
            # a slot whose PRN changed was (re)allocated: its carr_phase is a fresh
            # initial value (plutogpssim.c:1964), not the previous loop's result
            fresh = prev[c] != r["prn"]
            out[e, c] = capi.make_desc(carrier_mode, r["prn"], r["f_carr"], r["f_code"], r["delt"], r["carr_phase"],
                                       r["code_phase"], r["dwrd"].astype(np.uint64), r["iword"], r["ibit"],
                                       r["icode"], r["gain"], fresh)
            prev[c] = r["prn"]
    return out
