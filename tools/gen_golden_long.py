#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: BASELINE config[2] in full -- the reference's own 300 s user-motion stream (3000 epochs,
circle.csv, 2.6 MS/s): SHA-256 of the 3.6 GB it pushes and a per-epoch checksum list, without keeping the samples.
Needs oracle/_ref (make -C oracle).  Writes tests/golden/circle12_300s_meta.json."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import refdump  # noqa: E402
from pluto_gps_sim_b200.synth import checksum_host  # noqa: E402

EPOCHS = 3000
N = 300000


def main():
    h, nav, extra, max_chan = refdump.SCENARIOS["circle12"]
    extra = [os.path.join(refdump.REF_DIR, "circle.csv") if a == "@circle" else a for a in extra]
    with tempfile.TemporaryDirectory(dir="/tmp") as wd:
        iq_path = os.path.join(wd, "iq.bin")
        env = dict(os.environ, REF_EPOCHS=str(EPOCHS), REF_IQ_OUT=iq_path, REF_DESC_OUT=os.path.join(wd, "d.bin"))
        subprocess.run([os.path.join(refdump.REF_DIR, h), "-e", os.path.join(refdump.GOLDEN, nav)] + extra, env=env,
                       check=True, capture_output=True)
        sha = hashlib.sha256()
        sums = []
        with open(iq_path, "rb") as f:
            for e in range(EPOCHS):
                buf = f.read(N * 4)
                assert len(buf) == N * 4
                sha.update(buf)
                sums.append(int(checksum_host(np.frombuffer(buf, np.int16).reshape(N, 2))))
        recs = np.fromfile(os.path.join(wd, "d.bin"), refdump.DUMP_DTYPE).reshape(EPOCHS, max_chan)
    meta = {"scenario": "circle12 (BASELINE config[2], full length)", "epochs": EPOCHS, "samples_per_epoch": N,
            "reference_cmd": h + " -e " + nav + " " + " ".join(refdump.SCENARIOS["circle12"][2]),
            "iq_sha256": sha.hexdigest(), "epoch_checksums": sums,
            "prn_last": [int(p) for p in recs[-1]["prn"]],
            "carr_phase_end_hex_last": [float(x).hex() for x in recs[-1]["carr_phase_end"]]}
    with open(os.path.join(refdump.GOLDEN, "circle12_300s_meta.json"), "w") as f:
        json.dump(meta, f)
    print("circle12 300 s:", meta["iq_sha256"])


if __name__ == "__main__":
    main()
