#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: regenerate tests/golden/* from the reference compiled in
oracle/_ref (needs /root/reference at build time; run `make -C oracle` first).

For each scenario it stores
  <name>_desc.npy   gpsiq descriptors [E][C] made from the reference's own
                    per-epoch channel state (through gpsiq_make_desc)
  <name>_meta.json  SHA-256 of the reference's full int16 stream, per-epoch
                    mix64 checksums, post-epoch carrier phases (hex), fixture hashes
  <name>_head.npy   first HEAD samples of epoch 0 and last HEAD of the final epoch
and, for static12, the raw reference state dump (static12_dump.npy) that the
host-orchestration parity tests compare against.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import refdump  # noqa: E402
from pluto_gps_sim_b200.synth import checksum_host  # noqa: E402

HEAD = 8192
PLAN = [("static12", 10, 0), ("circle12", 310, 0), ("allsky32", 20, 0), ("static12int", 10, 1), ("circle12int", 310, 1)]


def file_sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    out = refdump.GOLDEN
    only = sys.argv[1:]          # optional: regenerate just these scenarios
    for name, epochs, mode in PLAN:
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as wd:
            recs, iq, js = refdump.run_reference(name, epochs, wd)
        desc = refdump.to_descriptors(recs, mode)
        np.save(os.path.join(out, name + "_desc.npy"), desc)
        head = np.stack([iq[0, :HEAD], iq[-1, -HEAD:]])
        np.save(os.path.join(out, name + "_head.npy"), head)
        meta = {
            "scenario": name,
            "reference_cmd": refdump.SCENARIOS[name][0] + " -e " + refdump.SCENARIOS[name][1] + " " + " ".join(refdump.SCENARIOS[name][2]),
            "nav_fixture_sha256": file_sha(os.path.join(out, refdump.SCENARIOS[name][1])),
            "epochs": epochs,
            "max_chan": int(recs.shape[1]),
            "samples_per_epoch": int(iq.shape[1]),
            "iq_sha256": hashlib.sha256(iq.tobytes()).hexdigest(),
            "epoch_checksums": [int(checksum_host(iq[e])) for e in range(epochs)],
            "carr_phase_end_hex": [[float(x).hex() for x in recs[e]["carr_phase_end"]] for e in range(epochs)],
            "prn_first": [int(p) for p in recs[0]["prn"]],
            "prn_last": [int(p) for p in recs[-1]["prn"]],
            "iq_min": int(iq.min()), "iq_max": int(iq.max()),
            "harness_summary": json.loads(js),
        }
        with open(os.path.join(out, name + "_meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        if name == "static12":
            np.save(os.path.join(out, name + "_dump.npy"), recs)
        print(name, "epochs", epochs, "sha", meta["iq_sha256"][:16], "prn", meta["prn_first"])


if __name__ == "__main__":
    main()
