#!/usr/bin/env python3
"""Record the measured DRAM traffic of k_synth_line from an `ncu --set full` report into profiles/traffic.json,
keyed by the SHA-256 of the kernel's source file, so that bench.py only quotes `roofline.traffic` for the kernel
revision that was actually captured.
usage: update_traffic.py <report.ncu-rep> <channels> <epochs per launch> <profiles/summary file it was summarised into>"""
import csv
import hashlib
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles", "traffic.json")
SRC = os.path.join(REPO, "pluto_gps_sim_b200", "csrc", "synth_line.cuh")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, channels, epochs, summary = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    row = [r for r in rows[2:] if "k_synth_line" in r[h.index("Kernel Name")]][0]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(k)
        tot += float(row[i]) * UNIT[u[i]]
    try:
        t = json.load(open(OUT))
    except Exception:
        t = {}
    t.setdefault("k_synth_line", {})["c%d" % channels] = {
        "dram_bytes_per_epoch": tot / epochs, "dram_bytes_per_launch": tot, "epochs_per_launch": epochs,
        "algorithmic_bytes_per_epoch": 300000 * 4, "source": summary,
        "kernel_src_sha256": hashlib.sha256(open(SRC, "rb").read()).hexdigest(),
    }
    json.dump(t, open(OUT, "w"), indent=1)
    print("traffic c%d: %.1f MB per launch of %d epochs (%.3fx algorithmic)" % (channels, tot / 1e6, epochs, tot / (epochs * 1.2e6)))


if __name__ == "__main__":
    main()
