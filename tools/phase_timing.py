#!/usr/bin/env python3
"""Time the scan and render phases of one batch separately (one stream, no pipelining), then the
pipelined streaming pair, with CUDA events.  Run on the GPU box.  usage: phase_timing.py [epochs]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pluto_gps_sim_b200 import Synthesizer, capi  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = 300000
base = np.load(os.path.join(REPO, "tests", "golden", "static12_desc.npy"))
C = base.shape[1]
desc = np.concatenate([base] * ((E + 9) // 10))[:E].copy()
desc["flags"] = 0
first = desc.copy()
first[0]["flags"] = capi.FLAG_RESET_CARRIER
d_first = torch.from_numpy(first.view(np.uint8).reshape(-1)).cuda()
d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(E * N * 2, dtype=torch.int16, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
sp = st.cuda_stream
s = Synthesizer(max_chan=C, max_epochs=E)


def ev():
    return torch.cuda.Event(enable_timing=True)


s.scan_device(d_first.data_ptr(), E, sp)
s.render_device(d_first.data_ptr(), E, d_out.data_ptr(), sp)
torch.cuda.synchronize()
ts, tr = [], []
for i in range(8):
    a, b, c = ev(), ev(), ev()
    a.record()
    s.scan_device(d_desc.data_ptr(), E, sp)
    b.record()
    s.render_device(d_desc.data_ptr(), E, d_out.data_ptr(), sp)
    c.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
    tr.append(b.elapsed_time(c))
print("sequential: scan %.3f ms  render %.3f ms  (min of 8: %.3f / %.3f)" % (np.median(ts), np.median(tr), min(ts), min(tr)))

# phases of the scan
tp, tsp, tc = [], [], []
for i in range(8):
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    s.prepare_device(d_desc.data_ptr(), E, None, sp)
    b.record()
    s.speculate_device(d_desc.data_ptr(), E, sp)
    c.record()
    s.chain_device(d_desc.data_ptr(), E, sp)
    d.record()
    s.render_device(d_desc.data_ptr(), E, d_out.data_ptr(), sp)
    torch.cuda.synchronize()
    tp.append(a.elapsed_time(b)); tsp.append(b.elapsed_time(c)); tc.append(c.elapsed_time(d))
print("scan phases: prepare %.3f  speculate %.3f  chain(+code scan join) %.3f ms" % (np.median(tp), np.median(tsp), np.median(tc)))


def run_batches(count):
    s.submit_device(d_desc.data_ptr(), E, sp)
    for _ in range(count - 1):
        s.submit_device(d_desc.data_ptr(), E, sp)
        s.fetch_device(d_out.data_ptr(), sp)
    s.fetch_device(d_out.data_ptr(), sp)


run_batches(3)
torch.cuda.synchronize()
if os.environ.get("GPSIQ_TRACE"):
    capi.lib.gpsiq_trace_dump(s._ctx, 1)
    run_batches(4)
    capi.lib.gpsiq_trace_dump(s._ctx, 1)
a, b = ev(), ev()
a.record()
run_batches(20)
b.record()
torch.cuda.synchronize()
print("pipelined: %.3f ms per batch of %d epochs -> %.3e samples/s" % (a.elapsed_time(b) / 20, E, 20 * E * N / (a.elapsed_time(b) * 1e-3)))
import time
t0 = time.perf_counter()
run_batches(20)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue time per batch: %.3f ms (then %.3f ms to drain)" % ((t1 - t0) / 20 * 1e3, (t2 - t1) * 1e3))
