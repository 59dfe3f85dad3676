#!/usr/bin/env python3
"""Every epoch of a large batch against the oracle (GPU box): device checksum of each epoch == oracle, post-epoch
carrier phases == literal recurrence.  usage: full_parity.py [epochs] [batches] [mode: pipe|phase] [est_err]"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import oracle_lib as ol  # noqa: E402
from pluto_gps_sim_b200 import Synthesizer, capi, checksum_host  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "pipe"
est_err = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
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
err = torch.full((C,), est_err, dtype=torch.float64, device="cuda")
zero = torch.zeros(C, dtype=torch.float64, device="cuda")
if mode == "pipe":
    s.submit_device(d_first.data_ptr(), E, sp)
    for b in range(1, B):
        s.submit_device(d_desc.data_ptr(), E, sp)
        s.fetch_device(d_out.data_ptr(), sp)
    s.fetch_device(d_out.data_ptr(), sp)
else:
    for b in range(B):
        d = d_first if b == 0 else d_desc
        s.prepare_device(d.data_ptr(), E, None, sp)
        if est_err:
            s.estimate_correct_device(zero.data_ptr(), err.data_ptr(), 1.0, sp)   # est <- est - est_err
        s.speculate_device(d.data_ptr(), E, sp)
        s.chain_device(d.data_ptr(), E, sp)
        s.render_device(d.data_ptr(), E, d_out.data_ptr(), sp)
torch.cuda.synchronize()
trace = s.carrier_trace(E)
sums = s.checksum_device(d_out.data_ptr(), E)
print("fallbacks", s.carrier_fallbacks, "line stats", s.line_stats, flush=True)
t0 = time.time()
bad_chain, bad_sum = [], []
for e in range(1, E):
    for c in range(C):
        want = ol.oracle_carr_nco(float(trace[e - 1, c]), float(desc[e, c]["carr_step"]), N)
        if want != trace[e, c]:
            bad_chain.append((e, c))
print("chain: %d bad links of %d (%.1f s)" % (len(bad_chain), (E - 1) * C, time.time() - t0), bad_chain[:10], flush=True)
t0 = time.time()
for e in range(1, E):
    stt = trace[e - 1].copy()
    iq, tr = ol.oracle_synth(desc[e:e + 1], N, carr_state=stt)
    if int(checksum_host(iq[0])) != int(sums[e]):
        bad_sum.append(e)
print("samples: %d bad epochs of %d (%.1f s)" % (len(bad_sum), E - 1, time.time() - t0), bad_sum[:20], flush=True)
if bad_sum:
    e = bad_sum[0]
    got = d_out.view(E, N, 2)[e].cpu().numpy()
    stt = trace[e - 1].copy()
    iq, _ = ol.oracle_synth(desc[e:e + 1], N, carr_state=stt)
    idx = np.argwhere((got != iq[0]).any(axis=1)).reshape(-1)
    print("epoch", e, "differing samples", len(idx), idx[:10], "tiles", sorted(set((idx // 1024).tolist()))[:10])
    print("got", got[idx[:4]].tolist(), "want", iq[0][idx[:4]].tolist())
sys.exit(1 if (bad_chain or bad_sum) else 0)
