"""Time-slice runner on ONE GPU (world 1): same stream structure as the multi-GPU run, no NCCL."""
import os, sys, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from pluto_gps_sim_b200 import Synthesizer, capi
from pluto_gps_sim_b200.timeslice import GpuSliceEngine, TimeSliceRunner
E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = np.load(os.path.join(REPO, "tests", "golden", "static12_desc.npy"))
desc = np.concatenate([base] * ((E + 9) // 10))[:E].copy(); desc["flags"] = 0
first = desc.copy(); first[0]["flags"] = 1
d_first = torch.from_numpy(first.view(np.uint8).reshape(-1)).cuda(); d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).cuda()
out = torch.empty(E * 300000 * 2, dtype=torch.int16, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
s = Synthesizer(max_chan=12, max_epochs=E)
r = TimeSliceRunner(GpuSliceEngine(s), 0, 1, deferred_render=True)
for i in range(3): r.step(d_first if i == 0 else d_desc, E, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(12): r.step(d_desc, E, out)
b.record(); r.finish(); torch.cuda.synchronize()
print("runner world=1: %.3f ms per step" % (a.elapsed_time(b) / 12))
s.close()
