#!/usr/bin/env python3
"""bench.py -- throughput of the GPS L1 C/A I/Q synthesis hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): Msamples/s of complex int16 I/Q, bit-exact vs the CPU
reference.  Workload at N=1: BASELINE config[1] -- static location, 2.6 MS/s
(delt = 1/2.6e6), 12 visible channels, 300 000 samples per 0.1 s epoch.  One
*step* = one pass of the hot path over a batch of EPOCHS consecutive epochs
(EPOCHS*300000 samples).  The descriptors are the committed reference-derived
golden descriptors of that scenario (tests/golden/static12_desc.npy, 10 epochs)
tiled in time to the batch length: same Doppler / code-phase / gain statistics,
carrier phase chaining through the whole run.

  value   whole-job Msamples/s with descriptors and output resident in HBM
          (CUDA events, max over ranks);
  e2e     the same through the public host-buffer call gpsiq_synth: pinned host
          descriptors -> H2D -> kernels -> D2H of the full int16 stream;
  roofline  the synthesis kernel against the measured HBM peak: algorithmic
          bytes = 4 B per complex sample written;
  cpu_baseline  the reference's own loop (oracle/_ref/ref_harness_O2, compiled from the
          reference source) timed on this box, one core -- its design point.

N > 1 (torchrun, one rank per GPU): the stream is time-sliced, rank r renders
slice r of every step and hands the carrier phases to rank r+1 over NCCL
(pluto_gps_sim_b200/timeslice.py); "weak" scaling: EPOCHS per rank per step is fixed.

--impl reference: the reference CPU implementation on all host cores (one
independent replica per core: one stream cannot be sliced on the CPU because
carr_phase chains across epochs), same metric/config; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_SAMPLES = 300000
GOLDEN_DESC = os.path.join(REPO, "tests", "golden", "static12_desc.npy")
NAV_FIXTURE = os.path.join(REPO, "tests", "golden", "brdc3540_synth.14n.gz")
REF_BIN = os.path.join(REPO, "oracle", "_ref", "ref_harness_O2")
REF_BIN_O0 = os.path.join(REPO, "oracle", "_ref", "ref_harness_O0")
REF_ARGS = ["-e", NAV_FIXTURE, "-l", "30.286502,120.032669,100", "-s", "2600000"]
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from ncu --set full captures, per EPOCH
# (300000 samples x 12 slots) so that it scales to the launch the roofline is quoted on:
#   k_synth_line   1255168128 B for a 1024-epoch launch (profiles/r01_w_synth_line_ncu_full.txt): 83.7 MB read
#                  (tables, anchors) + 1171.5 MB written of the 1228.8 MB of output (the rest is still in L2)
#   k_synth_fixed  37450240 B for an 8-epoch launch (profiles/r01_c_render_kernels_ncu_full.txt)
TRAFFIC_PER_EPOCH = {"k_synth_line": 1255168128 / 1024, "k_synth_fixed": 37450240 / 8}
WORKLOAD = "config[1]: static -l 30.286502,120.032669,100, synthetic brdc3540.14n, 2.6 MS/s, 12 channels, 300000 samples/epoch"


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        self.idx = gpu_index
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the busiest half of the samples = under load
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm
def run_reference_replicas(n_rep, epochs, binary=REF_BIN):
    """n_rep concurrent replicas of the compiled reference, `epochs` epochs each.
    Returns (sum of per-replica Msamples/s, list of loop seconds)."""
    env = dict(os.environ, REF_EPOCHS=str(epochs))
    env.pop("REF_IQ_OUT", None); env.pop("REF_DESC_OUT", None)
    procs = [subprocess.Popen([binary] + REF_ARGS, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(n_rep)]
    rates, secs = [], []
    for p in procs:
        out, _ = p.communicate()
        js = json.loads(out.strip().splitlines()[-1])
        rates.append(js["msamples_per_s"]); secs.append(js["loop_seconds"])
    return sum(rates), secs


def run_oracle_port(epochs):
    """Fallback CPU baseline when oracle/_ref is absent: the C restatement, one core."""
    import numpy as np
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    desc = np.load(GOLDEN_DESC)
    desc = np.concatenate([desc] * ((epochs + 9) // 10))[:epochs]
    t0 = time.perf_counter()
    ol.oracle_synth(desc, N_SAMPLES)
    dt = time.perf_counter() - t0
    return epochs * N_SAMPLES / dt / 1e6, [dt]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    have_ref = os.path.exists(REF_BIN)
    epochs = 20                                  # bounded sample per step: 6e6 samples per replica
    vals = []
    for i in range(args.warmup + args.steps):
        if have_ref:
            v, _ = run_reference_replicas(cores, epochs)
        else:
            v, _ = run_oracle_port(epochs)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    used = cores if have_ref else 1
    samples_per_step = epochs * N_SAMPLES * used
    line = {
        "impl": "reference",
        "metric": "Msamples/sec (complex I/Q) at 12 channels; bit-exact vs CPU ref",   # the same metric string as our arm
        "value": round(value, 3), "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * samples_per_step / (value * 1e6), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 NCO / int accumulate", "data": "synthetic (generated RINEX fixture)",
        "config": {"workload": WORKLOAD, "epochs_per_step_per_replica": epochs, "replicas": used,
                   "note": "independent replicas, one per host core: a single stream cannot be sliced on the CPU "
                           "(carr_phase chains across epochs); the reference itself is single-threaded by design"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": used,
                         "kind": "reference" if have_ref else "port",
                         "sample": "%d replicas x %d epochs x 300000 samples per step, gcc -O2 -march=x86-64-v3 build of "
                                   "the reference source" % (used, epochs)},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm
def ours_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from pluto_gps_sim_b200 import Synthesizer, capi
    from pluto_gps_sim_b200.timeslice import GpuSliceEngine, TimeSliceRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own "NCCL version ..." banner (printed when the environment
        # sets NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        opts = None
        if os.environ.get("GPSIQ_NCCL_HIPRI"):               # experiment: NCCL kernels on a high-priority stream
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)

    E = args.epochs
    if args.workload == "config3":
        base = np.load(os.path.join(REPO, "tests", "golden", "allsky32_desc.npy"))   # [20][32]
        args.no_cpu_baseline = True
    else:
        base = np.load(GOLDEN_DESC)                           # [10][12], reference-derived
    C = base.shape[1]
    desc = np.concatenate([base] * ((E + len(base) - 1) // len(base)))[:E].copy()
    desc["flags"] = 0                                         # carrier chains through the whole run ...
    first = desc.copy()
    first[0]["flags"] = capi.FLAG_RESET_CARRIER               # ... from the allocation phases of epoch 0
    samples_per_step = E * N_SAMPLES

    synth = Synthesizer(max_chan=C, max_epochs=E, device=local, kernel=args.kernel, tile_samples=args.tile)
    d_first = torch.from_numpy(first.view(np.uint8).reshape(-1)).cuda()
    d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(samples_per_step * 2, dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream()                 # everything below is enqueued on this stream
    torch.cuda.set_stream(stream)
    engine = GpuSliceEngine(synth)
    handoff = args.handoff if world > 1 else "nccl"
    if handoff == "mailbox" and not engine.mailbox_setup(rank, world):
        handoff = "nccl"                                     # decided collectively: every rank falls back alike
    runner = TimeSliceRunner(engine, rank, world, deferred_render=True, handoff=handoff)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    # N == 1: the streaming pair submit/fetch with one batch of lookahead (the serial carrier chain of
    # batch k+1 overlaps the sample kernels of batch k).  N > 1: time slices with the NCCL hand-off.
    sp = stream.cuda_stream

    def run_batches(count, first_desc=None):
        """`count` whole batches through the streaming pair, pipeline fill and drain included."""
        synth.submit_device((first_desc if first_desc is not None else d_desc).data_ptr(), E, sp)
        for _ in range(count - 1):
            synth.submit_device(d_desc.data_ptr(), E, sp)    # scan of the next batch ...
            synth.fetch_device(d_out.data_ptr(), sp)         # ... while this one is rendered
        synth.fetch_device(d_out.data_ptr(), sp)

    if world == 1:
        run_batches(max(args.warmup, 1), d_first)
    else:
        for i in range(args.warmup):
            runner.step(d_first if (i == 0 and rank == 0) else d_desc, E, d_out)
    barrier()
    l0 = synth.launch_count
    synth.timing_begin()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if world == 1:
        run_batches(args.steps)                              # exactly `steps` batches scanned AND rendered in here
    else:
        for i in range(args.steps):
            runner.step(d_desc, E, d_out)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        runner.finish()
    torch.cuda.synchronize()
    launches = synth.launch_count - l0
    nrec, scan_ms, synth_ms = synth.timing_collect()
    kn, kms, kep = synth.timing_sample_kernel()
    try:
        kiso_ms, kiso_ep = synth.timing_sample_kernel_isolated(20)
    except Exception:
        kiso_ms, kiso_ep = 0.0, 0
    fallbacks = synth.carrier_fallbacks
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, launches = float(tm[0]), int(ts[1])
    value = world * args.steps * samples_per_step / (ms * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call (e2e)
    synth2 = Synthesizer(max_chan=C, max_epochs=E, device=local, kernel=args.kernel, tile_samples=args.tile)
    nbytes_desc = E * C * 64
    h_desc = capi.lib.gpsiq_host_alloc(nbytes_desc)
    h_iq = capi.lib.gpsiq_host_alloc(samples_per_step * 4)
    assert h_desc and h_iq
    import ctypes
    ctypes.memmove(h_desc, first.ctypes.data, nbytes_desc)
    h_iq2 = capi.lib.gpsiq_host_alloc(samples_per_step * 4)
    assert h_iq2
    outs = [h_iq, h_iq2]

    def run_host_batches(count, first=None):
        """`count` batches through the host-buffer streaming pair gpsiq_submit / gpsiq_fetch: every batch's
        descriptors go host->device and its full int16 stream comes device->host inside this call sequence."""
        if first is not None:
            ctypes.memmove(h_desc, first.ctypes.data, nbytes_desc)
        synth2.submit_ptr(h_desc, E)
        ctypes.memmove(h_desc, desc.ctypes.data, nbytes_desc)
        for k in range(count - 1):
            synth2.submit_ptr(h_desc, E)
            synth2.fetch_ptr(outs[k & 1])                    # blocking: the host buffer is complete on return
        synth2.fetch_ptr(outs[(count - 1) & 1])

    run_host_batches(max(args.warmup, 1), first)             # warm-up (also seeds the carrier from epoch 0)
    barrier()
    t0 = time.perf_counter()
    run_host_batches(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * samples_per_step / float(te[0]) / 1e6
    capi.lib.gpsiq_host_free(h_desc); capi.lib.gpsiq_host_free(h_iq); capi.lib.gpsiq_host_free(h_iq2)

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel = k_synth_line (k_synth_fixed with --kernel 2); one launch covers `kep` epochs
        dom = "k_synth_fixed" if args.kernel == capi.KERNEL_FIXED_POINT else "k_synth_line"
        if kiso_ms > 0:
            # the dominant kernel timed ALONE (20 back-to-back re-launches of the last k_synth_fixed on an idle
            # device, CUDA events on its stream); inside the pipelined region it shares the SMs with the scan
            # kernels of the next batch, so its in-pipeline duration (kernel_ms_in_pipeline) is not the kernel's own
            kern_ms = kiso_ms
            kern_bytes = kiso_ep * N_SAMPLES * 4
            kern_name = "%s (one launch = %d epochs), timed alone" % (dom, kiso_ep)
        elif kn > 0:
            kern_ms = kms / kn
            kern_bytes = kep * N_SAMPLES * 4
            kern_name = "%s (one launch = %d epochs)" % (dom, kep)
        else:
            kern_ms = synth_ms / max(nrec, 1)
            kern_bytes = samples_per_step * 4
            kern_name = "k_synth_lanes"
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if os.path.exists(REF_BIN):
                v, secs = run_reference_replicas(1, 200)       # ~5 s of one core
                cpu = {"value": round(v, 3), "unit": "Msamples/s", "cores": 1, "kind": "reference",
                       "sample": "200 epochs (6e7 samples) of the same scenario, reference source built -O2 "
                                 "-march=x86-64-v3, single thread as designed"}
                if os.path.exists(REF_BIN_O0):
                    v0, _ = run_reference_replicas(1, 40, REF_BIN_O0)
                    cpu["shipped_flags_O0_value"] = round(v0, 3)
            else:
                v, secs = run_oracle_port(100)
                cpu = {"value": round(v, 3), "unit": "Msamples/s", "cores": 1, "kind": "port",
                       "sample": "100 epochs (3e7 samples), oracle C restatement"}
        line = {
            "metric": "Msamples/sec (complex I/Q) at %d channels; bit-exact vs CPU ref" % C,
            "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 NCO / int32 accumulate / int16 out",
            "data": "synthetic (reference-derived golden descriptors of config[1], tiled in time)",
            "config": {"workload": WORKLOAD if args.workload == "config1" else
                       "config[3]: static location, 10.0 MS/s, 32 channels (synthetic all-visible constellation), 300000 samples/epoch",
                       "epochs_per_step_per_gpu": E, "samples_per_step_per_gpu": samples_per_step,
                       "parallelism": ("time-slice x%d, %s carrier-phase hand-off" % (world, "NCCL" if handoff == "nccl" else
                                        "peer-memory mailbox (NCCL only for the advance all_gather)")) if world > 1 else "single GPU",
                       "l2_policy": "output per step %.1f MB > 126 MB L2; inputs are %d B of descriptors"
                                    % (samples_per_step * 4 / 1e6, nbytes_desc),
                       "kernel": args.kernel, "tile_samples": args.tile,
                       "carrier_scan_serial_fallbacks": fallbacks,
                       "carrier_scan_chains": (args.warmup + args.steps) * E * C},
            "e2e": {"value": round(e2e_value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": nbytes_desc,
                    "d2h_bytes_per_step": samples_per_step * 4},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": int(TRAFFIC_PER_EPOCH[dom] * (kern_bytes // (N_SAMPLES * 4))) if C == 12 else None, "peak_source": peak_src,
                         "kernel": kern_name, "kernel_ms_per_launch": round(kern_ms, 4),
                         "algorithmic_bytes_per_launch": kern_bytes,
                         "kernel_ms_in_pipeline": round(kms / kn, 4) if kn else None,
                         "step_level_achieved_gbs": round(value * 4 / 1e3, 2),
                         "scan_phase_ms_per_step": round(scan_ms / max(nrec, 1), 4),
                         "render_phase_ms_per_step": round(synth_ms / max(nrec, 1), 4)},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--epochs", type=int, default=1024, help="epochs per step per GPU (1024 -> 1228.8 MB of output)")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--handoff", choices=["nccl", "mailbox"], default=os.environ.get("GPSIQ_HANDOFF", "mailbox"),
                    help="carrier-phase hand-off between time slices (N > 1): the SM-free peer-memory mailbox "
                         "(copy engine + stream memory operations; falls back to NCCL if unavailable), or NCCL send/recv")
    ap.add_argument("--workload", choices=["config1", "config3"], default="config1",
                    help="config1 (default, the metric's configuration): 12 channels, 2.6 MS/s; config3: 32 channels, "
                         "10 MS/s, synthetic all-visible constellation (informative; no CPU baseline / reference arm)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
