#!/usr/bin/env python3
"""bench.py -- throughput of the GPS L1 C/A I/Q synthesis hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config1|config3] [--carrier float|int32]

Metric (BASELINE.json): Msamples/s of complex int16 I/Q, bit-exact vs the CPU
reference.  Workload at N=1: BASELINE config[1] -- static location, 2.6 MS/s
(delt = 1/2.6e6), 12 visible channels, 300 000 samples per 0.1 s epoch
(--workload config3: BASELINE config[3]/[4], 32 channels at 10 MS/s).  One
*step* = one pass of the hot path over a batch of EPOCHS consecutive epochs
(EPOCHS*300000 samples).  The descriptors are the committed reference-derived
golden descriptors of that scenario (tests/golden/*_desc.npy) tiled in time to
the batch length: same Doppler / code-phase / gain statistics, carrier phase
chaining through the whole run.

  value   whole-job Msamples/s with descriptors and output resident in HBM
          (CUDA events, max over ranks);
  e2e     the same through host buffers: pinned host descriptors -> H2D ->
          kernels -> D2H of the full int16 stream into pinned host memory
          (N = 1: gpsiq_submit / gpsiq_fetch; N > 1: the SAME time-sliced stream
          as `value`, every rank's slice rendered by gpsiq_fetch);
  parity  after the timed region every rank checks what it just rendered
          against the oracle (oracle/liboracle.so): a stride of epochs of the
          last batch bit for bit (device checksum == oracle checksum, end phase
          == oracle end phase, starting from the GPU's own carrier trace), a run
          of consecutive epochs of the carrier chain against the literal
          recurrence, and at N > 1 the slice boundary (rank r's first epoch from
          rank r-1's end phases).  A mismatch exits non-zero;
  roofline  the synthesis kernel against the measured HBM peak: algorithmic
          bytes = 4 B per complex sample written;
  cpu_baseline  the reference's own loop (oracle/_ref/ref_harness*, compiled from
          the reference source) timed on this box, one core -- its design point.

N > 1 (torchrun, one rank per GPU): the stream is time-sliced, rank r renders
slice r of every step (pluto_gps_sim_b200/timeslice.py); float carrier: the
phases go from rank to rank through the SM-free mailbox (NCCL fallback);
integer carrier: closed-form prefix, no ring.  "weak" scaling.

--impl reference: the reference CPU implementation on all host cores (one
independent replica per core: one stream cannot be sliced on the CPU because
carr_phase chains across epochs), same metric/config; rank 0 only.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# The pipeline keeps a dozen streams busy per GPU (render, per-set scan streams, ring, side streams, copies); with
# the default of 8 hardware work queues several streams share one, and a stream blocked in a wait (the ring's
# stream-memory wait on the next hand-off, an event wait) holds up unrelated work queued behind it.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_SAMPLES = 300000
GOLDEN = os.path.join(REPO, "tests", "golden")
REF_DIR = os.path.join(REPO, "oracle", "_ref")
LLH = "30.286502,120.032669,100"
# workload -> (description, golden descriptors per carrier mode, reference harness per carrier mode, reference args)
WORKLOADS = {
    "config1": {
        "text": "config[1]: static -l 30.286502,120.032669,100, synthetic brdc3540.14n, 2.6 MS/s, 12 channels, 300000 samples/epoch",
        "desc": {"float": "static12_desc.npy", "int32": "static12int_desc.npy"},
        "ref": {"float": "ref_harness_O2", "int32": "ref_harness_int_O2"},
        "ref_args": ["-e", os.path.join(GOLDEN, "brdc3540_synth.14n.gz"), "-l", LLH, "-s", "2600000"],
    },
    "config3": {
        "text": "config[3]: static location, 10.0 MS/s, 32 channels (synthetic all-visible constellation), 300000 samples/epoch",
        "desc": {"float": "allsky32_desc.npy"},
        "ref": {"float": "ref_harness32_O2"},
        "ref_args": ["-e", os.path.join(GOLDEN, "allsky32_synth.14n.gz"), "-l", LLH, "-s", "10000000"],
    },
}
REF_BIN_O0 = os.path.join(REF_DIR, "ref_harness_O0")
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu --set full summaries,
# keyed by the hash of the kernel's source at capture time (tools/update_traffic.py writes it)
TRAFFIC_JSON = os.path.join(REPO, "profiles", "traffic.json")
KERNEL_SRC = os.path.join(REPO, "pluto_gps_sim_b200", "csrc", "synth_line.cuh")


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_for(channels, epochs):
    """Measured DRAM traffic of one k_synth_line launch of `epochs` epochs -- only if the committed capture is of
    the kernel source that is being timed (else None: a stale number is worse than none)."""
    try:
        with open(TRAFFIC_JSON) as f:
            t = json.load(f)
        rec = t["k_synth_line"]["c%d" % channels]
        sha = hashlib.sha256(open(KERNEL_SRC, "rb").read()).hexdigest()
        if rec["kernel_src_sha256"] != sha:
            return None, "stale: %s was captured from another revision of synth_line.cuh" % rec["source"]
        return int(rec["dram_bytes_per_epoch"] * epochs), rec["source"]
    except Exception as e:
        return None, "no capture (%s)" % type(e).__name__


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


def bind_to_gpu_numa(local):
    """Pin this rank (and so its pinned host buffers, first touch) to the CPUs NVML reports as local to its GPU."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# --------------------------------------------------------------------------- reference arm
def run_reference_replicas(n_rep, epochs, binary, ref_args):
    """n_rep concurrent replicas of the compiled reference, `epochs` epochs each.
    Returns (sum of per-replica Msamples/s, list of loop seconds)."""
    env = dict(os.environ, REF_EPOCHS=str(epochs))
    env.pop("REF_IQ_OUT", None); env.pop("REF_DESC_OUT", None)
    procs = [subprocess.Popen([binary] + ref_args, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(n_rep)]
    rates, secs = [], []
    for p in procs:
        out, _ = p.communicate()
        js = json.loads(out.strip().splitlines()[-1])
        rates.append(js["msamples_per_s"]); secs.append(js["loop_seconds"])
    return sum(rates), secs


def run_oracle_port(epochs, desc_file, carrier_mode):
    """Fallback CPU baseline when oracle/_ref is absent: the C restatement, one core."""
    import numpy as np
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    desc = np.load(os.path.join(GOLDEN, desc_file))
    desc = np.concatenate([desc] * ((epochs + len(desc) - 1) // len(desc)))[:epochs]
    t0 = time.perf_counter()
    ol.oracle_synth(desc, N_SAMPLES, carrier_mode=carrier_mode)
    dt = time.perf_counter() - t0
    return epochs * N_SAMPLES / dt / 1e6, [dt]


def metric_name(channels):
    return "Msamples/sec (complex I/Q) at %d channels; bit-exact vs CPU ref" % channels


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    channels = 12 if args.workload == "config1" else 32
    cores = os.cpu_count() or 1
    ref_bin = os.path.join(REF_DIR, wl["ref"][args.carrier])
    have_ref = os.path.exists(ref_bin)
    epochs = 20 if channels == 12 else 6         # bounded sample per step and replica (6e6 / 1.8e6 samples)
    vals = []
    for i in range(args.warmup + args.steps):
        if have_ref:
            v, _ = run_reference_replicas(cores, epochs, ref_bin, wl["ref_args"])
        else:
            v, _ = run_oracle_port(epochs, wl["desc"][args.carrier], 1 if args.carrier == "int32" else 0)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    used = cores if have_ref else 1
    samples_per_step = epochs * N_SAMPLES * used
    line = {
        "impl": "reference",
        "metric": metric_name(channels),
        "value": round(value, 3), "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * samples_per_step / (value * 1e6), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 NCO / int accumulate" if args.carrier == "float" else "f64 code NCO, u32 carrier NCO / int accumulate",
        "data": "synthetic (generated RINEX fixture)",
        "config": {"workload": wl["text"], "epochs_per_step_per_replica": epochs, "replicas": used, "carrier": args.carrier,
                   "note": "independent replicas, one per host core: a single stream cannot be sliced on the CPU "
                           "(carr_phase chains across epochs); the reference itself is single-threaded by design"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": used,
                         "kind": "reference" if have_ref else "port",
                         "sample": "%d replicas x %d epochs x 300000 samples per step, gcc -O2 -march=x86-64-v3 build of "
                                   "the reference source (%s)" % (used, epochs, os.path.basename(ref_bin))},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- parity (inside the measured run)
def parity_check(synth, d_out, desc, E, C, carrier_mode, prev_end, stride, chain_n):
    """What this rank rendered last, against the oracle.  desc: the host descriptors of that batch ([E][C], no
    re-seed flags); prev_end: exact phases before its first epoch if known (slice boundary), else None.
    -> dict(checked_epochs, chain_epochs, boundary_checked, ok, first_bad)."""
    import numpy as np
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    from pluto_gps_sim_b200 import checksum_host

    trace = synth.carrier_trace(E)                                   # the GPU's post-epoch phases of the batch
    sums = synth.checksum_device(d_out.data_ptr(), E)                # device checksum of every epoch it rendered
    bad = None
    checked = 0

    why = []

    def check_epoch(e, start):
        st = np.array(start, dtype=np.float64)
        iq, tr = ol.oracle_synth(desc[e:e + 1], N_SAMPLES, carrier_mode=carrier_mode, carr_state=st)
        act = desc[e]["prn"] > 0
        ok_s, ok_p = int(checksum_host(iq[0])) == int(sums[e]), np.array_equal(tr[0][act], trace[e][act])
        if not ok_s:
            why.append("samples of epoch %d" % e)
        if not ok_p:
            why.append("end phases of epoch %d" % e)
        return ok_s and ok_p

    boundary = 0
    if prev_end is not None:
        boundary = 1
        if not check_epoch(0, prev_end):
            bad = 0
    for e in range(1, E, stride):
        if bad is not None:
            break
        checked += 1
        if not check_epoch(e, trace[e - 1]):
            bad = e
    # the chain itself: consecutive epochs, literal recurrence per slot (plutogpssim.c:2741-2748)
    chained = 0
    e0 = max(1, E // 2 - chain_n // 2)
    for e in range(e0, min(E, e0 + chain_n)):
        if bad is not None:
            break
        chained += 1
        for c in range(C):
            d = desc[e, c]
            if d["prn"] <= 0:
                continue
            if carrier_mode == 0:
                want = ol.oracle_carr_nco(float(trace[e - 1, c]), float(d["carr_step"]), N_SAMPLES)
            else:
                want = float((int(trace[e - 1, c]) + int(d["carr_step"]) * N_SAMPLES) % 2 ** 32)
            if want != trace[e, c]:
                bad = e
                why.append("carrier chain link into epoch %d slot %d" % (e, c))
                break
    return {"checked_epochs": checked + boundary, "chain_epochs": chained, "boundary_checked": boundary,
            "ok": bad is None, "first_bad": bad, "why": why, "end_phase": trace[E - 1].copy()}


# --------------------------------------------------------------------------- our arm
def ours_arm(args):
    import ctypes

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
    numa_cpus = bind_to_gpu_numa(local) if (world > 1 and not os.environ.get("GPSIQ_NO_NUMA_BIND")) else 0
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own "NCCL version ..." banner (printed when the environment
        # sets NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = WORKLOADS[args.workload]
    if args.carrier not in wl["desc"]:
        raise SystemExit("bench.py: no %s-carrier golden descriptors for %s" % (args.carrier, args.workload))
    carrier_mode = capi.CARRIER_INT32 if args.carrier == "int32" else capi.CARRIER_FLOAT
    E = args.epochs
    base = np.load(os.path.join(GOLDEN, wl["desc"][args.carrier]))          # [10][12] / [20][32], reference-derived
    C = base.shape[1]
    desc = np.concatenate([base] * ((E + len(base) - 1) // len(base)))[:E].copy()
    desc["flags"] = 0                                         # carrier chains through the whole run ...
    first = desc.copy()
    first[0]["flags"] = capi.FLAG_RESET_CARRIER               # ... from the allocation phases of epoch 0
    samples_per_step = E * N_SAMPLES
    nbytes_desc = E * C * 64

    def make_synth():
        return Synthesizer(max_chan=C, max_epochs=E, device=local, kernel=args.kernel, tile_samples=args.tile,
                           carrier_mode=carrier_mode)

    def use_pipelined(handoff):
        if handoff != "mailbox" or args.lockstep:
            return False
        return True if args.pipelined else world <= 4

    def make_runner(synth):
        engine = GpuSliceEngine(synth)
        if carrier_mode == capi.CARRIER_INT32:
            handoff = "prefix"
        else:
            handoff = args.handoff if world > 1 else "nccl"
            if handoff == "mailbox" and not engine.mailbox_setup(rank, world):
                handoff = "nccl"                              # decided collectively: every rank falls back alike
        # the mailbox hand-off never blocks the host, so the ranks need not run in lockstep: pipelined runner (rank 0
        # speculates from an estimate too; the next slice is prepared and its advances all-gathered one step ahead).
        # Measured on this pool's 8-GPU box (profiles/r02_v/x): 2.07 / 2.35 ms per step at 2 / 4 GPUs against 2.23 /
        # 2.44 for the lockstep runner, but 3.71 against 3.53 at 8 (every rank's estimate then rests on seven closed-form
        # advances, a slice speculated from a poor one chains serially on a ring that has no slack): pipelined up to
        # 4 GPUs, lockstep beyond, unless forced either way.
        return TimeSliceRunner(engine, rank, world, deferred_render=True, handoff=handoff,
                               pipelined=use_pipelined(handoff)), handoff

    synth = make_synth()
    d_first = torch.from_numpy(first.view(np.uint8).reshape(-1)).cuda()
    d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(samples_per_step * 2, dtype=torch.int16, device="cuda")
    stream = torch.cuda.Stream(priority=int(os.environ.get("GPSIQ_RENDER_PRIO", "0")))   # everything below is enqueued on this stream
    torch.cuda.set_stream(stream)
    runner, handoff = make_runner(synth) if world > 1 else (None, "none")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    # N == 1: the streaming pair submit/fetch with batches of lookahead (the carrier chain of the batches ahead
    # overlaps the sample kernels of this one).  N > 1: time slices.
    sp = stream.cuda_stream
    ahead = max(1, min(args.lookahead, capi.MAX_LOOKAHEAD))

    def run_batches(count, first_desc=None):
        """`count` whole batches through the streaming pair, pipeline fill and drain included."""
        sub = 0
        for i in range(min(ahead, count)):
            synth.submit_device((first_desc if (first_desc is not None and i == 0) else d_desc).data_ptr(), E, sp)
            sub += 1
        for _ in range(count):
            if sub < count:
                synth.submit_device(d_desc.data_ptr(), E, sp)    # scan of a later batch ...
                sub += 1
            synth.fetch_device(d_out.data_ptr(), sp)             # ... while this one is rendered

    if world == 1:
        run_batches(max(args.warmup, 1), d_first)
    else:
        for i in range(args.warmup):
            runner.step(d_first if (i == 0 and rank == 0) else d_desc, E, d_out, next_desc=d_desc)
    barrier()
    l0 = synth.launch_count
    fb0, sl0 = synth.carrier_fallbacks, synth.slice_stats   # (the device is idle here: reading the counters costs nothing)
    synth.timing_begin()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if world == 1:
        run_batches(args.steps)                              # exactly `steps` batches scanned AND rendered in here
    else:
        for i in range(args.steps):                          # (every step prepares the following one: one prepare per step)
            runner.step(d_desc, E, d_out, next_desc=d_desc)
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
    fallbacks = synth.carrier_fallbacks
    synth.check_device()                                     # a kernel-flagged error (incl. the chain's self-check) ends the run
    sl1 = synth.slice_stats
    sl = torch.tensor([float(sl1[0]), float(sl1[1]), float(sl1[0] - sl0[0]), float(sl1[1] - sl0[1]), float(fallbacks),
                       float(fallbacks - fb0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sl, op=dist.ReduceOp.SUM)
    slice_translated, slice_serial, slice_translated_timed, slice_serial_timed, fallbacks, fallbacks_timed = (int(v) for v in sl)
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, launches = float(tm[0]), int(ts[1])
    value = world * args.steps * samples_per_step / (ms * 1e-3) / 1e6

    # ---- parity of what was just rendered (the last batch / slice of the timed region is in d_out)
    par = None
    if not args.no_parity:
        t0 = time.perf_counter()
        stride = args.parity_stride or (16 if C <= 12 else 48)
        # slice boundary: rank r's first epoch starts from rank r-1's end phases of the same step
        prev_end = None
        if world > 1:
            mine = torch.from_numpy(synth.carrier_trace(E)[E - 1].copy()).cuda()
            ends = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(ends, mine)
            if rank > 0:
                prev_end = ends[rank - 1].cpu().numpy()
        par = parity_check(synth, d_out, desc, E, C, carrier_mode, prev_end, stride, 48 if C <= 12 else 16)
        par.pop("end_phase")
        par["seconds"] = round(time.perf_counter() - t0, 2)
        flag = torch.tensor([1.0 if par["ok"] else 0.0, float(par["checked_epochs"]), float(par["chain_epochs"]),
                             float(par["boundary_checked"])], dtype=torch.float64, device="cuda")
        if world > 1:
            mn = flag.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            sm = flag.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            par.update(ok=bool(mn[0] > 0), checked_epochs=int(sm[1]), chain_epochs=int(sm[2]), boundary_checked=int(sm[3]))
        par["what"] = ("last batch of the timed region, per rank: every %dth epoch rendered on the GPU == oracle (device "
                       "checksum of the int16 stream and end-of-epoch carrier phases, bit for bit), %s consecutive epochs of "
                       "the carrier chain == literal recurrence%s" % (stride, "48" if C <= 12 else "16",
                       ", each rank's first epoch from the previous rank's end phases" if world > 1 else ""))

    # the dominant kernel alone (after the parity check: the re-launches rewrite the last batch's samples)
    try:
        kiso_ms, kiso_ep = synth.timing_sample_kernel_isolated(20)
    except Exception:
        kiso_ms, kiso_ep = 0.0, 0

    # ---- end to end through host buffers (e2e)
    h_desc = capi.lib.gpsiq_host_alloc(nbytes_desc)
    outs = [capi.lib.gpsiq_host_alloc(samples_per_step * 4) for _ in range(2)]
    assert h_desc and all(outs)
    synth2 = make_synth() if not args.no_e2e else None
    e2e_s, e2e_how = float("inf"), "skipped (--no-e2e)"
    if args.no_e2e:
        pass
    elif world == 1:
        def run_host_batches(count, first_d=None):
            """`count` batches through the host-buffer streaming pair gpsiq_submit / gpsiq_fetch: every batch's
            descriptors go host->device and its full int16 stream comes device->host inside this call sequence."""
            if first_d is not None:
                ctypes.memmove(h_desc, first_d.ctypes.data, nbytes_desc)
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
        e2e_how = "gpsiq_submit / gpsiq_fetch, pinned host buffers"
    else:
        # the SAME time-sliced stream as `value`, host buffers: each rank uploads its slice's descriptors from pinned
        # host memory (on its scan stream) and renders the slice with gpsiq_fetch into pinned host memory
        runner2, _ = make_runner(synth2)
        hd = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(h_desc, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes_desc,)))
        for i in range(max(args.warmup, 1)):
            ctypes.memmove(h_desc, (first if (i == 0 and rank == 0) else desc).ctypes.data, nbytes_desc)
            runner2.step(hd, E, outs[i & 1], next_desc=None if (i == 0 and rank == 0) else hd)
            torch.cuda.synchronize()                             # (the upload of the staging buffer is done)
        ctypes.memmove(h_desc, desc.ctypes.data, nbytes_desc)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):                              # each step: scan slice k (+ H2D), render slice k-1 (+ D2H, blocking)
            runner2.step(hd, E, outs[i & 1], next_desc=hd)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        runner2.finish()
        torch.cuda.synchronize()
        e2e_how = "time slices: descriptors H2D from pinned memory per slice, slices rendered by gpsiq_fetch into pinned host buffers"
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * samples_per_step / float(te[0]) / 1e6
    capi.lib.gpsiq_host_free(h_desc)
    for o in outs:
        capi.lib.gpsiq_host_free(o)

    if rank == 0:
        peak, peak_src = peaks()
        if kiso_ms > 0:
            # the dominant kernel timed ALONE (20 back-to-back re-launches of the last k_synth_line on an idle
            # device, CUDA events on its stream); inside the pipelined region it shares the SMs with the scan
            # kernels of the next batch, so its in-pipeline duration (kernel_ms_in_pipeline) is not the kernel's own
            kern_ms = kiso_ms
            kern_ep = kiso_ep
            kern_name = "k_synth_line (one launch = %d epochs), timed alone" % kiso_ep
        elif kn > 0:
            kern_ms = kms / kn
            kern_ep = kep
            kern_name = "k_synth_line (one launch = %d epochs)" % kep
        else:
            kern_ms = synth_ms / max(nrec, 1)
            kern_ep = E
            kern_name = "k_synth_lanes"
        kern_bytes = kern_ep * N_SAMPLES * 4
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        traffic, traffic_src = traffic_for(C, kern_ep) if kern_name.startswith("k_synth_line") else (None, "n/a")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref_bin = os.path.join(REF_DIR, wl["ref"][args.carrier])
            if os.path.exists(ref_bin):
                n_ep = 200 if C <= 12 else 60                    # ~5-10 s of one core
                v, secs = run_reference_replicas(1, n_ep, ref_bin, wl["ref_args"])
                cpu = {"value": round(v, 3), "unit": "Msamples/s", "cores": 1, "kind": "reference",
                       "sample": "%d epochs (%.1e samples) of the same scenario, reference source built -O2 "
                                 "-march=x86-64-v3 (%s), single thread as designed" % (n_ep, n_ep * N_SAMPLES, wl["ref"][args.carrier])}
                if args.workload == "config1" and args.carrier == "float" and os.path.exists(REF_BIN_O0):
                    v0, _ = run_reference_replicas(1, 40, REF_BIN_O0, wl["ref_args"])
                    cpu["shipped_flags_O0_value"] = round(v0, 3)
            else:
                v, secs = run_oracle_port(100 if C <= 12 else 30, wl["desc"][args.carrier], carrier_mode)
                cpu = {"value": round(v, 3), "unit": "Msamples/s", "cores": 1, "kind": "port",
                       "sample": "oracle C restatement, one core"}
        par_text = {"none": "single GPU", "prefix": "time-slice x%d, integer carrier: closed-form prefix of the slices' advances "
                    "(one all_gather per step, no ring)" % world,
                    "nccl": "time-slice x%d, NCCL carrier-phase hand-off" % world,
                    "mailbox": "time-slice x%d, peer-memory mailbox carrier-phase hand-off (NCCL only for the advance all_gather)" % world}
        line = {
            "metric": metric_name(C),
            "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 NCO / int32 accumulate / int16 out" if carrier_mode == capi.CARRIER_FLOAT else
                     "f64 code NCO, u32 carrier NCO / int32 accumulate / int16 out",
            "data": "synthetic (reference-derived golden descriptors of %s, tiled in time)" % args.workload,
            "config": {"workload": wl["text"], "carrier": args.carrier,
                       "epochs_per_step_per_gpu": E, "samples_per_step_per_gpu": samples_per_step,
                       "parallelism": par_text[handoff if world > 1 else "none"],
                       "l2_policy": "output per step %.1f MB > 126 MB L2; inputs are %d B of descriptors"
                                    % (samples_per_step * 4 / 1e6, nbytes_desc),
                       "kernel": args.kernel, "tile_samples": args.tile, "lookahead_batches": ahead if world == 1 else 1,
                       "carrier_scan_serial_fallbacks": fallbacks,
                       "slice_chains_translated": slice_translated, "slice_chains_serial": slice_serial,
                       # the same counters over the timed region only (all ranks): what a poor start-phase estimate
                       # cost INSIDE the measurement -- a serially chained slice or epoch sits on the inter-GPU ring
                       "timed_region": {"slice_chains_translated": slice_translated_timed,
                                        "slice_chains_serial": slice_serial_timed,
                                        "carrier_scan_serial_fallbacks": fallbacks_timed},
                       "runner": ("pipelined" if (world > 1 and use_pipelined(handoff)) else
                                  "lockstep" if world > 1 else "submit/fetch"),
                       "carrier_scan_chains": (args.warmup + args.steps) * E * C * world,
                       "numa_bound_cpus": numa_cpus},
            "e2e": {"value": round(e2e_value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": nbytes_desc,
                    "d2h_bytes_per_step": samples_per_step * 4, "how": e2e_how},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": kern_name, "kernel_ms_per_launch": round(kern_ms, 4),
                         "algorithmic_bytes_per_launch": kern_bytes,
                         "kernel_ms_in_pipeline": round(kms / kn, 4) if kn else None,
                         "step_level_achieved_gbs": round(value / world * 4 / 1e3, 2),
                         "step_level_frac": round(value / world * 4 / 1e3 / peak, 5),
                         "scan_phase_ms_per_step": round(scan_ms / max(nrec, 1), 4),
                         "render_phase_ms_per_step": round(synth_ms / max(nrec, 1), 4)},
            "clocks": clocks,
        }
        if par is not None:
            line["parity"] = par
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if par is not None and not par["ok"]:
        sys.stderr.write("bench.py: PARITY MISMATCH against the oracle (rank %d: first bad epoch %s: %s)\n"
                         % (rank, par.get("first_bad"), par.get("why")))
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--epochs", type=int, default=1024, help="epochs per step per GPU (1024 -> 1228.8 MB of output)")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--lookahead", type=int, default=int(os.environ.get("GPSIQ_LOOKAHEAD", "2")),
                    help="batches scanned ahead of the one being rendered (N = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling / trace runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the rendered batch (profiling runs)")
    ap.add_argument("--parity-stride", type=int, default=0)
    ap.add_argument("--carrier", choices=["float", "int32"], default="float",
                    help="float: the shipped build (FLOAT_CARR_PHASE, plutogpssim.h:12); int32: the reference's integer "
                         "carrier NCO (its #else branches), closed form")
    ap.add_argument("--lockstep", action="store_true",
                    help="N > 1: force the lockstep runner (rank 0 waits for the ring before it speculates, advances gathered inside the step)")
    ap.add_argument("--pipelined", action="store_true",
                    help="N > 1: force the pipelined runner (default: pipelined up to 4 GPUs, lockstep beyond)")
    ap.add_argument("--handoff", choices=["nccl", "mailbox"], default=os.environ.get("GPSIQ_HANDOFF", "mailbox"),
                    help="float-carrier phase hand-off between time slices (N > 1): the SM-free peer-memory mailbox "
                         "(copy engine + stream memory operations; falls back to NCCL if unavailable), or NCCL send/recv")
    ap.add_argument("--workload", choices=["config1", "config3"], default="config1",
                    help="config1 (default, the metric's configuration): 12 channels, 2.6 MS/s; config3: 32 channels, "
                         "10 MS/s, synthetic all-visible constellation (BASELINE configs [3]/[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
