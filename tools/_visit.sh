mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms_in_pipeline"], d["roofline"]["kernel_ms_per_launch"], d["parity"]["ok"], d["config"]["carrier_scan_serial_fallbacks"])'
for sp in hi lo; do for rp in 0 -1; do
echo "== scan prio $sp render prio $rp"
GPSIQ_SCAN_PRIO=$sp GPSIQ_RENDER_PRIO=$rp timeout 600 python bench.py --steps 10 --warmup 3 --lookahead 2 --no-cpu-baseline 2>&1 | python -c "$P"
done; done
GPSIQ_TRACE=2 GPSIQ_SCAN_PRIO=lo GPSIQ_RENDER_PRIO=-1 timeout 600 python bench.py --steps 6 --warmup 3 --lookahead 2 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r02h_trace.txt
