mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms_in_pipeline"], d["roofline"]["kernel_ms_per_launch"], d["parity"]["ok"], d["config"]["carrier_scan_serial_fallbacks"], d["e2e"]["value"])'
for la in 1 2 2; do
echo "== lookahead $la"
timeout 600 python bench.py --steps 20 --warmup 3 --lookahead $la --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "== int32"; timeout 600 python bench.py --steps 20 --warmup 3 --lookahead 2 --no-cpu-baseline --carrier int32 2>&1 | python -c "$P"
echo "== config3"; timeout 600 python bench.py --steps 20 --warmup 3 --lookahead 2 --no-cpu-baseline --workload config3 2>&1 | python -c "$P"
GPSIQ_TRACE=2 timeout 600 python bench.py --steps 6 --warmup 3 --lookahead 2 --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r02h_trace.txt
