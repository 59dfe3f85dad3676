mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["kernel_ms_in_pipeline"], d["roofline"]["kernel_ms_per_launch"], d["parity"]["ok"], d["config"]["carrier_scan_serial_fallbacks"], d["e2e"]["value"])'
for la in 1 2 3; do
echo "== lookahead $la"
timeout 600 python bench.py --steps 20 --warmup 3 --lookahead $la --no-cpu-baseline 2>&1 | python -c "$P"
done
