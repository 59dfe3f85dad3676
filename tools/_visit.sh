mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for la in 1 2; do 
echo "== lookahead $la"
timeout 600 python bench.py --steps 10 --warmup 3 --lookahead $la --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms_in_pipeline'], d['roofline']['kernel_ms_per_launch'], d['parity']['ok'], d['config']['carrier_scan_serial_fallbacks'], d['e2e']['value'])"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > /dev/null 2>&1
