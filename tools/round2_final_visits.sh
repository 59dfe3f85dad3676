#!/bin/bash
# The GPU visits behind the end-of-round-2 numbers (profiles/r02_v .. r02_y), as run under gpurun.
#   bash tools/round2_final_visits.sh one      # 1 GPU : suite, bench lines, ncu of k_synth_line, launch list, trace   (r02_w)
#   bash tools/round2_final_visits.sh scale N  # N GPUs: bench.py --gpus N --no-e2e, config[1] (+ config[4] if "config3" follows)  (r02_v / r02_x)
#   bash tools/round2_final_visits.sh ring2    # 2 GPUs: the 2-GPU tests, pipelined vs lockstep runner, trace with host enqueue times
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d["config"]; print(d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["kernel_ms_in_pipeline"], d.get("parity",{}).get("ok"), c["slice_chains_translated"], c["slice_chains_serial"], c["carrier_scan_serial_fallbacks"], c.get("timed_region"), c["runner"])'
run() { n=$1; tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((RANDOM%90+10)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-e2e "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; echo "$tag rc $?"
  python -c "$P" < gpurun_out/${tag}.json; grep -h "PARITY\|Error\|error" gpurun_out/${tag}.err | head -5; }
case "$1" in
one)
  tag=r02w
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
  timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
  timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  timeout 200 python bench.py --workload config3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
  timeout 200 python bench.py --carrier int32 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_int32.json 2>/dev/null
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_synth_line -s 2 -c 1 -o gpurun_out/${tag}_line \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/${tag}_ncu_line.log 2>&1
  M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
  timeout 200 ncu --metrics $M --clock-control none -c 140 --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > /dev/null 2>&1
  GPSIQ_TRACE=2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > /dev/null 2> gpurun_out/${tag}_trace.txt
  # afterwards, here: python tools/ncu_summary.py gpurun_out/${tag}_line.ncu-rep profiles/r02_w_synth_line_ncu_full.txt
  #                   python tools/update_traffic.py gpurun_out/${tag}_line.ncu-rep 12 1024 profiles/r02_w_synth_line_ncu_full.txt
  ;;
scale)
  n=$2
  run $n r02x_n${n}
  if [ "$3" = "config3" ]; then run $n r02x_n${n}_config3 --workload config3; fi
  ;;
ring2)
  timeout 600 python -m pytest tests/test_gpu_timeslice.py -x -q > gpurun_out/r02t_pytest_timeslice.log 2>&1; tail -2 gpurun_out/r02t_pytest_timeslice.log
  run 2 r02v_n2_pipe
  run 2 r02v_n2_lock --lockstep
  GPSIQ_TS_FREE=1 run 2 r02y_n2_free
  GPSIQ_TRACE=2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 8 --warmup 3 --no-parity --no-e2e > /dev/null 2> gpurun_out/r02v_trace_n2.txt
  ;;
*) echo "usage: $0 one | scale N [config3] | ring2"; exit 2 ;;
esac
