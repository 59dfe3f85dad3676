#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu --set full of the dominant kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_synth_line -s 2 -c 1 -o gpurun_out/${tag}_line \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
