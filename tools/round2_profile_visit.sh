#!/bin/bash
# One GPU-box visit that produces everything profiles/ holds for the final code of round 2:
# GPU suite log, bench lines (float / int32 / 32 channels) + reference arms, launch list with instruction counts,
# ncu --set full of k_synth_line (12 and 32 channels) and of the scan kernels, sanitizer logs, a steady-state trace.
# usage (under gpurun): bash tools/round2_profile_visit.sh <tag>
tag=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2>&1
timeout 600 python bench.py --carrier int32 > gpurun_out/${tag}_bench_int32.json 2> gpurun_out/${tag}_bench_int32.err
timeout 600 python bench.py --impl reference --carrier int32 --steps 3 --warmup 1 > gpurun_out/${tag}_ref_int32.json 2>&1
timeout 600 python bench.py --workload config3 > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
timeout 600 python bench.py --impl reference --workload config3 --steps 3 --warmup 1 > gpurun_out/${tag}_ref_config3.json 2>&1
timeout 600 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_long.json 2> /dev/null
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 600 ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_synth_line -s 2 -c 1 -o gpurun_out/${tag}_line python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/${tag}_ncu_line.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_synth_line -s 2 -c 1 -o gpurun_out/${tag}_line32 python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/${tag}_ncu_line32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_carr_|k_line_anchor|k_line_check|k_prepare" -s 16 -c 8 -o gpurun_out/${tag}_scan python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/${tag}_ncu_scan.log 2>&1
GPSIQ_TRACE=2 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > /dev/null 2> gpurun_out/${tag}_trace.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck.log
tail -2 gpurun_out/${tag}_memcheck.log gpurun_out/${tag}_racecheck.log
ls -la gpurun_out | grep ${tag}
