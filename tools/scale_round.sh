#!/bin/bash
# One multi-GPU box visit: time-slice parity (2 GPUs) and the weak-scaling bench with both hand-offs.
# usage (under `gpurun --gpus N`): bash tools/scale_round.sh <tag> <N> [steps]
tag=${1:-x}; n=${2:-2}; steps=${3:-10}
mkdir -p gpurun_out
if [ "$n" -ge 2 ]; then
  timeout -k 5 150 python -m pytest tests/test_gpu_timeslice.py -x -q > gpurun_out/${tag}_pytest_timeslice.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest_timeslice.log
  tail -3 gpurun_out/${tag}_pytest_timeslice.log
fi
for h in mailbox nccl; do
  timeout -k 5 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps $steps --warmup 3 --no-cpu-baseline --handoff $h \
    > gpurun_out/${tag}_bench_n${n}_$h.json 2> gpurun_out/${tag}_bench_n${n}_$h.err
  echo "$h exit $?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/${tag}_bench_n${n}_$h.json
done
