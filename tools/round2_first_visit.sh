#!/bin/bash
# First GPU-box visit of the next round (nothing below could be run once round 1's GPU minutes were spent):
#   1. full GPU suite (includes the tests added on the CPU at the end of round 1: tests/test_zz_dropin.py -- the
#      reference's own main() driving the kernels -- and the radio-sink front-end test)
#   2. bench line
#   3. launch list + ncu --set full of the FINAL k_synth_line (TMA table fill, biased accumulator: DESIGN section 10, lead 5)
#   4. compute-sanitizer memcheck and racecheck on the smoke-sized path (shared-memory tables are
#      written by TMA bulk copies and read by all warps: racecheck is the tool that sees a missing barrier there)
# usage (under gpurun): bash tools/round2_first_visit.sh <tag>
tag=${1:-r02a}
mkdir -p gpurun_out
bash tools/gpu_round.sh "$tag"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' \
  > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' \
  > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_racecheck.log
tail -3 gpurun_out/${tag}_memcheck.log gpurun_out/${tag}_racecheck.log
