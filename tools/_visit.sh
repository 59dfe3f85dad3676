mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "rc $?"; tail -c 2500 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --carrier int32 > gpurun_out/r02c_bench_int32.json 2> gpurun_out/r02c_bench_int32.err; echo "rc $?"; tail -c 2500 gpurun_out/r02c_bench_int32.json; tail -5 gpurun_out/r02c_bench_int32.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload config3 > gpurun_out/r02c_bench_config3.json 2> gpurun_out/r02c_bench_config3.err; echo "rc $?"; tail -c 2500 gpurun_out/r02c_bench_config3.json; tail -5 gpurun_out/r02c_bench_config3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --workload config3 > gpurun_out/r02c_ref_config3.json 2>&1; tail -c 1500 gpurun_out/r02c_ref_config3.json
