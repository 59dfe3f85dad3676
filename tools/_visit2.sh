mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_timeslice.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_n2.json 2> gpurun_out/r02g_n2.err; echo rc $?; tail -c 1800 gpurun_out/r02g_n2.json; tail -3 gpurun_out/r02g_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --carrier int32 > gpurun_out/r02g_n2_int32.json 2> gpurun_out/r02g_n2_int32.err; echo rc $?; tail -c 1800 gpurun_out/r02g_n2_int32.json; tail -3 gpurun_out/r02g_n2_int32.err
