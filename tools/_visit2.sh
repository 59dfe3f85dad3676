mkdir -p gpurun_out
GPSIQ_TRACE=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --no-parity --no-e2e > /dev/null 2> gpurun_out/r02n_trace_n2.txt
grep -c "trace dev" gpurun_out/r02n_trace_n2.txt
