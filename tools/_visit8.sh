mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02k_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not 300s" 2>&1 | tail -3
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02k_n$n.json 2> gpurun_out/r02k_n$n.err; echo "n=$n rc $?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02k_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['config']['carrier_scan_serial_fallbacks'])"
grep PARITY gpurun_out/r02k_n$n.err
done
GPSIQ_NO_NUMA_BIND=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-parity > gpurun_out/r02k_n8_nobind.json 2> gpurun_out/r02k_n8_nobind.err; echo "rc $?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02k_n8_nobind.json').read().strip().splitlines()[-1]); print('nobind 8', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_n1.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02k_n1.json').read().strip().splitlines()[-1]); print(1, d['value'], d['ms_per_step'], d['e2e']['value'])"
