mkdir -p gpurun_out
for n in 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02m_n$n.json 2> gpurun_out/r02m_n$n.err; echo "n=$n rc $?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02m_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['config']['carrier_scan_serial_fallbacks'])"
grep PARITY gpurun_out/r02m_n$n.err
done
for c in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --lookahead 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('conn $c', 1, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel_ms_in_pipeline'])"
done
