mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02i_n$n.json 2> gpurun_out/r02i_n$n.err; echo "n=$n rc $?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02i_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
grep PARITY gpurun_out/r02i_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --carrier int32 > gpurun_out/r02i_n8_int32.json 2> gpurun_out/r02i_n8_int32.err; echo "rc $?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r02i_n8_int32.json').read().strip().splitlines()[-1]); print('int32 8', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_n1.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02i_n1.json').read().strip().splitlines()[-1]); print(1, d['value'], d['ms_per_step'], d['e2e']['value'])"
