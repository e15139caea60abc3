N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2f_bench_weak_${N}gpu.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_weak_2gpu.json').read().strip().splitlines()[-1]); print('A2', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 300 $TR scripts/extra_configs.py frame 2>/dev/null | grep '^{' | cut -c40-60,330-400 | head -3
