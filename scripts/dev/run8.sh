N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/nerfacto_bench.py nerfacto --steps 60 2>/dev/null | tail -1 > gpurun_out/r2c_nerfacto_${N}gpu.jsonl; python -c "
import json; d=json.loads(open('gpurun_out/r2c_nerfacto_8gpu.jsonl').read().strip().splitlines()[-1]); print('nerfacto8', d['value'], d['ms_per_step'], d['phase_ms'])"
timeout 300 $TR scripts/nerfacto_bench.py nerf --steps 60 2>/dev/null | tail -2 > gpurun_out/r2c_nerf_${N}gpu.jsonl; python -c "
import json
for l in open('gpurun_out/r2c_nerf_8gpu.jsonl'):
  if l.startswith('{'):
    d=json.loads(l); print('nerf8', d['value'], d['ms_per_step'], d['rays_per_gpu'])"
