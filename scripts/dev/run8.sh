N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --config Aprime --steps 40 --warmup 5 > gpurun_out/r2b_bench_Aprime_${N}gpu.json 2> gpurun_out/r2b_bench_Aprime_${N}gpu.err; tail -c 300 gpurun_out/r2b_bench_Aprime_${N}gpu.json | head -c 10; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_Aprime_8gpu.json').read().strip().splitlines()[-1]); print('Aprime8', d['value'], d['ms_per_step'])"
timeout 300 $TR scripts/extra_configs.py frame > gpurun_out/r2b_frame_sweep_${N}gpu.jsonl 2> gpurun_out/r2b_frame_${N}gpu.err; cut -c1-40,330-460 gpurun_out/r2b_frame_sweep_${N}gpu.jsonl; tail -2 gpurun_out/r2b_frame_${N}gpu.err
timeout 300 $TR scripts/nerfacto_bench.py nerfacto --steps 40 2>/dev/null | tail -1 > gpurun_out/r2b_nerfacto_${N}gpu.jsonl; python -c "
import json; d=json.loads(open('gpurun_out/r2b_nerfacto_8gpu.jsonl').read().strip().splitlines()[-1]); print('nerfacto8', d['value'], d['ms_per_step'], d['phase_ms'])"
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2b_bench_weak_${N}gpu.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_weak_8gpu.json').read().strip().splitlines()[-1]); print('A8', d['value'], d['ms_per_step'], d['e2e']['value'])"
