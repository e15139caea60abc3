timeout 900 python bench.py --steps 60 --warmup 10 > gpurun_out/r2_bench_check.json 2> gpurun_out/r2_bench_check.err; tail -2 gpurun_out/r2_bench_check.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2_bench_check.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], r['class'], r['frac'], r.get('tensor_lens'), [ (l['class'], l.get('tensor_lens')) for l in r['kernels'] if 'tensor_lens' in l])
PY
