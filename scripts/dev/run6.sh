python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -2 gpurun_out/r2_bench_default.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','vs_baseline','dtype','gpu_launches','clocks','e2e')})
print(r['class'], r['bound'], r['achieved'], r['peak'], r['frac'], r['traffic'], r.get('step_frac_of_tensor_peak'))
print(r['kernel_class_ms_per_step']); print(d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
