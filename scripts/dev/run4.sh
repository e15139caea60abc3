timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_full_gpu_tests.log; cat gpurun_out/r2_full_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], r['frac'], r.get('step_frac_of_tensor_peak'))
PY
timeout 300 python scripts/nerfacto_bench.py nerfacto --steps 60 2>/dev/null | tail -1 > gpurun_out/r02_nerfacto_bench_1gpu_final.json; python -c "
import json; d=json.load(open('gpurun_out/r02_nerfacto_bench_1gpu_final.json')); print('config4', d['value'], d['ms_per_step'])"
