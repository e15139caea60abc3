timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_full_gpu_tests.log; cat gpurun_out/r2_full_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], r['frac'], r.get('step_frac_of_tensor_peak'))
PY
