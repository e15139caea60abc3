timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_full_gpu_tests.log; cat gpurun_out/r2_full_gpu_tests.log
timeout 600 python bench.py --config Aprime --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r02_bench_Aprime_1gpu_v2.json
timeout 600 python bench.py --config 360gin --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r02_bench_360gin_1gpu_v2.json
python - <<PY
import json
for c in ('Aprime','360gin'):
  d=json.load(open('gpurun_out/r02_bench_%s_1gpu_v2.json'%c)); print(c, d['value'], d['ms_per_step'], d['roofline'].get('step_frac_of_tensor_peak'), d['clocks'])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"wgrad2_kernel|dense_tc_kernel" -s 30 -c 4 -o gpurun_out/r2_wide_bwd_final -f python bench.py --config Aprime --steps 1 --warmup 1 > gpurun_out/r2_ncu_final.log 2>&1; tail -1 gpurun_out/r2_ncu_final.log
