timeout 600 python -m pytest tests/test_gpu_layered.py -x -q 2>&1 | tail -15 > gpurun_out/r2_w2_tests.log; cat gpurun_out/r2_w2_tests.log
for P in 0 1; do HUGS_WGRAD_PAIRS=$P timeout 300 python bench.py --config Aprime --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_w2_aprime_p$P.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_w2_aprime_p$P.json')); print('pairs=$P', d['value'], d['ms_per_step'], d.get('kernels_ms', d.get('roofline')))
PY
done
timeout 300 python scripts/extra_configs.py frame > gpurun_out/r2_frame_sweep_1gpu.jsonl 2> gpurun_out/r2_frame_sweep.err; tail -3 gpurun_out/r2_frame_sweep.err; cut -c1-330 gpurun_out/r2_frame_sweep_1gpu.jsonl
