HUGS_NF_CC_HEADS=1 timeout 600 python -m pytest tests/test_gpu_nerfacto_hash.py -q -x 2>&1 | tail -2
for G in 0 1 0 1; do HUGS_NF_CC_HEADS=$G timeout 300 python scripts/nerfacto_bench.py nerfacto --steps 60 2>/dev/null | tail -1 > gpurun_out/r2_nerfacto_cc_$G.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_nerfacto_cc_$G.json')); print('cc_heads=$G', round(d['value']), round(d['ms_per_step'],2), round(d['phase_ms']['forward'],2), round(d['phase_ms']['backward'],2), d['loss'])
PY
done
