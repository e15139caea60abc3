timeout 600 python scripts/nerfacto_bench.py nerfacto --steps 1500 2>/dev/null | tail -1 > gpurun_out/r2_soak.json; python -c "
import json, math; d=json.load(open('gpurun_out/r2_soak.json')); print('soak 1500 steps', d['value'], d['ms_per_step'], d['loss'], math.isfinite(d['loss']))"
