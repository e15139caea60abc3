run() { timeout 300 python bench.py --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r2_nb.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_nb.json')); k=d['roofline']['kernel_class_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],3), 'wgrad', round(k['wgrad_nerf'],3), round(k['wgrad_prop'],3), d['clocks']['sm_mhz'])
PY
}
run bias; HUGS_DBG_NOBIAS=1 run nobias; run bias; HUGS_DBG_NOBIAS=1 run nobias
