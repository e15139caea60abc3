run() { timeout 300 python bench.py --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r2_hint.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_hint.json')); k=d['roofline']['kernel_class_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],3), 'fwd', round(k['chain_fwd_nerf'],3), round(k['chain_fwd_prop'],3), 'wgrad', round(k['wgrad_nerf'],3), d['clocks']['sm_mhz'])
PY
}
run hint0; HUGS_FEAT_L2_HINT=1 run hint1; run hint0; HUGS_FEAT_L2_HINT=1 run hint1
