timeout 900 python -m pytest tests/test_gpu_layered.py -q -x 2>&1 | tail -2
run() { timeout 300 python bench.py --config $2 --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r2_ab2_$2_$1.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_ab2_$2_$1.json')); k=d['roofline']['kernel_class_ms_per_step']; print('$1 $2', round(d['value']), round(d['ms_per_step'],3), 'wgrad', round(k['wgrad_nerf'],2), 'dgrad', round(k['chain_bwd_nerf'],2), 'fwd', round(k['chain_fwd_nerf'],2), d['clocks']['sm_mhz'])
PY
}
HUGS_WGRAD_PAIRS=0 run p0 Aprime
HUGS_WGRAD_PAIRS=1 run p1 Aprime
HUGS_WGRAD_PAIRS=0 run p0 360gin
HUGS_WGRAD_PAIRS=1 run p1 360gin
