timeout 900 python -m pytest tests/test_gpu_layered.py -q -x 2>&1 | tail -2
run() { timeout 300 python bench.py --config $1 --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/r2_pack_$1.json; python - <<PY
import json; d=json.load(open('gpurun_out/r2_pack_$1.json')); k=d['roofline']['kernel_class_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],3), 'adam_pack', round(k['adam_pack'],3), d['clocks']['sm_mhz'])
PY
}
run Aprime; run 360gin
