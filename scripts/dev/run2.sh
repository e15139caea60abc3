HUGS_WGRAD_PAIRS=1 timeout 600 ncu --set full --clock-control none -k regex:wgrad2_kernel -s 3 -c 1 -o gpurun_out/r2_wgrad2_pairs -f python bench.py --config Aprime --steps 1 --warmup 1 > gpurun_out/r2_ncu_w2.log 2>&1; tail -2 gpurun_out/r2_ncu_w2.log | cut -c1-200
HUGS_WGRAD_PAIRS=0 timeout 600 ncu --set full --clock-control none -k regex:wgrad_kernel -s 5 -c 1 -o gpurun_out/r2_wgrad1_wide -f python bench.py --config Aprime --steps 1 --warmup 1 > gpurun_out/r2_ncu_w1.log 2>&1; tail -2 gpurun_out/r2_ncu_w1.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
