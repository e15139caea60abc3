timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_d.log python -m pytest tests/test_gpu_nerfacto_hash.py -q -x -k "forward_vs_reference or loss_and_gradients or training_decreases" 2>&1 | tail -3; tail -3 gpurun_out/r2_sanitizer_d.log
timeout 300 python -m pytest tests/test_gpu_nerfacto_hash.py -q 2>&1 | tail -2
