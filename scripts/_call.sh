python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_gpu_tests.log
