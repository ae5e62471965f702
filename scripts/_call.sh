timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2j_gpu_tests.log
