timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2o_gpu_tests.log
timeout 300 python scripts/ab_probe.py C2 C3S C3 C4 C5 2>&1 | tee gpurun_out/r2o_ab.txt
