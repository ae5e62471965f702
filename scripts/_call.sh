timeout 120 python scripts/ab_probe.py C3S 2>&1 | tee gpurun_out/r2e_first.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2e_gpu_tests.log
for lib in _cemc_b200_base.so _cemc_b200.so; do
CEMC_B200_LIB=cemc_b200/$lib timeout 300 python scripts/ab_probe.py C2 C3S C3 C4 C5 2>&1
done | tee gpurun_out/r2e_ab.txt
tail -5 gpurun_out/r2e_gpu_tests.log
