for lib in _cemc_b200_prev.so _cemc_b200.so; do
CEMC_B200_LIB=cemc_b200/$lib timeout 300 python scripts/ab_probe.py C2 C3S C3 C4 C5 2>&1
done | tee gpurun_out/r2x_ab.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 6 | tee gpurun_out/r2x_tests.log
