for lib in _cemc_b200_prev.so _cemc_b200.so; do
CEMC_B200_LIB=cemc_b200/$lib timeout 300 python scripts/ab_probe.py C2 C3S C3 C4 2>&1
done | tee gpurun_out/r2y_ab.txt
