for lib in _cemc_b200_prev.so _cemc_b200.so _cemc_b200_noadapt.so; do
CEMC_B200_LIB=cemc_b200/$lib timeout 300 python scripts/v6_probe.py 2>&1
done | tee gpurun_out/r2t_v6.txt
