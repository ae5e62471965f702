python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python scripts/e2e_probe.py C2 C3S 2>&1 | tee gpurun_out/r2_e2e_probe2.txt
