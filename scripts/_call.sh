python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "site_split or golden or layered or LAYERED or case2 or case3 or recompute" 2>&1 | tail -5
python scripts/order_probe2.py 2>&1 | tee gpurun_out/r2_order_probe.txt
python scripts/pt_probe.py 2>&1 | grep -v "variant [12]:" | tee gpurun_out/r2_pt_probe2.txt
