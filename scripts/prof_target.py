"""Small single-launch workloads for ncu (never a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scripts.probe import setup, KB

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
if which == "c2":
    T = np.linspace(200, 1000, 16)
    ft, gpu = setup(10, ["Al", "Mg"], {"Al": 0.5, "Mg": 0.5}, 256, np.repeat(T * KB, 16), np.zeros(256))
    gpu.run_sgc(n); gpu.synchronize()
    gpu.run_sgc(n); gpu.synchronize()
elif which in ("c3", "c3s"):
    ft, gpu = setup(20, ["Al", "Mg", "Si"], {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, 64, np.linspace(300, 900, 64) * KB)
    if which == "c3":
        gpu.run_canonical(n); gpu.synchronize()
        gpu.run_canonical(n); gpu.synchronize()
    else:
        gpu.run_sgc(n); gpu.synchronize()
if which == "c3s":
    gpu.run_sgc(n); gpu.synchronize()
