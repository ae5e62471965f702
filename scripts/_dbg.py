import sys; sys.path.insert(0,'/root/repo')
from cemc_b200 import workloads as wl
w = wl.c3s_almgsi_sgc(R=1)
gpu = wl.make_updater(w)
gpu.set_batch(4)
gpu.run_sgc(40); gpu.synchronize()
