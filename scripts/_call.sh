PT_L=12 PT_N=512 PT_STEPS=1728 PT_ROUNDS=6 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/pt_multi.py > gpurun_out/r02_pt_multi_n8.log 2>&1
tail -n 1 gpurun_out/r02_pt_multi_n8.log
