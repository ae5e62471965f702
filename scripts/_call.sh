python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench.err
python bench.py > gpurun_out/r02_bench.json 2>> gpurun_out/r02_bench.err
tail -c 300 gpurun_out/r02_bench.json; tail -n 3 gpurun_out/r02_bench.err
