timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 3 | tee gpurun_out/r02_gpu_tests_final.log
python -c "from cemc_b200 import _lib; print(_lib.source_hash())" > gpurun_out/r02_source_hash.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --variant 6 > gpurun_out/r02_bench_under_ncu.log 2>&1
for w in "C2 6" "C3S 1" "C3 8" "C4 9"; do set -- $w
  ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -f -o /tmp/r02_full_$1 python scripts/prof_wl.py $1 $2 4000 > gpurun_out/r02_full_$1.log 2>&1
  ncu -i /tmp/r02_full_$1.ncu-rep --page raw --csv > gpurun_out/r02_raw_$1.csv 2>/dev/null
  ncu -i /tmp/r02_full_$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_src_$1.csv 2>/dev/null
done
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 3 python scripts/sanitize_target.py 1 2 3 4 6 8 9 > gpurun_out/r02_san_$tool.log 2>&1
  tail -n 1 gpurun_out/r02_san_$tool.log
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python scripts/sanitize_target.py 1 2 3 4 6 8 9 > gpurun_out/r02_san_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 60 python scripts/sanitize_target.py 2 3 4 6 > gpurun_out/r02_san_racecheck_c1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench.err
