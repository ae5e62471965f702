python -c "from cemc_b200 import _lib; print(_lib.source_hash())" > gpurun_out/r02_source_hash.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
for w in "C2 3" "C3S 1" "C3 8" "C4 9"; do set -- $w
  ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1 -c 1 -f -o gpurun_out/r02_full_$1 python scripts/prof_wl.py $1 $2 4000 > gpurun_out/r02_full_$1.log 2>&1
  ncu -i gpurun_out/r02_full_$1.ncu-rep --page raw --csv > gpurun_out/r02_raw_$1.csv 2>/dev/null
  ncu -i gpurun_out/r02_full_$1.ncu-rep --page source --csv --print-source cuda > gpurun_out/r02_src_$1.csv 2>/dev/null
done
rm -f gpurun_out/r02_full_C3.ncu-rep gpurun_out/r02_full_C4.ncu-rep
ls -la gpurun_out/r02_*
tail -2 gpurun_out/r02_full_C2.log
