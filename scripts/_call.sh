timeout 900 compute-sanitizer --tool racecheck --print-limit 60 python scripts/sanitize_target.py 2 3 4 6 > gpurun_out/r02_san_racecheck_c1.log 2>&1
tail -2 gpurun_out/r02_san_racecheck_c1.log
grep -o "in cemc_[a-z_]*.cuh:[0-9]*" gpurun_out/r02_san_racecheck_c1.log | sort | uniq -c | sort -rn | head
