ncu --set full --clock-control none --import-source on -k regex:nl_run_wtma -s 3 -c 1 -o gpurun_out/prof_srif_r01 python bench.py --workload srif6 --filter-steps 100 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_srif.log 2>&1
tail -3 gpurun_out/ncu_srif.log
