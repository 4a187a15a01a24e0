ncu --set full --clock-control none --import-source on -k regex:mc_chisquare -s 3 -c 1 -o gpurun_out/prof_mc_r01_icdf python bench.py --filter-steps 300 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mc.log 2>&1
tail -2 gpurun_out/ncu_mc.log | cut -c1-200
