#!/bin/bash
# Round-2 profiling job, second pass (after the strict-kernel rewrite and the fused OD scheduler): launch list of the default
# bench command, DRAM traffic per launch of every workload's dominant kernel, full-set captures of the two kernels that changed.
# (gpurun brings back at most 64 MiB: the raw pages are exported here and only two reports are kept.)
set -x
B="python bench.py --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_default.csv $B --steps 2 --warmup 3 > gpurun_out/r02_b_launch.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_hybrid6.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:od_run_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_hybrid6_fused_od.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:hybrid_run_strict -s 3 -c 1 --csv --log-file gpurun_out/traffic_hybrid6_strict.csv $B --workload hybrid6_strict --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_srif6.csv $B --workload srif6 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mc_chisquare -s 3 -c 1 --csv --log-file gpurun_out/traffic_mc_jerk3.csv $B --workload mc_jerk3 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:vanilla_tile -s 3 -c 1 --csv --log-file gpurun_out/traffic_vanilla32.csv $B --workload vanilla32 --steps 1 --warmup 3 > /dev/null 2>&1
python tools/traffic_to_json.py gpurun_out/r02_traffic.json > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hybrid_run_strict -s 3 -c 1 -f -o gpurun_out/r02b_prof_hybrid6_strict $B --workload hybrid6_strict --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:od_run_kernel -s 1 -c 1 -f -o gpurun_out/r02b_prof_od_run $B --no-sub --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
for r in r02b_prof_hybrid6_strict r02b_prof_od_run; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; done
ls -la gpurun_out/; cat gpurun_out/r02_traffic.json | head -80
