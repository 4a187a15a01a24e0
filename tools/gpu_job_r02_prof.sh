#!/bin/bash
# Round-2 profiling job (run under gpurun): launch list of the default bench command, DRAM traffic per launch of the
# dominant kernel of each workload, full-set captures of the three hybrid kernels (TMA production, strict, fused OD).
set -x
B="python bench.py --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv $B --steps 2 --warmup 3 > gpurun_out/r02_b_launch.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_hybrid6.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:od_run_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_hybrid6_fused_od.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:hybrid_run_strict -s 3 -c 1 --csv --log-file gpurun_out/traffic_hybrid6_strict.csv $B --workload hybrid6_strict --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_srif6.csv $B --workload srif6 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mc_chisquare -s 3 -c 1 --csv --log-file gpurun_out/traffic_mc_jerk3.csv $B --workload mc_jerk3 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:vanilla_tile -s 3 -c 1 --csv --log-file gpurun_out/traffic_vanilla32.csv $B --workload vanilla32 --steps 1 --warmup 3 > /dev/null 2>&1
python tools/traffic_to_json.py gpurun_out/r02_traffic.json > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nl_run_wtma -s 3 -c 1 -f -o gpurun_out/r02_prof_hybrid6 $B --no-sub --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hybrid_run_strict -s 3 -c 1 -f -o gpurun_out/r02_prof_hybrid6_strict $B --workload hybrid6_strict --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:od_run_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_od_run $B --no-sub --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nl_run_wtma -s 3 -c 1 -f -o gpurun_out/r02_prof_srif6 $B --workload srif6 --filter-steps 200 --steps 1 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/r02_traffic.json | head -60
