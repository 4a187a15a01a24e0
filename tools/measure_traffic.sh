#!/bin/bash
# DRAM traffic per launch of the dominant kernel of each bench workload (ncu, one launch each, bench configuration).
# Writes gpurun_out/traffic_<workload>.csv; tools/traffic_to_json.py folds them into profiles/rNN_traffic.json.
set -x
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:mc_chisquare -s 3 -c 1 --csv --log-file gpurun_out/traffic_mc_jerk3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_hybrid6.csv python bench.py --workload hybrid6 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:nl_run_wtma -s 3 -c 1 --csv --log-file gpurun_out/traffic_srif6.csv python bench.py --workload srif6 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:vanilla_tile -s 3 -c 1 --csv --log-file gpurun_out/traffic_vanilla32.csv python bench.py --workload vanilla32 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:vanilla_tile -s 3 -c 1 --csv --log-file gpurun_out/traffic_vanilla64.csv python bench.py --workload vanilla64 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mc_chisquare -s 3 -c 1 --csv --log-file gpurun_out/traffic_mc_robot_info.csv python bench.py --workload mc_robot_info --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mc_chisquare -s 3 -c 1 --csv --log-file gpurun_out/traffic_mc_robot_sqrt.csv python bench.py --workload mc_robot_sqrt --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 4 gpurun_out/traffic_*.csv
