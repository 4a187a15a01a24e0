#!/bin/bash
# Final round-2 job (run under gpurun): both bench arms with the strict (bit-exact) kernel as the headline, the launch
# list of the default command, DRAM traffic of one launch of the headline kernel and of its fused OD run.
set -x
python bench.py > gpurun_out/r02g_bench_default_n1.json 2> gpurun_out/r02g_bench.err
python bench.py --impl reference > gpurun_out/r02g_bench_reference_n1.json 2>> gpurun_out/r02g_bench.err
B="python bench.py --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g_launches_default.csv $B --steps 2 --warmup 3 > gpurun_out/r02g_b_launch.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:hybrid_run_strict -s 3 -c 1 --csv --log-file gpurun_out/r02g_traffic_hybrid6.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:od_run_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02g_traffic_hybrid6_fused_od.csv $B --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
tail -3 gpurun_out/r02g_bench.err; wc -c gpurun_out/r02g_*
