# ncu --set full captures of the kernels changed late in round 1 (one launch each, the bench configurations)
set -x
ncu --set full --clock-control none --import-source on -k regex:nl_run_wtma -s 3 -c 1 -f -o gpurun_out/prof_srif_r01b python bench.py --workload srif6 --filter-steps 100 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:smooth_all -s 2 -c 1 -f -o gpurun_out/prof_smooth_r01 python tools/bench_smooth.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:vanilla_tile -s 3 -c 1 -f -o gpurun_out/prof_tile64_r01 python bench.py --workload vanilla64 --trials 13320 --filter-steps 50 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
