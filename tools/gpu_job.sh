set -x
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_c.json
python -c "import json;d=json.load(open('gpurun_out/bench_r01_c.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['step_ms'],d['e2e'],d['cpu_baseline'])"
for w in mc_robot_info mc_robot_sqrt; do
python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_$w.json
python -c "import json;d=json.load(open('gpurun_out/bench_r01_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['cpu_baseline'])"
done
python bench.py --workload srif6 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_srif6.json
python -c "import json;d=json.load(open('gpurun_out/bench_r01_srif6.json'));print('srif6',d['value'],d['roofline'],d['config'])"
ncu --set full --clock-control none --import-source on -k regex:vanilla_tile -s 3 -c 1 -o gpurun_out/prof_tile_r01b python bench.py --workload vanilla32 --trials 23680 --filter-steps 50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tile.log 2>&1
