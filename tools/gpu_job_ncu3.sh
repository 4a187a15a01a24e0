# final captures of round 1: the n = 64 pair kernel (full set) and the launch list of the default bench command
set -x
ncu --set full --clock-control none --import-source on -k regex:vanilla_tile -s 3 -c 1 -f -o gpurun_out/prof_tile64_r01b python bench.py --workload vanilla64 --trials 17760 --filter-steps 50 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launch.log 2>&1
tail -5 gpurun_out/launches_final.csv | cut -c1-200
bash tools/measure_traffic.sh > /dev/null 2>&1
python tools/traffic_to_json.py gpurun_out/traffic.json > /dev/null 2>&1; cat gpurun_out/traffic.json | head -50
