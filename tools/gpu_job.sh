set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_b.json
ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
