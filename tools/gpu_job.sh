set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_icdf.json | cut -c1-200
python -c "import json;d=json.load(open('gpurun_out/bench_icdf.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['config']['nis_mean'],d['config']['nees_mean'])"
for w in mc_robot_info mc_robot_sqrt; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_icdf_$w.json
python -c "import json;d=json.load(open('gpurun_out/bench_icdf_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
