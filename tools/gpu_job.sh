python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r01_d.json
python -c "import json;d=json.load(open('gpurun_out/bench_r01_d.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
python bench.py --workload srif6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/srif_d.json
python -c "import json;d=json.load(open('gpurun_out/srif_d.json'));print('srif',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
