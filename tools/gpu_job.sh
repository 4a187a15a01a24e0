python -m pytest tests/test_gpu_parity_nl.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload srif6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/srif_c.json
python -c "import json;d=json.load(open('gpurun_out/srif_c.json'));print('srif',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
