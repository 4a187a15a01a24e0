python -m pytest tests/test_gpu_parity_mc.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload mc_robot_sqrt --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('mc_robot_sqrt',d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
