set -x
python -m pytest tests/test_gpu_parity_tile.py tests/test_gpu_parity_nl.py -m gpu -x -q 2>&1 | tail -15
python bench.py --workload vanilla32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/tile_c.json
python -c "import json;d=json.load(open('gpurun_out/tile_c.json'));print(d['value'],d['roofline']['kernel_ms'],d['roofline']['machine_tflops'])"
