python -m pytest tests/test_gpu_parity_tile.py tests/test_gpu_fullsize.py -m gpu -x -q -k "tile or vanilla32" 2>&1 | tail -5
for w in 8 12; do
GKB_TILE_WARPS=$w python bench.py --workload vanilla32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('n32 warps',$w,d['value'],d['roofline']['kernel_ms'],d['roofline']['machine_tflops'])"
done
