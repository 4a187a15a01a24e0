# large-state kernel check: parity tests, racecheck of the warp-pair shapes, the n = 32 bench line (regression), n = 48 / 64 timings
python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q 2>&1 | tail -5
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q -k "64-2-4 or 48-5-5" 2>&1 | tail -6
python bench.py --workload vanilla32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_v32.json
python -c "import json;d=json.load(open('gpurun_out/t_v32.json'));print('n32',d['value'],d['roofline'].get('kernel_ms'),d['roofline']['frac'])"
python bench.py --workload vanilla64 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_v64.json
python -c "import json;d=json.load(open('gpurun_out/t_v64.json'));print('n64',d['value'],d['roofline'].get('kernel_ms'),d['roofline']['machine_tflops'])"
GKB_BENCH_TILE_N=48 python bench.py --workload vanilla32 --trials 35520 --filter-steps 100 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_v48.json
python -c "import json;d=json.load(open('gpurun_out/t_v48.json'));print('n48',d['value'],d['roofline'].get('kernel_ms'),d['roofline']['machine_tflops'])"
