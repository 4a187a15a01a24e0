# large-state kernel check: parity tests, then the n = 32 bench line (regression) and an n = 64 timing
python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q 2>&1 | tail -5
python bench.py --workload vanilla32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_v32.json
python -c "import json;d=json.load(open('gpurun_out/t_v32.json'));print('n32',d['value'],d['roofline'].get('kernel_ms'),d['roofline']['frac'])"
for n in 48 64; do
GKB_BENCH_TILE_N=$n python bench.py --workload vanilla32 --trials 23680 --filter-steps 100 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_v$n.json
python -c "import json;d=json.load(open('gpurun_out/t_v$n.json'));print('n$n',d['value'],d['roofline'].get('kernel_ms'))"
done
