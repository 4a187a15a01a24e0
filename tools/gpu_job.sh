python -m pytest tests/test_gpu_parity_nl.py -m gpu -x -q 2>&1 | tail -4
for p in tensor plain; do
GKB_BENCH_EVERY_STEP=1 GKB_NL_PATH=$p python bench.py --workload hybrid6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('every-step $p',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm'])"
done
python bench.py --workload hybrid6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('final-only',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
