# SRIF fast-epoch / SmoothAll check: parity tests of the NLDKF path, then the srif6 bench line and the smoother timing.
python -m pytest tests/test_gpu_parity_nl.py tests/test_gpu_fullsize.py tests/test_gpu_parity_batch.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_smooth.py 2>&1 | tail -1
for w in srif6; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_$w.json
python -c "import json;d=json.load(open('gpurun_out/t_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline'].get('kernel_ms'),d['roofline']['frac'],d['e2e']['value'])"
done
