python -m pytest tests/test_gpu_parity_mc.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_t.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -o "mc_finish_kernel.*" gpurun_out/launches_t.csv | tail -2 | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['value'])"
