timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python bench.py --workload hybrid6 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/hyb_e.json
python -c "import json;d=json.load(open('gpurun_out/hyb_e.json'));print('hyb',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'],d['cpu_baseline'])"
timeout 120 python bench.py --workload srif6 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/srif_g.json
python -c "import json;d=json.load(open('gpurun_out/srif_g.json'));print('srif',d['value'],d['roofline']['kernel_ms'],d['cpu_baseline'])"
timeout 300 python bench.py --workload vanilla32 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/tile_d.json
python -c "import json;d=json.load(open('gpurun_out/tile_d.json'));print('tile',d['value'],d['roofline']['kernel_ms'],d['cpu_baseline'])"
