set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --workload srif6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/srif_b.json
python -c "import json;d=json.load(open('gpurun_out/srif_b.json'));print('srif',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
GKB_NL_PATH=plain python bench.py --workload srif6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/srif_plain.json
python -c "import json;d=json.load(open('gpurun_out/srif_plain.json'));print('srif plain',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
python bench.py --workload hybrid6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/hyb_c.json
python -c "import json;d=json.load(open('gpurun_out/hyb_c.json'));print('hyb',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
for w in mc_robot_info mc_robot_sqrt; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
