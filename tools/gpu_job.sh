set -x
python -m pytest tests/test_gpu_parity_nl.py -m gpu -x -q 2>&1 | tail -5
for p in tensor bulk plain; do
  GKB_NL_PATH=$p python bench.py --workload hybrid6 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/hyb_$p.json
  python -c "import json;d=json.load(open('gpurun_out/hyb_$p.json'));print('$p',d['value'],d['roofline']['kernel_ms'],d['roofline']['hbm']['frac'])"
done
