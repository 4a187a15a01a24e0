# Strict (reference-order) hybrid kernel: bit-exactness tests on every A/B path, then the hybrid6_strict bench line per path.
for p in prefetch plainload l2 regs; do
  echo "== GKB_STRICT_PATH=$p"
  GKB_STRICT_PATH=$p python -m pytest tests/test_gpu_strict.py tests/test_gpu_crosscheck.py -m gpu -x -q 2>&1 | tail -3
  GKB_STRICT_PATH=$p python bench.py --workload hybrid6_strict --filter-steps 200 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/strict_$p.json
  python -c "import json;d=json.load(open('gpurun_out/strict_$p.json'));print('$p',d['value'],d['ms_per_step'],d['roofline'].get('kernel_ms'),d['roofline']['frac'],d['clocks'])"
done
python -m pytest tests/test_gpu_parity_nl.py tests/test_gpu_od.py -m gpu -x -q 2>&1 | tail -3
