for v in "0 8" "1500 8" "3000 8" "6000 8" "0 6" "0 4"; do
  set -- $v
  GKB_TILE_STAGGER_NS=$1 GKB_TILE_WARPS=$2 python bench.py --workload vanilla32 --trials 47360 --filter-steps 100 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
  python -c "import json;d=json.load(open('gpurun_out/t.json'));print('$v',d['value'],d['roofline']['kernel_ms'],d['roofline']['machine_tflops'])"
done
