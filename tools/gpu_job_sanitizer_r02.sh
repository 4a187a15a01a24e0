#!/bin/bash
# compute-sanitizer on the kernels written / rewritten in round 2 (strict shared-memory kernel + scheduler, fused OD run).
for tool in memcheck racecheck synccheck; do
  echo "== $tool: strict kernel (scheduler forced, register twin) + fused OD run"
  timeout 900 compute-sanitizer --tool $tool python -m pytest "tests/test_gpu_strict.py::test_strict_scheduler_and_register_twin_are_bit_identical" "tests/test_gpu_strict.py::test_strict_hybrid_appd_constants_every_step" tests/test_gpu_od.py -m gpu -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|hazards|RACECHECK SUMMARY|error" | tail -6
done
