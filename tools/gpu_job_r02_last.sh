#!/bin/bash
# Last verification of round 2 (run under gpurun): GPU suite, smoke, both bench arms, and a full-set ncu capture of the
# headline (strict) kernel at the bench size with its raw page exported on the box.
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02h_bench_default_n1.json 2> gpurun_out/r02h_bench.err
python bench.py --impl reference > gpurun_out/r02h_bench_reference_n1.json 2>> gpurun_out/r02h_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hybrid_run_strict -s 3 -c 1 -f -o gpurun_out/r02h_prof_strict python bench.py --no-cpu-baseline --no-sub --steps 1 --warmup 3 > /dev/null 2>&1
ncu -i gpurun_out/r02h_prof_strict.ncu-rep --page raw --csv > gpurun_out/r02h_prof_strict.raw.csv 2>/dev/null
ncu -i gpurun_out/r02h_prof_strict.ncu-rep --page details > gpurun_out/r02h_prof_strict.details.txt 2>/dev/null
ls -la gpurun_out/r02h_*; find gpurun_out -name "r02h_prof_strict.ncu-rep" -size +40M -delete
tail -2 gpurun_out/r02h_bench.err
