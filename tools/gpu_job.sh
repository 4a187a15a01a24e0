# The round-end check, as one gpurun command:  gpurun --timeout 1500 -- 'bash tools/gpu_job.sh > gpurun_out/job.log 2>&1'
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_final_reference.json | cut -c1-300
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
