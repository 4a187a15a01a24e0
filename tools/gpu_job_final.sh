# Round-end check with every workload's bench line:  gpurun --timeout 1500 -- 'bash tools/gpu_job_final.sh > gpurun_out/job_final.log 2>&1'
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_final_reference.json | cut -c1-200
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-200
for w in mc_robot_info mc_robot_sqrt hybrid6 srif6 vanilla32 vanilla64; do
python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final_$w.json | cut -c1-200
done
GKB_BENCH_EVERY_STEP=1 python bench.py --workload hybrid6 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_final_hybrid6_every.json | cut -c1-200
python tools/bench_smooth.py 2>&1 | tail -1 | tee gpurun_out/bench_final_smooth.json
python tools/bench_lti.py 2>&1 | tail -4 | tee gpurun_out/bench_final_lti.json
