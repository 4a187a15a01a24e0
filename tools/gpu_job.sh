ncu --set full --clock-control none --import-source on -k regex:vanilla_tile -s 3 -c 1 -o gpurun_out/prof_tile_r01c python bench.py --workload vanilla32 --trials 35520 --filter-steps 50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tile.log 2>&1
tail -2 gpurun_out/ncu_tile.log | cut -c1-200
