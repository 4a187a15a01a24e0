python tools/bench_smooth.py 2>&1 | tail -2
