python tools/bench_lti.py 2>&1 | tail -4
