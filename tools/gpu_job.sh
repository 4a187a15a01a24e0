set -x
which compute-sanitizer
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity_tile.py tests/test_gpu_parity_nl.py tests/test_gpu_parity_batch.py -m gpu -x -q 2>&1 | tail -15
echo memcheck rc=$?
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q -k "every_step" 2>&1 | tail -15
echo racecheck tile rc=$?
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity_nl.py -m gpu -x -q -k "tma_staged" 2>&1 | tail -15
echo racecheck nl rc=$?
