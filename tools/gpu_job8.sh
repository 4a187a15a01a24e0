set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_n8.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --workload hybrid6 2>&1 | tail -1 | tee gpurun_out/bench_r01_hyb_n8.json | cut -c1-400
