# experiment: the n = 32 shape in pair mode (two warps per filter, T in registers) at 8 and 10 filters per SM
cp gokalman_b200/libgokalman_b200.so /tmp/lib_orig.so
for T in 512 640; do
cp build/exp/libexp_$T.so gokalman_b200/libgokalman_b200.so
echo "== pair32, $T threads per CTA"
python -m pytest tests/test_gpu_parity_tile.py -m gpu -x -q 2>&1 | tail -2
python bench.py --workload vanilla32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t_exp.json
python -c "import json;d=json.load(open('gpurun_out/t_exp.json'));print('n32',d['value'],d['roofline'].get('kernel_ms'),d['roofline']['machine_tflops'])"
done
cp /tmp/lib_orig.so gokalman_b200/libgokalman_b200.so
