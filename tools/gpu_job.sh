for C in c6 c7; do
cp gokalman_b200/libgokalman_b200_$C.so gokalman_b200/libgokalman_b200.so
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/t.json
python -c "import json;d=json.load(open('gpurun_out/t.json'));print('$C',d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
