mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_webp_gpu.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r2p_webp_tests.txt; tail -1 gpurun_out/r2p_webp_tests.txt
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2p_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/r2p_pytest_gpu.log | tail -1
timeout 400 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2p_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], {k: (v.get('value'), v.get('e2e', {}).get('value'), v.get('e2e', {}).get('images_per_sec')) for k, v in d.get('configs', {}).items()})
print('webp d2h/step', d['configs']['4']['e2e']['d2h_bytes_per_step'])"
B200_TRACE=2 timeout 100 python bench.py --configs 4 --only-configs --skip-cpu-baseline --steps 6 > gpurun_out/r2p_trace_webp.json 2> gpurun_out/r2p_trace_webp.err; grep "trace\] jpeg" gpurun_out/r2p_trace_webp.err | tail -3
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2p_bench_reference.json 2>> gpurun_out/r2p_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
