mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_png_gpu.py tests/test_baseline_configs_gpu.py tests/test_webp_gpu.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2k_tests.txt; tail -3 gpurun_out/r2k_tests.txt
timeout 150 python bench.py --configs 3 --skip-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2k_bench3.json 2> gpurun_out/r2k_bench3.err
python -c "
import json
d=json.load(open('gpurun_out/r2k_bench3.json'))
c=d['configs']
print('png e2e', c['3']['e2e']['value'], 'device', c['3']['value'], 'out/in', c['3']['out_over_in_bytes'])
print({k:(round(v['ms'],3),v['launches']) for k,v in c['3']['device']['roofline']['all_kernels'].items()})
"
(timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_all.py 2>&1 | tail -25) > gpurun_out/r2k_memcheck.txt; tail -3 gpurun_out/r2k_memcheck.txt
(timeout 250 compute-sanitizer --tool racecheck python tools/sanitize_all.py 2>&1 | tail -25) > gpurun_out/r2k_racecheck.txt; tail -3 gpurun_out/r2k_racecheck.txt
