mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_png_gpu.py tests/test_webp_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2i_tests.txt; tail -3 gpurun_out/r2i_tests.txt
timeout 150 python bench.py --configs 3,4 --skip-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2i_bench34.json 2> gpurun_out/r2i_bench34.err
python -c "
import json
d=json.load(open('gpurun_out/r2i_bench34.json'))
c=d['configs']
print('png e2e', c['3']['e2e']['value'], 'device', c['3']['value'], 'out/in', c['3']['out_over_in_bytes'])
print({k:(round(v['ms'],3),v['launches']) for k,v in c['3']['device']['roofline']['all_kernels'].items()})
print('webp e2e', c['4']['e2e']['value'], c['4']['e2e']['images_per_sec'])
"
export B200_GRAPHS=0
timeout 150 ncu --set full --import-source on --clock-control none -k regex:"k_png_hashmatch|k_dfl_emit|k_dfl_hist|k_dfl_tables|k_dfl_len|k_png_match" -c 7 -f -o gpurun_out/r2i_png2 python tools/profile_legs.py png > /dev/null 2>&1
timeout 150 ncu --set full --import-source on --clock-control none -k regex:"k_vp8_encode|k_vp8_rgb_to_yuv" -c 2 -f -o gpurun_out/r2i_webp2 python tools/profile_legs.py webp > /dev/null 2>&1
ls -la gpurun_out | grep r2i
