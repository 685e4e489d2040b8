#!/bin/bash
# Round-end evidence run (on a GPU box): GPU test suite, both bench arms, launch lists and ncu --set full captures.
# Everything lands in gpurun_out/; tools/summarise_profiles.py cuts the committed summaries from it.
# The ncu runs replay nothing from a CUDA graph (B200_GRAPHS=0): every launch is listed under its own name.
TAG=${1:-r2}
mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -1
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
export B200_GRAPHS=0
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --configs 1 --batch 16 --e2e-batch 16 --skip-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_group.csv python tools/profile_group.py 8 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_legs.csv python tools/profile_legs.py > /dev/null 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_fused_same|k_chroma420_refdct|k_idct_plane" -c 3 -f -o gpurun_out/${TAG}_transform python bench.py --steps 1 --warmup 3 --configs 1 --batch 64 --only-value > /dev/null 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_geb_emit|k_geb_hist|k_geb_len|k_geb_classify|k_gd_write|k_gd_round0|k_ge_tables|k_ge_groups|k_gd_unstuff_scatter|k_gd_dc_gather" -c 10 -f -o gpurun_out/${TAG}_entropy python tools/profile_group.py 8 > /dev/null 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_png_filter|k_png_match|k_png_hashmatch|k_png_parse|k_png_unfilter|k_dfl_emit|k_dfl_hist|k_dfl_tables" --launch-skip 0 -c 8 -f -o gpurun_out/${TAG}_png python tools/profile_legs.py png > /dev/null 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_vp8_encode|k_vp8_rgb_to_yuv|k_resize|k_lanczos|k_planes_to_rgb|k_ycc" -c 6 -f -o gpurun_out/${TAG}_webp python tools/profile_legs.py webp > /dev/null 2>&1
unset B200_GRAPHS
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], {k: (v.get('value'), v.get('e2e', {}).get('value')) for k, v in d.get('configs', {}).items()})
r=json.load(open('gpurun_out/${TAG}_bench_reference.json')); print('reference', r['value'])"
ls -la gpurun_out | tail -14
# host-side traces of the PNG and WebP legs (per-image stage times under 16 concurrent callers) and the per-image-call probe
B200_TRACE=2 timeout 120 python bench.py --configs 3 --skip-cpu-baseline --steps 5 > gpurun_out/${TAG}_trace_png.json 2> gpurun_out/${TAG}_trace_png.err
B200_TRACE=2 timeout 120 python bench.py --configs 4 --skip-cpu-baseline --steps 6 > gpurun_out/${TAG}_trace_webp.json 2> gpurun_out/${TAG}_trace_webp.err
grep "trace\] png" gpurun_out/${TAG}_trace_png.err | tail -4; grep "trace\] jpeg" gpurun_out/${TAG}_trace_webp.err | tail -4
B200_COALESCE=0 timeout 100 python tools/coalesce_probe.py 1024 16,64 > gpurun_out/${TAG}_coalesce0.json 2>/dev/null; B200_COALESCE=1 timeout 100 python tools/coalesce_probe.py 1024 16,64 > gpurun_out/${TAG}_coalesce1.json 2>/dev/null
cat gpurun_out/${TAG}_coalesce0.json gpurun_out/${TAG}_coalesce1.json
