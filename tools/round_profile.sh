#!/bin/bash
# Round-end evidence run (on a GPU box): GPU test suite, both bench arms, launch lists and ncu --set full captures.
# Everything lands in gpurun_out/; tools/summarise_profiles.py cuts the committed summaries from it.
TAG=${1:-r1c}
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -1
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --e2e-batch 16 --skip-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_group.csv python tools/profile_group.py 8 > /dev/null 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_fused_same|k_chroma420_refdct|k_idct_plane" -c 3 -f -o gpurun_out/${TAG}_transform python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --e2e-batch 16 > /dev/null 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_geb_emit|k_geb_hist|k_geb_len|k_geb_classify|k_gd_write|k_gd_round0|k_ge_tables" -c 7 -f -o gpurun_out/${TAG}_entropy python tools/profile_group.py 8 > /dev/null 2>&1
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['roofline']['all_kernels'])
r=json.load(open('gpurun_out/${TAG}_bench_reference.json')); print('reference', r['value'])"
ls -la gpurun_out | tail -12
