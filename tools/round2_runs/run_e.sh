mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2m_pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/r2m_pytest_gpu.log | tail -1
timeout 400 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>> gpurun_out/r2m_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2m_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], {k: (v.get('value'), v.get('e2e', {}).get('value')) for k, v in d.get('configs', {}).items()})
print({k:(round(v['ms'],3),v['launches']) for k,v in d['configs']['3']['device']['roofline']['all_kernels'].items() if k.startswith('k_png')})
r=json.load(open('gpurun_out/r2m_bench_reference.json')); print('reference', r['value'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
