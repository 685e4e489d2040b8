mkdir -p gpurun_out
for t in 16 24 32 48; do
  timeout 120 python bench.py --configs 3 --only-configs --skip-cpu-baseline --steps 10 --warmup 3 --png-threads $t --png-batch $((t*2)) > gpurun_out/r2n_png_t$t.json 2> gpurun_out/r2n_png_t$t.err
  python -c "
import json
d=json.load(open('gpurun_out/r2n_png_t$t.json')); c=d['configs']['3']
print('threads $t png e2e', c['e2e']['value'], 'MP/s', c['e2e']['images_per_sec'], 'img/s; steps', c['e2e']['steps'], 'images/step', c['e2e']['images_per_step'])"
done
