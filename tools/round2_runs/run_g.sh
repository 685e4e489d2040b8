mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_n2.json 2> gpurun_out/r2o_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2o_n2.json')); print('n2 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['images_per_sec'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])"
timeout 200 $TR 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2o_n2_reference.json 2>> gpurun_out/r2o_n2.err
cat gpurun_out/r2o_n2_reference.json | cut -c1-300
tail -3 gpurun_out/r2o_n2.err
