#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2n2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n2_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['config']['strong_scaling'].items(): print(k, v.get('mlups'), v.get('ms_per_step'), v.get('strong_efficiency'))
print(d['config']['slab_bit_identical'])
PY
echo done
