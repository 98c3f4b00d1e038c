#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pointwise_losses_boxes.py tests/test_gpu_nn_layers.py tests/test_gpu_conv_bn.py tests/test_gpu_fused_conv.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err
tail -3 gpurun_out/micro.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r['kernel'], r['ms'], r['frac_hbm'])
PY
HB_BN_DEBUG=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eager-baseline > gpurun_out/bench_a0.json 2> gpurun_out/bench_a0.err
grep "\[hb\]" gpurun_out/bench_a0.err | sort | uniq -c
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench_a0.json').read().strip().split('\n')[-1])
print(b['ms_per_step'], b['value'], {k[:12]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
s=b['secondary']; print(s['ms_per_step'], s['images_per_s'], {k[:12]:(v['ms'],v['frac']) for k,v in s['roofline']['per_family'].items()})
PY
