#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zoo.py -q -k "mobileone" -s 2>&1 | grep -E "zoo eval|zoo train|zoo reparam|passed|failed|Error|error|assert" | tail -12
timeout 500 python bench.py --model mobileone_s0 --gpus 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/m1.json 2> gpurun_out/m1.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/m1.json').read().strip().split('\n')[-1])
    print('mobileone_s0', round(b['ms_per_step'],3), round(b['value'],1), b['config'].get('launch'), {k[:12]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/m1.err').read()[-2500:])
PY
