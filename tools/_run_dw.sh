#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn_layers.py tests/test_gpu_zoo.py -x -q -k "depthwise or rexnet or frelu" 2>&1 | tail -5
timeout 400 python bench.py --model rexnet1_0x --gpus 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/rex_dw.json 2> gpurun_out/rex_dw.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/rex_dw.json').read().strip().split('\n')[-1])
print('rexnet', round(b['ms_per_step'],3), round(b['value'],1), {k[:12]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
PY
