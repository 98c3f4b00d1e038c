#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_trainer.py tests/test_gpu_nn_layers.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err
tail -3 gpurun_out/micro.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r['kernel'], r['ms'], r['frac_hbm'])
PY
