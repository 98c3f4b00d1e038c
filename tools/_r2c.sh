#!/bin/bash
# second verification run: re-checks the tests adjusted after _r2b.sh, the TridentNet / PyConvResNet GPU tests, and times the new families
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_zoo.py -q -s -k "mobileone or yolov1 or f3b" 2>&1 | grep -E "^\[zoo|zoo (eval|train|reparam)|passed|failed|Error|error|assert|^E  |FAILED" | cut -c1-700 > gpurun_out/r2c_tests.log
tail -40 gpurun_out/r2c_tests.log
echo "== tests done at $(( $(date +%s) - S )) s"
for m in res2net50_26w_4s sknet50 convnext_tiny tridentnet50; do
  timeout 110 python bench.py --model $m --gpus 1 --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-secondary > gpurun_out/r2c_$m.json 2> gpurun_out/r2c_$m.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/r2c_$m.json').read().strip().split('\n')[-1])
    print('$m', round(b['ms_per_step'], 2), 'ms', round(b['value'], 1), 'img/s', b['config'].get('launch'), {k[:12]: (v['ms'], v['frac']) for k, v in b['roofline']['per_family'].items()})
except Exception as e:
    print('$m failed', e); print(open('gpurun_out/r2c_$m.err').read()[-1500:])
PY
done
echo "== bench done at $(( $(date +%s) - S )) s"
