#!/bin/bash
# third verification run: Trainer classes on the device path, TridentNet tests, ConvNeXt graph capture, then the whole GPU suite
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_trainer.py -q -s -k "trainer_class" 2>&1 | grep -E "^\[trainer|passed|failed|Error|error|assert|^E  |FAILED" | cut -c1-500 > gpurun_out/r2d_trainer.log; tail -25 gpurun_out/r2d_trainer.log
timeout 300 python -m pytest tests/test_gpu_zoo.py -q -s -k "trident" 2>&1 | grep -E "^\[zoo|passed|failed|Error|error|assert|^E  |FAILED" | cut -c1-500 > gpurun_out/r2d_trident.log; tail -12 gpurun_out/r2d_trident.log
echo "== targeted tests done at $(( $(date +%s) - S )) s"
timeout 110 python bench.py --model convnext_tiny --gpus 1 --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-secondary > gpurun_out/r2d_convnext_tiny.json 2> gpurun_out/r2d_convnext_tiny.err
python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/r2d_convnext_tiny.json').read().strip().split('\n')[-1])
    print('convnext_tiny', round(b['ms_per_step'], 2), 'ms', round(b['value'], 1), 'img/s', b['config'].get('launch'), {k[:12]: (v['ms'], v['frac']) for k, v in b['roofline']['per_family'].items()})
except Exception as e:
    print('convnext_tiny failed', e)
print(open('gpurun_out/r2d_convnext_tiny.err').read()[-600:])
PY
echo "== bench done at $(( $(date +%s) - S )) s"
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2d_all.log; cat gpurun_out/r2d_all.log
echo "== full suite done at $(( $(date +%s) - S )) s"
