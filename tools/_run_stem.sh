#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_conv_bn.py tests/test_gpu_zoo.py -x -q 2>&1 | tail -5
run() { tag=$1; model=$2; shift; shift; env "$@" timeout 400 python bench.py --model $model --gpus 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/stem_$tag.json 2> gpurun_out/stem_$tag.err; python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/stem_$tag.json').read().strip().split('\n')[-1])
    print('$tag', round(b['ms_per_step'],3), round(b['value'],1), {k[:12]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/stem_$tag.err').read()[-1500:])
PY
}
run rex rexnet1_0x A=1
run yolo yolov4 A=1
run unet unet3p A=1
