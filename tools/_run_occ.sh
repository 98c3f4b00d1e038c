#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --model rexnet1_0x --gpus 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/rex_$tag.json 2> gpurun_out/rex_$tag.err; grep "\[hb\]" gpurun_out/rex_$tag.err | sort | uniq -c | head -30; python - <<PY
import json
b=json.loads(open('gpurun_out/rex_$tag.json').read().strip().split('\n')[-1])
print('$tag', b['ms_per_step'], {k[:12]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
PY
}
run default A=1
run fwd3 HB_BN_CAP_FWD=3
run occ HB_BN_USE_OCC=1 HB_BN_DEBUG=1
run noquad HB_DISABLE_DW_QUAD=1
