#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_tests.log; cat gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cat gpurun_out/final_ref.json | cut -c1-700
for m in repvgg_a1 yolov4 unet3p; do timeout 500 python bench.py --model $m --gpus 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/final_$m.json 2> gpurun_out/final_$m.err; done
timeout 600 python bench.py --micro > gpurun_out/final_micro.json 2> gpurun_out/final_micro.err
python - <<'PY'
import json
def last(p):
    return json.loads(open(p).read().strip().split('\n')[-1])
b=last('gpurun_out/final_bench.json')
print('A0', b['ms_per_step'], b['value'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], 'clocks', b['clocks'])
print(' roofline', {k:b['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','algorithmic_bytes')})
print(' fam', {k[:14]:(v['ms'],v['frac']) for k,v in b['roofline']['per_family'].items()})
print(' cpu', b['cpu_baseline']); print(' eager', b.get('gpu_eager_baseline'))
s=b['secondary']; print(' rex', s['ms_per_step'], s['images_per_s'], {k[:14]:(v['ms'],v['frac']) for k,v in s['roofline']['per_family'].items()})
for m in ('repvgg_a1','yolov4','unet3p'):
    try:
        x=last(f'gpurun_out/final_{m}.json'); print(m, x['ms_per_step'], x['value'], 'e2e', x['e2e']['value'], {k[:14]:(v['ms'],v['frac']) for k,v in x['roofline']['per_family'].items()})
    except Exception as e: print(m, 'failed', e)
PY
