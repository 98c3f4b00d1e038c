cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dev_conv_check.py 2>&1 | grep -E "^time|ALL OK|BAD" > gpurun_out/conv_check.log; cat gpurun_out/conv_check.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
HB_BENCH_DETAIL=1 timeout 400 $B > gpurun_out/e_a0.json 2> gpurun_out/e_a0.err
python -c "
import json
d=json.load(open('gpurun_out/e_a0.json'))
print('a0', round(d['ms_per_step'],3), round(d['value'],1), {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()})
"
