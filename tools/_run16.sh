cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dev_conv_check.py 2>&1 | grep -E "^time|ALL OK|BAD" > gpurun_out/conv_check2.log; cat gpurun_out/conv_check2.log
timeout 900 python -m pytest tests/test_gpu_nn_layers.py tests/test_gpu_fused_conv.py tests/test_gpu_conv_bn.py -q -x 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-eager-baseline"
timeout 400 $B > gpurun_out/f_a0.json 2> gpurun_out/f_a0.err
python -c "
import json
d=json.load(open('gpurun_out/f_a0.json'))
print('a0', round(d['ms_per_step'],3), round(d['value'],1), {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()})
"
