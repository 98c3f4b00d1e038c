cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
timeout 300 $B > gpurun_out/c_a0.json 2> gpurun_out/c_a0.err
for m in rexnet1_0x repvgg_a1 unet3p yolov4; do
  HB_BENCH_DETAIL=1 timeout 400 $B --model $m > gpurun_out/c_$m.json 2> gpurun_out/c_$m.err
  echo "== $m rc=$?"; tail -c 600 gpurun_out/c_$m.err | tail -4
done
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"; tail -3 gpurun_out/micro.err
for f in a0 rexnet1_0x repvgg_a1 unet3p yolov4; do python -c "
import json,sys
try:
  d=json.load(open('gpurun_out/c_$f.json'))
  print('$f', round(d['ms_per_step'],3), round(d['value'],1), d['config']['launch'], {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()}, d['roofline'].get('whole_step_tflops'))
except Exception as e: print('$f', 'ERR', e)
"; done
python -c "
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r)
"
