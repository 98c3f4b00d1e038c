cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zoo.py -q -s 2>&1 | grep -v Warning | tail -70 > gpurun_out/r2_zoo1.log
tail -45 gpurun_out/r2_zoo1.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zoo.py 2>&1 | tail -5
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
HB_BENCH_DETAIL=1 timeout 400 $B --model yolov4 > gpurun_out/c_yolov4.json 2> gpurun_out/c_yolov4.err; echo "yolo rc=$?"; grep -v Warning gpurun_out/c_yolov4.err | tail -3
python -c "
import json
d=json.load(open('gpurun_out/c_yolov4.json'))
print('yolov4', round(d['ms_per_step'],3), round(d['value'],1), d['config']['launch'], {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()}, d['roofline'].get('whole_step_tflops'))
"
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r['kernel'], r['ms'], r['GB/s'], r['frac_hbm'])
"
