cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zoo.py -q -k "yolo" 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
HB_BENCH_DETAIL=1 timeout 600 $B --model yolov4 > gpurun_out/g_yolov4.json 2> gpurun_out/g_yolov4.err; echo "yolo rc=$?"; grep -v "Warning\|DETAIL\|warn" gpurun_out/g_yolov4.err | tail -5
python -c "
import json
d=json.load(open('gpurun_out/g_yolov4.json'))
print('yolov4', round(d['ms_per_step'],3), round(d['value'],1), d['config']['launch'], d['host_enqueue_ms_per_step'], {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()}, d['roofline'].get('whole_step_tflops'))
"
