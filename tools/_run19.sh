cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
timeout 600 $B --model yolov4 > gpurun_out/g_yolov4.json 2> gpurun_out/g_yolov4.err; echo "yolo rc=$?"; grep -v "Warning\|DETAIL\|warn" gpurun_out/g_yolov4.err | grep -B30 "capture failed" | tail -40
