cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zoo.py -q -s 2>&1 | grep -v Warning > gpurun_out/r2_zoo3.log
grep -E "^\[zoo|^E  |\[trajectory|FAILED|passed|failed" gpurun_out/r2_zoo3.log | cut -c1-600
