cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dev_traj_diag.py 2>&1 | grep -v Warn > gpurun_out/traj_diag.log; tail -5 gpurun_out/traj_diag.log
timeout 900 python -m pytest tests/test_gpu_zoo.py -q -s 2>&1 | grep -v Warning > gpurun_out/r2_zoo2.log
grep -E "^\[zoo|^E  |\[trajectory|FAILED|passed|failed" gpurun_out/r2_zoo2.log | cut -c1-400
