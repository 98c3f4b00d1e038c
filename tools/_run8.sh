cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/dev_frozen_step_diag.py 2>&1 | grep -v Warn > gpurun_out/frozen_diag.log; tail -40 gpurun_out/frozen_diag.log
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_optim.py -q -s -x 2>&1 | grep -v Warning > gpurun_out/r2_trainer1.log
grep -E "^\[trainer|^E  |FAILED|passed|failed|Error" gpurun_out/r2_trainer1.log | cut -c1-500 | head -40
