cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py -q -s 2>&1 | grep -v Warning > gpurun_out/r2_trainer3.log
grep -E "^\[trainer|^E  |FAILED|passed|failed|Error" gpurun_out/r2_trainer3.log | cut -c1-400 | head -30
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_trainer.py 2>&1 | tail -4
