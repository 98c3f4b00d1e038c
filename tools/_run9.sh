cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py -q -s 2>&1 | grep -v Warning > gpurun_out/r2_trainer2.log
grep -E "^\[trainer|^E  |FAILED|passed|failed|Error" gpurun_out/r2_trainer2.log | cut -c1-600 | head -40
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_trainer.py 2>&1 | tail -8
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r['kernel'], r['ms'], r['GB/s'], r['frac_hbm'])
"
