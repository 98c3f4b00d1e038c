cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nn_layers.py tests/test_gpu_trainer.py -q -x 2>&1 | tail -5
timeout 600 python bench.py --micro > gpurun_out/micro.json 2> gpurun_out/micro.err; echo "micro rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/micro.json'))
for r in d['rows']: print(r['kernel'], r['ms'], r['GB/s'], r['frac_hbm'], r.get('TFLOP/s'))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop_kernel -c 4 -o gpurun_out/r02_fprop_full python tools/dev_tensor_bound_profile.py > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_fprop_full.ncu-rep
