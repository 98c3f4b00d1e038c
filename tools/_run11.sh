cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in "" "--no-overlap"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $mode > gpurun_out/n2$mode.json 2> gpurun_out/n2$mode.err
echo "rc=$? mode=$mode"; tail -c 400 gpurun_out/n2$mode.err | tail -3
python -c "
import json
d=json.load(open('gpurun_out/n2$mode.json'))
print('N=2 $mode', round(d['ms_per_step'],3), round(d['value'],1), d['config']['allreduce'], d['host_enqueue_ms_per_step'], d['e2e']['value'])
"
done
