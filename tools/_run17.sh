cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {  # name, nproc, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps 10 --warmup 3 $3 > gpurun_out/n$2_$1.json 2> gpurun_out/n$2_$1.err
  echo "rc=$? $1 N=$2"
  python -c "
import json
txt=open('gpurun_out/n$2_$1.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
d=json.loads(line[-1])
print('$1 N=$2', round(d['ms_per_step'],3), round(d['value'],1), d['config'].get('allreduce'), d['config']['launch'], 'host', d['host_enqueue_ms_per_step'], 'e2e', round(d['e2e']['value'],1))
"
}
run repvgg_a0 8 ""
run repvgg_a0_noovl 8 "--no-overlap"
run repvgg_a1 8 "--model repvgg_a1"
run yolov4 8 "--model yolov4"
run unet3p 4 "--model unet3p"
