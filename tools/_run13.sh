cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
for m in repvgg_a0 rexnet1_0x; do
timeout 400 $B --model $m > gpurun_out/d_$m.json 2> gpurun_out/d_$m.err
python -c "
import json
d=json.load(open('gpurun_out/d_$m.json'))
print('$m', round(d['ms_per_step'],3), round(d['value'],1), {k[:10]:(v['ms'],v['frac']) for k,v in d['roofline']['per_family'].items()})
"
done
