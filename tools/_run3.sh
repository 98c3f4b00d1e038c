cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
HB_BENCH_DETAIL=1 timeout 300 $B > gpurun_out/ab_f.json 2> gpurun_out/ab_f.err
HB_BENCH_DETAIL=1 HB_DISABLE_BN_OUT_STATS=1 timeout 300 $B > gpurun_out/ab_g.json 2> gpurun_out/ab_g.err
HB_BENCH_DETAIL=1 HB_FORCE_CONV_STATS=1 timeout 300 $B > gpurun_out/ab_h.json 2> gpurun_out/ab_h.err
HB_BENCH_DETAIL=1 HB_FUSED_FPROP=1 timeout 300 $B > gpurun_out/ab_i.json 2> gpurun_out/ab_i.err
HB_BENCH_DETAIL=1 HB_DISABLE_BN_OUT_STATS=1 HB_DISABLE_CONV_STATS=1 timeout 300 $B --no-direct-grads > gpurun_out/ab_j.json 2> gpurun_out/ab_j.err
for f in f g h i j; do python -c "
import json,sys
d=json.load(open('gpurun_out/ab_$f.json'))
print('$f', round(d['ms_per_step'],3), {k[:10]:v['ms'] for k,v in d['roofline']['per_family'].items()})
"; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
