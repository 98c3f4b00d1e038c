cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --no-eager-baseline"
HB_BENCH_DETAIL=1 timeout 300 $B > gpurun_out/ab_a.json 2> gpurun_out/ab_a.err
HB_BENCH_DETAIL=1 HB_DISABLE_FUSED_FPROP=1 timeout 300 $B > gpurun_out/ab_b.json 2> gpurun_out/ab_b.err
HB_BENCH_DETAIL=1 HB_DISABLE_FUSED_FPROP=1 HB_DISABLE_CONV_STATS=1 timeout 300 $B > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err
HB_BENCH_DETAIL=1 HB_DISABLE_CONV_STATS=1 timeout 300 $B > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err
HB_BENCH_DETAIL=1 HB_DISABLE_FUSED_FPROP=1 HB_DISABLE_CONV_STATS=1 HB_DISABLE_BN_OUT_STATS=1 timeout 300 $B > gpurun_out/ab_e.json 2> gpurun_out/ab_e.err
for f in a b c d e; do python -c "
import json,sys
d=json.load(open('gpurun_out/ab_$f.json'))
print('$f', round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_family'].items()})
"; done
