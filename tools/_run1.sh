cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
HB_BENCH_DETAIL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_pytest1.log; head -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
