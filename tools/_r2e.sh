#!/bin/bash
# fourth run: 32-bit index arithmetic in the multi-tensor pack kernel - timing, then every conv-touching GPU test file, then the bench line
mkdir -p gpurun_out
timeout 120 python tools/micro_pack.py repvgg_a0 32 2>&1 | tail -2 | tee gpurun_out/r2e_pack.log
timeout 300 python -m pytest tests/test_gpu_conv_bn.py tests/test_gpu_conv_composites.py tests/test_gpu_fused_conv.py tests/test_gpu_repvgg.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2e_tests.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-eager-baseline --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; cut -c1-330 gpurun_out/r2e_bench.json
