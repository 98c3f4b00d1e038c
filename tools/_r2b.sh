#!/bin/bash
# one-shot verification run of the second session of round 2 (GPU budget: ~10 minutes of box time)
mkdir -p gpurun_out
S=$(date +%s)
timeout 420 python -m pytest tests/test_gpu_zoo.py -q -s -k "f3 or yolov1 or mobileone" 2>&1 | grep -E "^\[zoo|zoo (eval|train|reparam)|passed|failed|Error|error|assert|^E  |FAILED|PASSED" | cut -c1-600 > gpurun_out/r2b_new_tests.log
tail -60 gpurun_out/r2b_new_tests.log
echo "== new tests done at $(( $(date +%s) - S )) s"
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_zoo.py -x 2>&1 | tail -5 > gpurun_out/r2b_all_tests.log; cat gpurun_out/r2b_all_tests.log
echo "== other gpu test files done at $(( $(date +%s) - S )) s"
timeout 300 python -m pytest tests/test_gpu_zoo.py -q -k "not f3 and not yolov1 and not mobileone" 2>&1 | tail -5 > gpurun_out/r2b_zoo_tests.log; cat gpurun_out/r2b_zoo_tests.log
echo "== zoo tests done at $(( $(date +%s) - S )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python bench.py --gpus 1 --steps 10 --warmup 3 --no-secondary > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 400 gpurun_out/r2b_bench.err; cut -c1-900 gpurun_out/r2b_bench.json
echo "== bench done at $(( $(date +%s) - S )) s"
