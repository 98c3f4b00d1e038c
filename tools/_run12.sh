cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_a0.csv python tools/dev_one_step.py repvgg_a0 256 > gpurun_out/ncu_a0.log 2>&1; echo "a0 rc=$?"
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_rex.csv python tools/dev_one_step.py rexnet1_0x 256 > gpurun_out/ncu_rex.log 2>&1; echo "rex rc=$?"
python tools/summarize_launches.py gpurun_out/r02_step_a0.csv gpurun_out/r02_launches_a0.md gpurun_out/r02_traffic.json
python tools/summarize_launches.py gpurun_out/r02_step_rex.csv gpurun_out/r02_launches_rex.md -
