set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# full capture of the fused kernel: 10 frames = 9 pairs = one 7-pair launch + one 2-pair launch
ncu --set full --clock-control none --import-source on -k regex:k_tvl1_pair -c 2 -f -o gpurun_out/r2_fused_full python scripts/one_pair.py 1920 1080 10 0 > gpurun_out/r2_ncu_fused.log 2>&1
# the two stand-alone inner-loop kernels (unfused schedule) at 1080p, skipping the 3 warm-up launches of each
ncu --set full --clock-control none -k regex:k_estimate -s 3 -c 3 -f -o gpurun_out/r2_unfused_u python scripts/unfused_kernels.py 1920 1080 > gpurun_out/r2_ncu_unfused.log 2>&1
ncu --set full --clock-control none -k regex:k_estimate_dual -s 3 -c 3 -f -o gpurun_out/r2_unfused_dual python scripts/unfused_kernels.py 1920 1080 >> gpurun_out/r2_ncu_unfused.log 2>&1
# launch list of one bench step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out | tail -8
