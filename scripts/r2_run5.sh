set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_farneback_gpu.py tests/test_png_gpu.py tests/test_tvl1_gpu.py -m gpu -q -x --durations=5 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -n 15 gpurun_out/r2d_pytest.log
python scripts/gpu_probe3.py 1920 1080 default 17 8 1,0 > gpurun_out/r2d_probe3.log 2>&1; cat gpurun_out/r2d_probe3.log
python scripts/farn_probe.py > gpurun_out/r2d_farn_probe.log 2>&1; cat gpurun_out/r2d_farn_probe.log
timeout 600 python bench.py --workload farn_720p --steps 5 --warmup 3 > gpurun_out/r2d_bench_farn.json 2> gpurun_out/r2d_bench_farn.err; tail -c 2500 gpurun_out/r2d_bench_farn.json; tail -n 3 gpurun_out/r2d_bench_farn.err
