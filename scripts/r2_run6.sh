set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/farn_tma_debug.py default > gpurun_out/r2e_dbg_default.log 2>&1; tail -n 3 gpurun_out/r2e_dbg_default.log
python scripts/farn_tma_debug.py hx6 > gpurun_out/r2e_dbg_hx6.log 2>&1; tail -n 3 gpurun_out/r2e_dbg_hx6.log
timeout 300 compute-sanitizer --tool memcheck python scripts/farn_tma_debug.py hx6 > gpurun_out/r2e_san_hx6.log 2>&1; grep -v "^$" gpurun_out/r2e_san_hx6.log | head -40
timeout 900 python -m pytest tests/test_farneback_gpu.py tests/test_png_gpu.py -m gpu -q -x --durations=5 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -n 12 gpurun_out/r2e_pytest.log
python scripts/farn_probe.py > gpurun_out/r2e_farn_probe.log 2>&1; cat gpurun_out/r2e_farn_probe.log
timeout 600 python bench.py --workload farn_720p --steps 5 --warmup 3 > gpurun_out/r2e_bench_farn.json 2> gpurun_out/r2e_bench_farn.err; tail -c 3000 gpurun_out/r2e_bench_farn.json; tail -n 3 gpurun_out/r2e_bench_farn.err
