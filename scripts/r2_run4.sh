set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/gpu_probe3.py 1920 1080 default,t256 17 8,6,4 1,0 > gpurun_out/r2c_probe3.log 2>&1
python scripts/gpu_phase.py 1920 1080 8 1 > gpurun_out/r2c_phase_l1.log 2>&1
python scripts/gpu_phase.py 1920 1080 8 0 > gpurun_out/r2c_phase_auto.log 2>&1
cat gpurun_out/r2c_probe3.log
timeout 1200 python -m pytest tests/test_tvl1_gpu.py tests/test_tvl1_kernels_gpu.py tests/test_list_gpu.py tests/test_cpp_host_gpu.py tests/test_quantise_gpu.py tests/test_farneback_gpu.py -m gpu -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -n 25 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 600 python bench.py --workload tvl1_340x256 --list 256 --steps 1 --warmup 1 > gpurun_out/r2c_list256.json 2> gpurun_out/r2c_list256.err
tail -c 1500 gpurun_out/r2c_list256.json; tail -n 5 gpurun_out/r2c_list256.err
