set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 2>&1 | tee $O/r2q_probe3_1080.log
python scripts/gpu_phase.py 1920 1080 8 1 2>&1 | head -2
python scripts/gpu_probe3.py 340 256 default 64 8 0 2>&1 | tail -1
timeout 600 python -m pytest tests/test_tvl1_gpu.py tests/test_cpp_host_gpu.py -m gpu -q -x 2>&1 | tail -n 3
