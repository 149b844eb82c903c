set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 2>&1 | tee $O/r2p_probe3_1080.log
python scripts/gpu_probe3.py 340 256 default 64 8 16,0 2>&1 | tee $O/r2p_probe3_340.log
python scripts/gpu_probe3.py 640 360 default 33 8 0 2>&1 | tail -1
python scripts/gpu_probe3.py 1280 720 default 17 8 0 2>&1 | tail -1
timeout 600 python -m pytest tests/test_tvl1_gpu.py tests/test_list_gpu.py -m gpu -q -x 2>&1 | tail -n 3
