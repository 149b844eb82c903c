set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/farn_probe.py default 1280x720,1920x1080,340x256 2>&1 | tee gpurun_out/r2r_farn_probe.log
timeout 600 python -m pytest tests/test_farneback_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -n 3
