set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/farn_probe.py default 1280x720,1920x1080,340x256 > gpurun_out/r2j_farn_probe.log 2>&1; cat gpurun_out/r2j_farn_probe.log
timeout 300 python -m pytest tests/test_farneback_gpu.py -m gpu -q -x 2>&1 | tail -n 3
