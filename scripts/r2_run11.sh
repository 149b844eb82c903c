set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/farn_probe.py default,fb4,fb3 1280x720,1920x1080 > gpurun_out/r2i_farn_probe.log 2>&1; cat gpurun_out/r2i_farn_probe.log
