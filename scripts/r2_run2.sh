set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/gpu_probe3.py 1920 1080 t256 17 4,3,5,6 1,0,8,12,16 > gpurun_out/r2b_probe3_t256.log 2>&1
python scripts/gpu_phase.py 1920 1080 4 1 t256 > gpurun_out/r2b_phase_t256_l1.log 2>&1
python scripts/gpu_phase.py 1920 1080 4 0 t256 > gpurun_out/r2b_phase_t256_auto.log 2>&1
python scripts/gpu_probe3.py 340 256 t256,default 64 4,8 0 > gpurun_out/r2b_probe3_340.log 2>&1
cat gpurun_out/r2b_probe3_t256.log gpurun_out/r2b_probe3_340.log
