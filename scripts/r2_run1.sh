set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python scripts/gpu_probe3.py 1920 1080 default 17 > gpurun_out/r2a_probe3.log 2>&1
python scripts/gpu_phase.py 1920 1080 8 1 > gpurun_out/r2a_phase_l1.log 2>&1
python scripts/gpu_phase.py 1920 1080 8 0 > gpurun_out/r2a_phase_auto.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -n 3 gpurun_out/r2a_probe3.log; tail -n 5 gpurun_out/r2a_pytest.log
