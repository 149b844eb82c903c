set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_tvl1_gpu.py -m gpu -q -x > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log; tail -n 4 $O/r2k_pytest.log
python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 > $O/r2k_probe3.log 2>&1; cat $O/r2k_probe3.log
python scripts/gpu_phase.py 1920 1080 8 1 > $O/r2k_phase_l1.log 2>&1; head -3 $O/r2k_phase_l1.log
python scripts/gpu_probe3.py 340 256 default 64 8 0 > $O/r2k_probe3_340.log 2>&1; cat $O/r2k_probe3_340.log
