set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
for pf in 0 1 0 1; do DFB_PREFETCH=$pf python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 2>&1 | sed "s/^/prefetch=$pf /"; done > $O/r2l_probe3.log; cat $O/r2l_probe3.log
DFB_PREFETCH=1 python scripts/gpu_probe3.py 340 256 default 64 8 0 2>&1 | tail -1
DFB_PREFETCH=0 python scripts/gpu_probe3.py 340 256 default 64 8 0 2>&1 | tail -1
timeout 600 python -m pytest tests/test_tvl1_gpu.py -m gpu -q -x 2>&1 | tail -n 3
