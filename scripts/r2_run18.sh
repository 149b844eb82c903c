set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
python scripts/gpu_probe3.py 340 256 default 64 8 16,0,24,32,48,63 2>&1 | tee $O/r2o_probe3_340.log
python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 2>&1 | tee $O/r2o_probe3_1080.log
python scripts/gpu_probe3.py 256 256 default 64 8 16,0 2>&1 | tail -2
timeout 600 python -m pytest tests/test_tvl1_gpu.py tests/test_list_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -n 3
timeout 600 python bench.py --workload tvl1_340x256 --list 256 --steps 1 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('list256', d['value'], d['list'])"
