set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2t_pytest.log; tail -n 4 $O/r2t_pytest.log
python scripts/gpu_probe3.py 340 256 default 64 8 0 2>&1 | tail -1
python scripts/gpu_probe3.py 1920 1080 default 15 8 0 2>&1 | tail -1
timeout 600 python bench.py --workload tvl1_340x256 --list 256 --steps 1 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('list256', d['value'])"
