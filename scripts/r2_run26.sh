set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python bench.py --workload tvl1_340x256 --list 256 --steps 1 --warmup 1 2>$O/r2u_list.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('list256', d['value'], d['list'])"
echo "list rc=$?"; tail -n 3 $O/r2u_list.err
timeout 300 python -m pytest tests/test_list_gpu.py tests/test_tvl1_gpu.py -m gpu -q -x 2>&1 | tail -n 3
