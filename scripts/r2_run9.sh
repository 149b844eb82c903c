set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=$GRAFT_REPO_ROOT/gpurun_out
for i in 1 2; do
( cd .r1check && timeout 600 python bench.py --workload farn_720p --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('r1 farn', d['value'], d['ms_per_step'], d['e2e']['value'])" )
timeout 600 python bench.py --workload farn_720p --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('r2 farn', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_share_of_step'])"
done
