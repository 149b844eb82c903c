set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_pytest.log; tail -n 5 $O/r2s_pytest.log
python scripts/farn_probe.py default 1280x720,1920x1080,340x256 2>&1 | tee $O/r2s_farn_probe.log
timeout 600 python bench.py --workload farn_720p --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('farn bench', d['value'], d['e2e']['value'], d['e2e_quantised']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['gpu_launches'])"
