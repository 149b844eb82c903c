set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_farneback_gpu.py tests/test_pipeline_gpu.py tests/test_list_gpu.py -m gpu -q -x > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log; tail -n 5 $O/r2h_pytest.log
python scripts/farn_probe.py > $O/r2h_farn_probe.log 2>&1; cat $O/r2h_farn_probe.log
for i in 1 2; do
timeout 600 python bench.py --workload farn_720p --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('r2 farn', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'])"
done
