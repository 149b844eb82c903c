set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_tvl1_gpu.py tests/test_decode_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log; tail -n 6 $O/r2g_pytest.log
python scripts/gpu_probe3.py 1920 1080 default 15 8 1,0 > $O/r2g_probe3.log 2>&1; cat $O/r2g_probe3.log
python scripts/gpu_phase.py 1920 1080 8 1 > $O/r2g_phase_l1.log 2>&1; head -3 $O/r2g_phase_l1.log
python scripts/farn_probe.py > $O/r2g_farn_probe.log 2>&1; cat $O/r2g_farn_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2g_bench_n1.json 2> $O/r2g_bench_n1.err; python -c "
import json;d=json.loads([l for l in open('$O/r2g_bench_n1.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['parity_aee_px'],d['roofline']['frac'])"
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > $O/r2_compute_sanitizer.txt 2>&1; tail -n 4 $O/r2_compute_sanitizer.txt
