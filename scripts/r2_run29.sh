# after double-buffering the convergence partials: the stalling list case x3, the 1024-clip list, the full GPU suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do
  echo "== auto lanes, 2 workers, list 256, try $i"
  timeout 90 python bench.py --workload tvl1_340x256 --list 256 --steps 1 --warmup 1 --workers-per-gpu 2 > gpurun_out/l256_$i.out 2> gpurun_out/l256_$i.err
  echo "  rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/l256_$i.out | head -1; grep -i "watchdog\|error" gpurun_out/l256_$i.err | head -3
done
echo "== list 1024"
timeout 200 python bench.py --workload tvl1_340x256 --list 1024 --steps 2 --warmup 1 --workers-per-gpu 2 > gpurun_out/list1024_n1.json 2> gpurun_out/list1024_n1.err
echo "  rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/list1024_n1.json | head -1
echo "== pytest"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
