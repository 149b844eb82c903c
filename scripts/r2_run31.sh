# validation of the final tree (proxy fence, partials, watchdog) + refreshed evidence
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1_quick.json 2>$O/bench_n1_quick.err; echo "tvl1 1080p $(grep -o '"value": [0-9.]*' $O/bench_n1_quick.json | head -2 | tr '\n' ' ')"
timeout 300 python bench.py --workload tvl1_340x256 > $O/bench_tvl1_340x256.json 2>$O/bench_340.err; echo "340 $(grep -o '"value": [0-9.]*' $O/bench_tvl1_340x256.json | head -2 | tr '\n' ' ')"
timeout 200 python bench.py --workload tvl1_340x256 --list 1024 --steps 2 --warmup 1 > $O/list1024_n1.json 2>$O/list1024_n1.err; echo "list $(grep -o '"value": [0-9.]*' $O/list1024_n1.json | head -1)"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/r2_farn_launches.csv python scripts/farn_pairs.py 1280 720 9 > $O/r2_farn_launches.log 2>&1; echo "ncu farn rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches_bench.log 2>&1; echo "ncu tvl1 rc=$?"
