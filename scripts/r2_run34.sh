# final tree: GPU suite, smoke, the default bench line, ncu --set full of the fused kernel (7-pair and 2-pair launches)
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$? $(grep -o '"value": [0-9.]*' $O/r2_bench_n1.json | head -2 | tr '\n' ' ')"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tvl1_pair -c 2 -f -o $O/r2_fused_full python scripts/one_pair.py 1920 1080 10 0 > $O/r2_ncu_fused.log 2>&1; echo "ncu rc=$?"
