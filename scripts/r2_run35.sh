# last validation of the round: GPU suite + smoke + quick bench on the committed tree
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_last.json 2>$O/bench_last.err; echo "bench rc=$? $(grep -o '"value": [0-9.]*' $O/bench_last.json | head -2 | tr '\n' ' ')"
