# the driver's N>1 launch shape on the final tree: default bench, both arms
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n2.json | head -2 | tr '\n' ' ')"
tail -3 $O/bench_n2.err
