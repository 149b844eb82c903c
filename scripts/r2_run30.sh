# chained vs free fused launches of two list workers on one GPU; refreshed bench lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
one() { # serial try
  DFB_TVL1_SERIAL_LAUNCHES=$1 timeout 90 python bench.py --workload tvl1_340x256 --list 256 --steps 2 --warmup 1 --workers-per-gpu 2 > gpurun_out/ser$1_$2.out 2> gpurun_out/ser$1_$2.err
  echo "serial=$1 try $2 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ser$1_$2.out | head -1) $(grep -i 'watchdog' gpurun_out/ser$1_$2.err | head -1)"
}
one 1 1; one 0 1; one 1 2; one 0 2; one 0 3
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_n1.json | head -1)"
timeout 400 python bench.py --workload farn_720p > gpurun_out/bench_farn_720p.json 2> gpurun_out/bench_farn.err; echo "farn rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_farn_720p.json | head -1)"
timeout 300 python bench.py --workload tvl1_340x256 --no-cpu-baseline > gpurun_out/bench_tvl1_340x256.json 2> gpurun_out/bench_340.err; echo "340 rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_tvl1_340x256.json | head -1)"
