set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10; nproc
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --workload tvl1_340x256 --list 1024 --steps 1 --warmup 1 > gpurun_out/r2_list1024_n$N.json 2> gpurun_out/r2_list1024_n$N.err
  tail -c 900 gpurun_out/r2_list1024_n$N.json; tail -n 2 gpurun_out/r2_list1024_n$N.err
done
python bench.py --gpus 8 --workload tvl1_340x256 --list 1024 --steps 1 --warmup 1 > gpurun_out/r2_list1024_threads8.json 2> gpurun_out/r2_list1024_threads8.err
tail -c 900 gpurun_out/r2_list1024_threads8.json; tail -n 2 gpurun_out/r2_list1024_threads8.err
