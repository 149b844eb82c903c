set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=4 > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log; tail -n 9 $O/r2z_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.json
timeout 600 python bench.py --workload tvl1_340x256 --list 1024 --steps 1 --warmup 1 > $O/r2_list1024_n1.json 2> $O/r2_list1024_n1.err
timeout 600 python bench.py --workload tvl1_340x256 --steps 5 --warmup 3 --pairs 63 > $O/r2_bench_tvl1_340x256.json 2> $O/r2_bench_340.err
timeout 600 python bench.py --workload farn_720p --steps 10 --warmup 3 > $O/r2_bench_farn_720p.json 2> $O/r2_bench_farn.err
timeout 600 python bench.py --workload tvl1_340x256 --chain --steps 5 --warmup 2 --pairs 63 > $O/r2_chain_340x256.json 2> $O/r2_chain_340.err
timeout 600 python bench.py --workload tvl1_455x256 --chain --chain-src 1920x1080 --steps 5 --warmup 2 --pairs 32 > $O/r2_chain_1080p_ns256.json 2> $O/r2_chain_ns.err
timeout 600 python bench.py --workload tvl1_1080p --chain --steps 3 --warmup 1 --pairs 16 > $O/r2_chain_1080p.json 2> $O/r2_chain_1080.err
ncu --set full --clock-control none --import-source on -k regex:k_tvl1_pair -c 2 -f -o $O/r2_fused_full python scripts/one_pair.py 1920 1080 10 0 > $O/r2_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_box_solve_update_tma -s 90 -c 2 -f -o $O/r2_farn_box_full python scripts/farn_pairs.py 1280 720 9 > $O/r2_ncu_farn.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/r2_farn_launches.csv python scripts/farn_pairs.py 1280 720 9 > $O/r2_farn_launches.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > $O/r2_compute_sanitizer.txt 2>&1; tail -n 4 $O/r2_compute_sanitizer.txt
python scripts/gpu_phase.py 1920 1080 8 1 > $O/r2_phase_l1.log 2>&1; cat $O/r2_phase_l1.log
