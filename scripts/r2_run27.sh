cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
one() { # lanes workers
  echo "== lanes=$1 workers=$2"
  DFB_TVL1_LANES=$1 timeout 70 python bench.py --workload tvl1_340x256 --list 128 --steps 1 --warmup 1 --workers-per-gpu $2 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value', d['value'])"
}
one 0 1
one 37 2
one 16 2
one 32 2
