# arithmetic variants of the TV-L1 inner loop (rho as two chained FMAs / two reciprocals instead of the shared one)
cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 120 python scripts/gpu_probe3.py 1920 1080 default,mA,mB,mAB 17 8 0 2>&1 | grep lanes; done
timeout 60 python scripts/gpu_probe3.py 340 256 default,mA,mB,mAB 64 8 0 2>&1 | grep lanes
