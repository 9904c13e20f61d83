#!/bin/bash
# round-1 call Y: ncu --set full of the radix passes of the graph build (one launch of each variant of the scatter + the histogram)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 250 ncu --set full --clock-control none --import-source on -k regex:'rs_scatter3|rs_hist' -c 4 -o gpurun_out/y_build_full -f \
    python tools/build_only.py --workload products --reps 1 > gpurun_out/y_build_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/y_build_full.ncu-rep
