#!/bin/bash
# round-1 third GPU pass: GAT v2 + hot-column cache policy A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests5.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests5.log
python tools/sweep.py --workloads products --windows -1 --shapes 8:2:18,16:2:18 --us 18 \
   --policies off,h0c0,h2c1,h2c0,h2c2 --hot-mb 32,64,96 > gpurun_out/sweep5.log 2>&1; echo "sweep rc=$?"
grep BEST gpurun_out/sweep5.log
python tools/bench_configs.py --only c3 > gpurun_out/configs5.log 2>&1; echo "configs rc=$?"; tail -5 gpurun_out/configs5.log
python bench.py --no-cpu-baseline > gpurun_out/bench5.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench5.log
