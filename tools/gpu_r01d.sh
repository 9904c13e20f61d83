#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests6.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests6.log
python tools/proto_colblock.py > gpurun_out/colblock.log 2>&1; echo "colblock rc=$?"; cat gpurun_out/colblock.log | tail -20
python tools/sweep.py --workloads products --windows -1 --shapes 8:2:18,16:2:18 --us 18 \
   --policies h2c1,h2c0 --hot-mb 32,64 --persist-mb 128 > gpurun_out/sweep6_persist.log 2>&1; echo "sweep rc=$?"
grep -E "BEST|persisting" gpurun_out/sweep6_persist.log
