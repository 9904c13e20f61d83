#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/sweep.py --workloads products --windows -1 --shapes 8:2:18,16:2:18 --us 18 \
   --policies off,h0c0,h2c1,h2c0,h0c1 --hot-mb 32,64,96 > gpurun_out/sweep7.log 2>&1; echo "sweep rc=$?"
grep -E "BEST" gpurun_out/sweep7.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)['BEST']; print(d['F'],d['weighted'],d['policy'],d['hot_mb'],d['ms'])"
python tools/sweep.py --workloads products --windows -1 --shapes 8:2:18,16:2:18 --us 18 \
   --policies h2c1,h2c0 --hot-mb 32,64 --persist-mb 128 > gpurun_out/sweep7_persist.log 2>&1; echo "sweep rc=$?"
grep -E "BEST" gpurun_out/sweep7_persist.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)['BEST']; print('persist',d['F'],d['weighted'],d['policy'],d['hot_mb'],d['ms'])"
