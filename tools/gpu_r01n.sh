#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests10.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests10.log
python tools/run_gat.py > gpurun_out/gat10.log 2>&1; tail -1 gpurun_out/gat10.log
python tools/run_gat.py --heads 1 --channels 41 > gpurun_out/gat10_1x41.log 2>&1; tail -1 gpurun_out/gat10_1x41.log
python tools/bench_configs.py --only c2,c3 > gpurun_out/configs10.log 2>&1; echo "configs rc=$?"; cut -c1-330 gpurun_out/configs10.log
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench10.log 2>&1; tail -1 gpurun_out/bench10.log | cut -c1-200
