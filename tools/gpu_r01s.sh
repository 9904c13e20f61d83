#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/mgpu_check.py --workload medium --feature-groups 2 > gpurun_out/mgpu2_medium_${N}_pf2.log 2>&1; echo "mgpu medium 4x2 rc=$?"; grep -E "^\{" gpurun_out/mgpu2_medium_${N}_pf2.log
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench2_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench2_n$N.log | cut -c1-330
timeout 900 $TR --master-port 29531 tools/bench_c5.py --scale 1.0 --feature-groups 2 > gpurun_out/c5b_n${N}_pf2.log 2>&1; echo "c5 rc=$?"; tail -1 gpurun_out/c5b_n${N}_pf2.log
