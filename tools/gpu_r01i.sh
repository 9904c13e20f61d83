#!/bin/bash
# 2-GPU pass: full GPU test suite, grid parity (1x2), C5 papers100M-shaped bf16 at P=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests7.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests7.log
python tools/bench_c5.py --scale 0.05 > gpurun_out/c5_1gpu_s005.log 2>&1; echo "c5 1gpu rc=$?"; tail -1 gpurun_out/c5_1gpu_s005.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/mgpu_check.py --workload medium --feature-groups 2 > gpurun_out/mgpu_medium_2_pf2.log 2>&1; echo "mgpu 1x2 rc=$?"; grep -E "^\{" gpurun_out/mgpu_medium_2_pf2.log
timeout 900 $TR --master-port 29531 tools/bench_c5.py --scale 1.0 > gpurun_out/c5_n2.log 2>&1; echo "c5 n2 rc=$?"; tail -1 gpurun_out/c5_n2.log
timeout 900 $TR --master-port 29532 tools/bench_c5.py --scale 1.0 --exchange allgather > gpurun_out/c5_n2_allgather.log 2>&1; echo "c5 n2 ag rc=$?"; tail -1 gpurun_out/c5_n2_allgather.log
timeout 900 $TR --master-port 29533 tools/bench_c5.py --scale 1.0 --feature-groups 2 > gpurun_out/c5_n2_pf2.log 2>&1; echo "c5 n2 pf2 rc=$?"; tail -1 gpurun_out/c5_n2_pf2.log
