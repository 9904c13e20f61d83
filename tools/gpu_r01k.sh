#!/bin/bash
# 8-GPU pass: C5 (papers100M-shaped bf16) on 4x2 and 2x4 grids, products bench with the default grid
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for pf in 2 4; do
timeout 600 $TR --master-port 2953$pf tools/bench_c5.py --scale 1.0 --feature-groups $pf > gpurun_out/c5_n${N}_pf$pf.log 2>&1; echo "c5 pf=$pf rc=$?"; tail -1 gpurun_out/c5_n${N}_pf$pf.log
done
timeout 600 $TR --master-port 29541 tools/bench_c5.py --scale 1.0 --feature-groups 2 --locality 0.9 > gpurun_out/c5_n${N}_pf2_loc09.log 2>&1; echo "c5 loc rc=$?"; tail -1 gpurun_out/c5_n${N}_pf2_loc09.log
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_auto.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n${N}_auto.log
