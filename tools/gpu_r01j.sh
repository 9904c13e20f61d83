#!/bin/bash
# N-GPU pass: products bench with Pf in {1,2}; C5 (papers100M-shaped bf16) with Pf in {1,2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for pf in 1 2; do
timeout 600 $TR --master-port 2952$pf bench.py --gpus $N --steps 10 --warmup 3 --feature-groups $pf > gpurun_out/bench_n${N}_pf$pf.log 2>&1; echo "bench pf=$pf rc=$?"; tail -1 gpurun_out/bench_n${N}_pf$pf.log | cut -c1-330
done
for pf in 1 2; do
timeout 900 $TR --master-port 2953$pf tools/bench_c5.py --scale 1.0 --feature-groups $pf > gpurun_out/c5_n${N}_pf$pf.log 2>&1; echo "c5 pf=$pf rc=$?"; tail -1 gpurun_out/c5_n${N}_pf$pf.log
done
