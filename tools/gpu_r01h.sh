#!/bin/bash
# 8-GPU grid pass: parity of 4x2 / 2x4 grids on the medium graph, bench with Pf in {1,2,4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for pf in 2 4; do
timeout 300 $TR --master-port 2951$pf tools/mgpu_check.py --workload medium --feature-groups $pf > gpurun_out/mgpu_medium_${N}_pf$pf.log 2>&1; echo "mgpu medium pf=$pf rc=$?"; tail -1 gpurun_out/mgpu_medium_${N}_pf$pf.log
done
for pf in 2 4 8; do
timeout 600 $TR --master-port 2952$pf bench.py --gpus $N --steps 10 --warmup 3 --feature-groups $pf > gpurun_out/bench_n${N}_pf$pf.log 2>&1; echo "bench pf=$pf rc=$?"; tail -1 gpurun_out/bench_n${N}_pf$pf.log | cut -c1-330
done
