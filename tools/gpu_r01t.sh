#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
python tools/bench_epoch.py --workload medium --epochs 3 > gpurun_out/epoch_medium_1.log 2>&1; echo "epoch medium 1gpu rc=$?"; tail -1 gpurun_out/epoch_medium_1.log
python tools/bench_epoch.py --epochs 5 > gpurun_out/epoch_products_1.log 2>&1; echo "epoch products 1gpu rc=$?"; tail -1 gpurun_out/epoch_products_1.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 tools/bench_epoch.py --workload medium --epochs 3 --check > gpurun_out/epoch_medium_$N.log 2>&1; echo "epoch medium rc=$?"; tail -3 gpurun_out/epoch_medium_$N.log | cut -c1-600
timeout 300 $TR --master-port 29572 tools/bench_epoch.py --workload medium --epochs 3 --check --feature-groups 2 > gpurun_out/epoch_medium_${N}_pf2.log 2>&1; echo "epoch medium pf2 rc=$?"; tail -3 gpurun_out/epoch_medium_${N}_pf2.log | cut -c1-600
timeout 600 $TR --master-port 29573 tools/bench_epoch.py --epochs 5 --check > gpurun_out/epoch_products_$N.log 2>&1; echo "epoch products rc=$?"; tail -1 gpurun_out/epoch_products_$N.log | cut -c1-600
