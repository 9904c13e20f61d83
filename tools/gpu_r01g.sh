#!/bin/bash
# N-GPU scaling pass (trimmed): parity on the medium graph, bench push at N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/mgpu_check.py --workload medium > gpurun_out/mgpu_medium_$N.log 2>&1; echo "mgpu medium rc=$?"; tail -1 gpurun_out/mgpu_medium_$N.log
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n$N.log
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --exchange allgather > gpurun_out/bench_n${N}_allgather.log 2>&1; echo "bench-ag rc=$?"; tail -1 gpurun_out/bench_n${N}_allgather.log
