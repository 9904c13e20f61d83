#!/bin/bash
# 8-GPU attribution of the hop time on the 4x2 grid: full / no remote pushes / no inter-hop tick / neither
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $TR --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/attr_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/attr_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")"
}
run full A=1
run selfpush RGBMP_DEBUG_PUSH=self
run notick RGBMP_DEBUG_TICK=0
run neither RGBMP_DEBUG_PUSH=self RGBMP_DEBUG_TICK=0
