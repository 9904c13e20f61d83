#!/bin/bash
# round-1 call U: full GPU test suite (new: eval memo, PTA patch, build variants), bench.py (+ guarded epoch),
# graph-build A/B + launch list, epoch with / without the eval memo, config timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --durations=8 > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-t0))"; tail -14 gpurun_out/u_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$? t=$((SECONDS-t0))"; cut -c1-2200 gpurun_out/u_bench.json; tail -3 gpurun_out/u_bench.err
for v in 1 2; do
  for wl in products reddit arxiv; do
    RGBMP_BUILD_VARIANT=$v timeout 200 python tools/build_only.py --workload $wl >> gpurun_out/u_build_v$v.log 2>&1
  done
  echo "build variant $v rc=$? t=$((SECONDS-t0))"; grep -E "^\{" gpurun_out/u_build_v$v.log | cut -c1-400
done
RGBMP_EVAL_MEMO_MB=0 timeout 300 python tools/bench_epoch.py --epochs 5 > gpurun_out/u_epoch_products_nomemo.log 2>&1; echo "epoch nomemo rc=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/u_epoch_products_nomemo.log
for v in 2 1; do
  RGBMP_BUILD_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/u_build_launches_v$v.csv python tools/build_only.py --workload products --reps 1 > gpurun_out/u_build_ncu_v$v.log 2>&1
  echo "ncu build v$v rc=$? t=$((SECONDS-t0))"
done
timeout 600 python tools/bench_configs.py --only c2,c3,c4 > gpurun_out/u_configs.log 2>&1; echo "configs rc=$? t=$((SECONDS-t0))"; grep -E "^\{" gpurun_out/u_configs.log | cut -c1-300
