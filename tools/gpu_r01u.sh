#!/bin/bash
# round-1 call U: full GPU test suite (new: eval memo, PTA patch, build variants), graph-build A/B + launch list,
# bench.py, epoch with / without the eval memo, config timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/u_pytest_gpu.log
for v in 1 2; do
  for wl in products reddit arxiv; do
    RGBMP_BUILD_VARIANT=$v timeout 300 python tools/build_only.py --workload $wl >> gpurun_out/u_build_v$v.log 2>&1
  done
  echo "build variant $v rc=$?"; grep -E "^\{" gpurun_out/u_build_v$v.log | cut -c1-400
done
for v in 1 2; do
  RGBMP_BUILD_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/u_build_launches_v$v.csv python tools/build_only.py --workload products --reps 1 > gpurun_out/u_build_ncu_v$v.log 2>&1
  echo "ncu build v$v rc=$?"
done
timeout 900 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/u_bench.json
timeout 600 python tools/bench_epoch.py --epochs 5 > gpurun_out/u_epoch_products_memo.log 2>&1; echo "epoch memo rc=$?"; tail -1 gpurun_out/u_epoch_products_memo.log
RGBMP_EVAL_MEMO_MB=0 timeout 600 python tools/bench_epoch.py --epochs 5 > gpurun_out/u_epoch_products_nomemo.log 2>&1; echo "epoch nomemo rc=$?"; tail -1 gpurun_out/u_epoch_products_nomemo.log
timeout 900 python tools/bench_configs.py --only c2,c3,c4 > gpurun_out/u_configs.log 2>&1; echo "configs rc=$?"; grep -E "^\{" gpurun_out/u_configs.log | cut -c1-300
