#!/bin/bash
# round-1 call V: per-kernel launch list of the graph build (both variants), epoch after the memo generation fix
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
FILT='regex:ee_|rs_|scan_|rowptr_|gather_i32|deg_hist|longrow_|row_key|degree_norm|iota_'
for v in 2 1; do
  RGBMP_BUILD_VARIANT=$v timeout 200 ncu -k "$FILT" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/v_build_launches_v$v.csv python tools/build_only.py --workload products --reps 1 > gpurun_out/v_build_ncu_v$v.log 2>&1
  echo "ncu build v$v rc=$?"
done
RGBMP_EPOCH_TRACE=1 timeout 200 python tools/bench_epoch.py --epochs 5 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_eval_memo.py -q 2>&1 | tail -2
