#!/bin/bash
# round-1 call Z: ballot-ranked radix passes -- bit-exactness tests, build timing, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graph_build.py tests/test_rowgen.py -q -x 2>&1 | tail -3
for wl in products reddit arxiv; do timeout 100 python tools/build_only.py --workload $wl 2>&1 | tail -1; done
FILT='regex:ee_|rs_|scan_|rowptr_|gather_i32|deg_hist|longrow_|row_key|degree_norm|iota_'
timeout 200 ncu -k "$FILT" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/z_build_launches_v3b.csv python tools/build_only.py --workload products --reps 1 > gpurun_out/z_build_ncu.log 2>&1
echo "ncu rc=$?"
