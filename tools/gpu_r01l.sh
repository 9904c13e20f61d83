#!/bin/bash
# 1-GPU pass: tests, bench (folded default + weighted), GAT timing, ncu launch list + full captures (SpMM final, GAT)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests8.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests8.log
python bench.py > gpurun_out/bench8.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench8.log
python bench.py --fold 0 --no-cpu-baseline > gpurun_out/bench8_weighted.log 2>&1; echo "bench-w rc=$?"; tail -1 gpurun_out/bench8_weighted.log | cut -c1-400
python tools/run_gat.py > gpurun_out/gat8.log 2>&1; echo "gat rc=$?"; tail -1 gpurun_out/gat8.log
python tools/run_gat.py --heads 1 --channels 41 > gpurun_out/gat8_1x41.log 2>&1; echo "gat rc=$?"; tail -1 gpurun_out/gat8_1x41.log
python tools/bench_configs.py > gpurun_out/configs8.log 2>&1; echo "configs rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|stage_rows|pack_rows|row_scale" -c 200 --csv \
    --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu8a.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spmm_rows|spmm_long" -s 40 -c 4 -o gpurun_out/prof_spmm_v3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu8b.log 2>&1
echo "ncu spmm rc=$?"
python tools/run_gat.py --iters 1 > gpurun_out/plain8g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gat_" -s 9 -c 9 -o gpurun_out/prof_gat_v2 \
    python tools/run_gat.py --iters 1 > gpurun_out/ncu8c.log 2>&1
echo "ncu gat rc=$?"
