#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests9.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests9.log
python bench.py > gpurun_out/bench9.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench9.log | cut -c1-300
python bench.py --fold 0 --no-cpu-baseline > gpurun_out/bench9_weighted.log 2>&1; echo "bench-w rc=$?"; tail -1 gpurun_out/bench9_weighted.log | cut -c1-300
python tools/sweep.py --workloads products --windows -1 --shapes 8:2:18,16:2:18 --us 18 --policies off,default > gpurun_out/sweep9.log 2>&1
grep BEST gpurun_out/sweep9.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)['BEST']; print(d['F'],d['weighted'],d['policy'],d['hot_mb'],d['ms'])"
python tools/run_gat.py > gpurun_out/gat9.log 2>&1; tail -1 gpurun_out/gat9.log
python tools/bench_configs.py > gpurun_out/configs9.log 2>&1; echo "configs rc=$?"; cut -c1-300 gpurun_out/configs9.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain9.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|stage_rows|pack_rows|row_scale" -c 200 --csv \
    --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu9a.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spmm_rows|spmm_long" -s 40 -c 4 -o gpurun_out/prof_spmm_v3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu9b.log 2>&1
echo "ncu spmm rc=$?"
