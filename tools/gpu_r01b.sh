#!/bin/bash
# round-1 second GPU pass: tests, shape/cache-policy A/B, bench, ncu launch list + full capture of the v2 SpMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests4.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests4.log
python tools/sweep.py --workloads products,arxiv --windows -1 --minimal --us 2,18,20 > gpurun_out/sweep4.log 2>&1; echo "sweep rc=$?"
python tools/sweep.py --workloads products,arxiv --windows -1 --minimal --us 2,18,20 --nostream > gpurun_out/sweep4_nostream.log 2>&1; echo "sweep-nostream rc=$?"
grep BEST gpurun_out/sweep4.log; echo ----; grep BEST gpurun_out/sweep4_nostream.log
python bench.py > gpurun_out/bench4.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench4.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench4_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench4_ref.log
python tools/bench_configs.py > gpurun_out/configs4.log 2>&1; echo "configs rc=$?"; tail -20 gpurun_out/configs4.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|stage_rows|pack_rows" -c 200 --csv \
    --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4a.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 35 -c 2 -o gpurun_out/prof_spmm_v2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4b.log 2>&1
echo "ncu full rc=$?"
