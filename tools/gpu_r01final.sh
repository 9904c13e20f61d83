#!/bin/bash
# round-1 final validation: full GPU suite, smoke, bench.py (both arms), config timings incl. C1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/f_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-700 gpurun_out/f_bench_ref.json
timeout 600 python tools/bench_configs.py --only c1,c2 > gpurun_out/f_configs.log 2>&1; echo "configs rc=$?"; grep -E "^\{" gpurun_out/f_configs.log | cut -c1-400
