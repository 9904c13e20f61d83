#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests11.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests11.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/smoke11.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke11.log
