cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_khop_cta.py tests/test_gpu_khop.py tests/test_gpu_hand_pins.py tests/test_pta_patch.py -x -q > gpurun_out/r3i_pytest.txt 2>&1
tail -3 gpurun_out/r3i_pytest.txt
timeout 300 python tools/khop_latency.py > gpurun_out/r3i_khop_latency.txt 2>&1
grep -v chunk gpurun_out/r3i_khop_latency.txt | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{\"case'): continue
    d=json.loads(l); print(d['case'],'|',d['call'][:12],'|',d['path'][:12],d['us_per_call_queued'],d['us_per_call_synchronised'],d['us_per_hop_queued'])"
