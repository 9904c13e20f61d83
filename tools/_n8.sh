cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/r3o_bench_n8.json 2> gpurun_out/r3o_bench_n8.err
echo rc=$?
python - <<PY
import json
for l in open("gpurun_out/r3o_bench_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=8", round(d["value"],1), "GTEPS", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d["config"]["parallelism"], "build", round(d["config"]["graph_build_ms"]), "nccl", round(d["config"].get("nccl_init_ms",0)), d.get("check"), (d.get("epoch") or {}).get("epoch_ms"))
PY
