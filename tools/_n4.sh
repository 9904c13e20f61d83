cd $GRAFT_REPO_ROOT
N=$1
for fg in 0 1; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$fg bench.py --gpus $N --steps 10 --warmup 3 --feature-groups $fg --no-epoch > gpurun_out/r3n_bench_n${N}_fg$fg.json 2> gpurun_out/r3n_bench_n${N}_fg$fg.err
echo rc=$?
python - <<PY
import json
for l in open("gpurun_out/r3n_bench_n${N}_fg$fg.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N fg=$fg", round(d["value"],1), "GTEPS", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d["config"]["parallelism"], "build", round(d["config"]["graph_build_ms"]), "nccl", round(d["config"].get("nccl_init_ms",0)), d.get("check"))
PY
done
