cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3m_bench_n2.json 2> gpurun_out/r3m_bench_n2.err
echo rc=$?
cut -c1-1200 gpurun_out/r3m_bench_n2.json
tail -5 gpurun_out/r3m_bench_n2.err
