set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 60 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
grep -v "^+" gpurun_out/r2_bench_n4.err | tail -4
