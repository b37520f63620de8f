set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
grep -v "^+" gpurun_out/r2_bench_n2.err | tail -5
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['exchange'][:30], d['config']['exchange_verified_bit_equal_to_nccl_path'], 'e2e', d['e2e']['value'])
s=d.get('strong'); print('strong', s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['frames_in_flight'], s['exchange'])
"
timeout 300 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
