set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 60 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep -v "^+" gpurun_out/r2_bench_n8.err | tail -6
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['exchange_verified_bit_equal_to_nccl_path'], d['config']['exchange'][:30])
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
for k in ('strong','strong_c5'):
    s=d.get(k); print(k, s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['frames_in_flight'], s['exchange'], s['one_frame_at_a_time'])
"
