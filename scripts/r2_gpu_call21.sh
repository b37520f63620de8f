set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 60 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep -v "^+" gpurun_out/r2_bench_n8.err | tail -6
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['exchange_verified_bit_equal_to_nccl_path'], d['config']['exchange'][:30])
for k,v in d['config']['calibration_ms_per_step'].items(): print(' ', k, round(v,4))
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
for k in ('strong','strong_c5'):
    s=d.get(k); print(k, s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['frames_in_flight'], s['exchange'], s['one_frame_at_a_time'])
    for ex, r in s['by_exchange'].items():
        print('   ', ex, {m: (round(v['ms_per_step'],4), round(v['speedup_vs_n1'],3)) for m, v in r.items()})
"
timeout 300 python bench.py --gpus 8 --single-process --steps 60 > gpurun_out/r2_bench_n8_single_process.json 2> gpurun_out/r2_bench_n8_sp.err
timeout 300 python bench.py --gpus 8 --single-process --steps 60 --workload c4 > gpurun_out/r2_bench_n8_single_process_c4.json 2>> gpurun_out/r2_bench_n8_sp.err
tail -3 gpurun_out/r2_bench_n8_sp.err
python -c "
import json
for f in ('gpurun_out/r2_bench_n8_single_process.json','gpurun_out/r2_bench_n8_single_process_c4.json'):
    d=json.load(open(f))
    print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['by_frames_in_flight'], d['per_frame_ms_cuda_events'], d['e2e']['value'], d['e2e']['frame_in_host_memory_equals_single_gpu_frame'])
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 60 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n4.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['exchange'][:20], 'e2e', d['e2e']['value'])
s=d.get('strong'); print('strong', s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['frames_in_flight'], s['exchange'])
"
