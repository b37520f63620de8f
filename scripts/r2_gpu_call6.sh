set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L | wc -l
timeout 120 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 60 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep -v "^+" gpurun_out/r2_bench_n8.err | tail -12
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'], d['config']['exchange_verified_bit_equal_to_nccl_path'])
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
s=d.get('strong'); print('strong', s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['one_frame_at_a_time'])
print(d['one_frame_at_a_time_l2_flushed'])
"
timeout 300 python bench.py --gpus 8 --single-process --steps 60 > gpurun_out/r2_bench_n8_single_process.json 2> gpurun_out/r2_bench_n8_sp.err
tail -3 gpurun_out/r2_bench_n8_sp.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8_single_process.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['by_frames_in_flight'], d['per_frame_ms_cuda_events'], d['e2e'])
"
timeout 300 python bench.py --gpus 8 --single-process --steps 60 --workload c4 > gpurun_out/r2_bench_n8_single_process_c4.json 2>> gpurun_out/r2_bench_n8_sp.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8_single_process_c4.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['by_frames_in_flight'], d['per_frame_ms_cuda_events'], d['e2e'])
"
