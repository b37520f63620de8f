set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -15
for ex in peer peer-nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 3 --exchange $ex > gpurun_out/r2_bench_n2_$ex.json 2> gpurun_out/r2_bench_n2_$ex.err
grep -v "^+" gpurun_out/r2_bench_n2_$ex.err | tail -12
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2_$ex.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'], d['config']['exchange_verified_bit_equal_to_nccl_path'])
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
s=d.get('strong'); print('strong', s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['one_frame_at_a_time'])
print(d['one_frame_at_a_time_l2_flushed'])
"
done
