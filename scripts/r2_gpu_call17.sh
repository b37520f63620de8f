set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -5
for ex in push peer; do
echo "== exchange $ex"
TRAY_BENCH_STRONG_EXCHANGE=$ex timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/strong_only.py c4 2>&1 | grep -E "^\{|Error|error|Traceback" | cut -c1-900
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_bench_n2_push.json 2> gpurun_out/r2_bench_n2_push.err
grep -v "^+" gpurun_out/r2_bench_n2_push.err | tail -8
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2_push.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'], d['config']['exchange_verified_bit_equal_to_nccl_path'], d['config']['exchange'][:40])
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
s=d.get('strong'); print('strong', s['ms_per_step'], s['n1_ms_per_step'], s['speedup_vs_n1'], s['frames_in_flight'], s['one_frame_at_a_time'])
"
