set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t4_pytest.log
cat gpurun_out/r2_t4_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -20 gpurun_out/r2_bench_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'], d['config']['exchange_verified_bit_equal_to_nccl_path'])
print('e2e', d['e2e']['value'], d['e2e']['synchronous_value'])
print('strong', json.dumps(d.get('strong'), indent=0))
print(d['one_frame_at_a_time_l2_flushed'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-700
