set -x
cd "$GRAFT_REPO_ROOT"
python -c "
from tray_racing_b200 import cuda
for kb in (8, 16, 32, 64):
    print('l1 gather', kb, 'KiB', [round(cuda.l1_gather_probe(kb << 10, 200), 1) for _ in range(2)], 'GB/s')
print('l2', cuda.bandwidth_probe(48 << 20, 50))
" 2>&1 | tee gpurun_out/r2_l1_probe.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t26_pytest.log; cat gpurun_out/r2_t26_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['frame_path'][:10])
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'frac', d['roofline']['frac'], d['roofline']['issue']['frac_of_issue_peak'], 'l1', d['roofline']['l1_gather'])
print(d['config']['calibration_ms_per_step'])
r=json.load(open('gpurun_out/r2_bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['cores'])
"
TRAY_BENCH_STRONG=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_bench_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o gpurun_out/r2_hairball python scripts/render_frames.py --scene hairball --frames 3 > gpurun_out/r2_ncu_full.log 2>&1
TRAY_CUDA_OVERLAP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2_hairball_frame python scripts/render_frames.py --scene hairball --frames 3 >> gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
