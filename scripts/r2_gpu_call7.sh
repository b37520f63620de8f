set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t7_pytest.log; cat gpurun_out/r2_t7_pytest.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>/dev/null
TRAY_BENCH_STRONG=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o gpurun_out/r2_hairball python scripts/render_frames.py --scene hairball --frames 3 > gpurun_out/r2_ncu_full.log 2>&1
TRAY_CUDA_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/r2_hairball_frame python scripts/render_frames.py --scene hairball --frames 3 >> gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'])
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print(d['roofline']['frac'], d['roofline']['issue']['frac_of_issue_peak'], d['mrays_s'])
"
head -c 400 gpurun_out/r2_bench_reference.json
