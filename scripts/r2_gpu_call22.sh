set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t22_pytest.log; cat gpurun_out/r2_t22_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['frame_path'][:10])
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'frac', d['roofline']['frac'], d['roofline']['issue']['frac_of_issue_peak'])
print(d['next_rows']['single_process_group'])
r=json.load(open('gpurun_out/r2_bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['cores'])
"
TRAY_BENCH_STRONG=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_bench_launches.csv
timeout 300 python scripts/sweep_knobs.py hairball "{'TRAY_CUDA_TRI_WEIGHT': [2, 3, 4, 6], 'TRAY_CUDA_REFILL_MIN': [2, 4, 8]}" 2>&1 | grep -v "^+"
