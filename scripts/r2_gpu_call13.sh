set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t13_pytest.log; cat gpurun_out/r2_t13_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
for wl in c1 c2 c4 c5; do
  timeout 400 python bench.py --workload $wl --steps 60 > gpurun_out/r2_bench_${wl}_n1.json 2> gpurun_out/r2_bench_${wl}.err; tail -2 gpurun_out/r2_bench_${wl}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_${wl}_n1.json'))
print('$wl', {k: d[k] for k in ('value','ms_per_step')}, d['config']['frames_in_flight'], d['config']['frame_path'][:12], 'e2e', d['e2e']['value'], d['one_frame_at_a_time_l2_flushed'], d['mrays_s'], d['roofline']['frac'])
"
done
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['frames_in_flight'], d['config']['calibration_ms_per_step'])
print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print(d['strong'])
print(d['next_rows']['device_builder'])
"
