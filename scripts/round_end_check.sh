set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/v9_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v9_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v9_ref.json 2> gpurun_out/v9_ref.err
python bench.py > gpurun_out/v9_bench.json 2> gpurun_out/v9_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v9_launches.csv python bench.py --steps 6 --warmup 3 > gpurun_out/v9_ncu_bench.log 2>&1
cat gpurun_out/v9_gputests.log gpurun_out/v9_smoke.log; head -c 600 gpurun_out/v9_bench.json; echo; head -c 300 gpurun_out/v9_ref.json; echo; wc -l gpurun_out/v9_launches.csv
