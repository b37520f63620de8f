set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -5
timeout 300 python scripts/r2_perf.py sanmiguel --w 3840 --h 2160 --shards 8 --device-build 2>&1 | grep -v "^+"
timeout 300 python scripts/r2_perf.py caldera --w 3840 --h 2160 --shards 8 --tlas 2>&1 | grep -v "^+"
timeout 300 python scripts/r2_perf.py hairball 2>&1 | grep -v "^+"
