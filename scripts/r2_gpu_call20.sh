set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -5
for ex in 1 0; do
TRAY_CUDA_GROUP_EXCHANGE=$ex timeout 300 python bench.py --gpus 2 --single-process --steps 40 --workload c4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('group exchange $ex c4', d['value'], d['ms_per_step'], d['by_frames_in_flight'], d['e2e']['value'], d['e2e']['frame_in_host_memory_equals_single_gpu_frame'])"
done
